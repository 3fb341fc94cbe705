"""Import shim for the *unmodified* reference tree (TEST INFRASTRUCTURE ONLY).

The reference (sfchng/neural_invertible_warp) is pure Python/PyTorch and imports a handful
of packages that are not installed in this image (easydict, lpips, visdom, ipdb, termcolor,
roma, imageio, matplotlib).  None of them takes part in the render arithmetic except
``roma.rigid_points_registration`` (Kabsch fit, off the render path).  This module installs
minimal stand-ins into ``sys.modules`` and puts the reference tree on ``sys.path`` so that
``oracle/make_golden.py`` can execute the real reference on CPU and mint golden vectors.

It is only usable where the reference tree exists (this build container: /root/reference).
Nothing under ``neural_invertible_warp_b200/`` imports this file; the GPU box never sees the
reference, so tests that need it are skipped there.
"""
import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NIW_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "nerf.py"))


class _AttrDict(dict):
    """Attribute-style dict with recursive conversion (stand-in for easydict.EasyDict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            return _AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(_AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v


def _kabsch(target, source):
    """roma.rigid_points_registration(x=target?, ...) stand-in.

    roma's signature is rigid_points_registration(x, y) -> (R, t) with y ~ R x + t.  The
    reference calls it as (target, source) (model/nerf_inn_llff.py:569), i.e. it fits
    source ~ R target + t.  Plain batched Kabsch with a determinant fix.
    """
    import torch
    x, y = target, source
    xm, ym = x.mean(dim=-2, keepdim=True), y.mean(dim=-2, keepdim=True)
    xc, yc = x - xm, y - ym
    M = yc.transpose(-1, -2) @ xc
    U, _, Vh = torch.linalg.svd(M)
    d = torch.det(U @ Vh)
    D = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], dim=-1))
    R = U @ D @ Vh
    t = (ym.squeeze(-2) - (R @ xm.transpose(-1, -2)).squeeze(-1))
    return R, t


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Permissive(types.ModuleType):
    """A module whose every attribute is another permissive, callable object."""
    __path__ = []  # looks like a package so that 'import a.b.c' recurses into the finder

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Permissive(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return self

    def __mro_entries__(self, bases):
        return (object,)


def _install_permissive_packages(roots):
    import importlib.abc
    import importlib.machinery

    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path=None, target=None):
            if fullname.split(".")[0] in roots:
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            return _Permissive(spec.name)

        def exec_module(self, module):
            pass

    sys.meta_path.insert(0, _Finder())


_installed = False


def install():
    """Install the stubs and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if "easydict" not in sys.modules:
        try:
            import easydict  # noqa: F401
        except ImportError:
            _stub("easydict", EasyDict=_AttrDict)
    for name in ("lpips", "visdom", "ipdb", "termcolor", "roma", "imageio"):
        try:
            importlib.import_module(name)
        except ImportError:
            pass

    class _LPIPS:
        def __init__(self, *a, **k):
            pass

        def to(self, *a, **k):
            return self

    if "lpips" not in sys.modules:
        _stub("lpips", LPIPS=_LPIPS)
    if "visdom" not in sys.modules:
        _stub("visdom", Visdom=object)
    if "ipdb" not in sys.modules:
        _stub("ipdb", set_trace=lambda *a, **k: None)
    if "termcolor" not in sys.modules:
        _stub("termcolor", colored=lambda s, *a, **k: str(s))
    if "roma" not in sys.modules:
        _stub("roma", rigid_points_registration=_kabsch)
    if "imageio" not in sys.modules:
        _stub("imageio")
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        _install_permissive_packages(("matplotlib", "mpl_toolkits"))
    # barf_inn_dtu eagerly imports the COLMAP/hloc initialiser (pycolmap, cupy, h5py ...)
    if "utils.colmap_initialization.sfm" not in sys.modules:
        _stub("utils.colmap_initialization.sfm", compute_sfm_pdcnet=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


@contextlib.contextmanager
def in_reference_dir():
    """The reference resolves options/*.yaml relative to the cwd (options.py:46,59-63)."""
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def load_reference_options(yaml_name, model, overrides=None, parent_override=None):
    """Build an ``opt`` with the reference's own loader (options.py:54-85).

    ``barf_llff.yaml`` inherits from ``options/nerf_llff.yaml`` which is missing upstream;
    ``parent_override`` substitutes ``nerf_inn_llff.yaml`` (identical arch/nerf/camera fields).
    """
    install()
    import yaml
    with in_reference_dir(), contextlib.redirect_stdout(open(os.devnull, "w")):
        import options as ref_options
        if parent_override is None:
            opt = ref_options.load_options("options/%s.yaml" % yaml_name)
        else:
            with open("options/%s.yaml" % yaml_name) as f:
                child = _AttrDict(yaml.safe_load(f))
            child.pop("_parent_", None)
            parent = ref_options.load_options("options/%s.yaml" % parent_override)
            opt = ref_options.override_options(parent, child, key_stack=[])
    opt = _AttrDict(opt)
    opt.model = model
    opt.yaml = yaml_name
    if overrides:
        _deep_update(opt, overrides)
    opt.device = "cpu"
    opt.H, opt.W = opt.data.image_size
    return opt


def _deep_update(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_update(dst[k], v)
        else:
            dst[k] = v


def import_reference(module):
    install()
    with in_reference_dir(), contextlib.redirect_stdout(open(os.devnull, "w")):
        return importlib.import_module(module)
