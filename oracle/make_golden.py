"""Mint golden vectors by executing the UNMODIFIED reference on CPU (build container only).

    python oracle/make_golden.py            # writes tests/golden/*.pt

Inputs are produced by ``neural_invertible_warp_b200.synthetic`` from fixed seeds (only the seeds
and small index/uniform tensors are stored); outputs are what the reference's own functions
return (model/nerf.py, model/barf.py, camera.py, model/nvp/nvp_ndr.py, model/barf_inn_llff.py,
model/barf_inn_dtu.py).  torch.rand / torch.randperm calls made inside the reference are
recorded so that the same draws can be replayed to the oracle port and to the CUDA path.
"""
import contextlib
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from neural_invertible_warp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


@contextlib.contextmanager
def record_rng():
    """Record every torch.rand / torch.randperm result drawn inside the block."""
    log = dict(rand=[], randperm=[])
    o_rand, o_perm = torch.rand, torch.randperm

    def rand(*a, **k):
        t = o_rand(*a, **k)
        log["rand"].append(t.clone())
        return t

    def randperm(*a, **k):
        t = o_perm(*a, **k)
        log["randperm"].append(t.clone())
        return t

    torch.rand, torch.randperm = rand, randperm
    try:
        yield log
    finally:
        torch.rand, torch.randperm = o_rand, o_perm


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def load_nerf(module, p):
    sd = module.state_dict()
    for k, v in p.items():
        assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
    module.load_state_dict({**{k: v for k, v in sd.items() if k not in p}, **p})


def grad_digest(named):
    """Small per-tensor summary of gradients: (sum, abs-sum, l2) + the first 8 entries."""
    out = {}
    for k, g in named.items():
        g = g.detach().double().flatten()
        out[k] = dict(sum=g.sum().item(), abssum=g.abs().sum().item(), l2=g.norm().item(),
                      head=g[:8].float().clone(), numel=g.numel())
    return out


# ------------------------------------------------------------------------------------------


def golden_camera():
    camera = ref_shim.import_reference("camera")
    opt = ref_shim._AttrDict(H=6, W=8, device="cpu", camera=dict(model="perspective", ndc=False))
    B = 3
    pose = syn.llff_poses(11, B, noise=0.2)
    pose[..., 3] += torch.tensor([0.3, -0.2, 0.5])
    intr = syn.intrinsics(B, opt.H, opt.W, 0.81)
    intr[1, 0, 0] *= 1.1
    intr[2, 0, 2] += 0.7
    ray_idx = syn.ray_indices(5, opt.H, opt.W, 10)
    center, ray = camera.get_center_and_ray(opt, pose, intr=intr)
    c_cam, g_cam = camera.get_unwarped_center_and_ray(opt, intr=intr, ray_idx=ray_idx)
    c_w, g_w = camera.get_unwarped_center_and_ray(opt, intr=intr, ray_idx=ray_idx, pose_init=pose)
    wu = torch.randn(4, 6, generator=torch.Generator().manual_seed(3)) * 0.3
    SE3 = camera.lie.se3_to_SE3(wu)
    comp = camera.pose.compose([SE3[:3], pose])
    c_ndc, r_ndc = camera.convert_NDC(opt, center[:, ray_idx], ray[:, ray_idx], intr=intr)
    # pose gradient through ray generation
    pose_g = pose.clone().requires_grad_(True)
    c2, r2 = camera.get_center_and_ray(opt, pose_g, intr=intr)
    wc = syn.uniforms(21, *c2.shape) - 0.5
    wr = syn.uniforms(22, *r2.shape) - 0.5
    ((c2 * wc).sum() + (r2 * wr).sum()).backward()
    save("camera", dict(H=opt.H, W=opt.W, B=B, pose_seed=11, pose=pose, intr=intr, ray_idx=ray_idx,
                        center=center, ray=ray, center_cam=c_cam, grid_cam=g_cam, center_w=c_w,
                        grid_w=g_w, wu=wu, SE3=SE3, composed=comp, center_ndc=c_ndc, ray_ndc=r_ndc,
                        pose_grad=pose_g.grad.clone(), wc_seed=21, wr_seed=22))


def _opt(yaml_name, model, parent=None, **over):
    return ref_shim.load_reference_options(yaml_name, model, overrides=over, parent_override=parent)


def golden_sampler():
    nerf_mod = ref_shim.import_reference("model.nerf")
    out = {}
    for tag, yaml_name, N, Nf, rng, param in (("llff", "nerf_inn_llff", 128, 64, [1, 0], "inverse"),
                                               ("dtu", "nerf_inn_dtu", 64, 128, [1.2, 5.2], "metric")):
        opt = _opt(yaml_name, "nerf", nerf=dict(sample_intvs=N, sample_intvs_fine=Nf,
                                                 depth=dict(range=rng, param=param)))
        g = nerf_mod.Graph.__new__(nerf_mod.Graph)
        B, R = 2, 6
        torch.manual_seed(100)
        with record_rng() as log:
            depth = nerf_mod.Graph.sample_depth(g, opt, B, num_rays=R)
        # a smooth un-normalised pdf like a composited ray: some rays opaque, some not
        gen = torch.Generator().manual_seed(7)
        x = torch.linspace(0, 1, N)[None, None]
        mu = torch.rand(B, R, 1, generator=gen)
        sg = 0.02 + 0.2 * torch.rand(B, R, 1, generator=gen)
        amp = torch.rand(B, R, 1, generator=gen) * 1.2
        pdf = torch.exp(-0.5 * ((x - mu) / sg) ** 2)
        pdf = pdf / pdf.sum(-1, keepdim=True) * amp.clamp(max=1.0)
        pdf[0, 0] = 0.0  # fully transparent ray
        pdf[0, 1, : N // 2] = 0.0  # leading zeros -> repeated cdf values
        fine = nerf_mod.Graph.sample_depth_from_pdf(g, opt, pdf=pdf)
        # recover idx exactly as the reference computes it (model/nerf.py:348-354)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), pdf.cumsum(dim=-1)], dim=-1)
        grid = torch.linspace(0, 1, Nf + 1)
        unif = 0.5 * (grid[:-1] + grid[1:]).repeat(*cdf.shape[:-1], 1)
        idx = torch.searchsorted(cdf, unif, right=True)
        merged = torch.cat([depth, fine], dim=2).sort(dim=2).values
        out[tag] = dict(N=N, Nf=Nf, range=rng, param=param, u=log["rand"][0], depth=depth, pdf=pdf,
                        fine=fine, idx=idx, merged=merged)
    save("sampler", out)


def golden_nerf_mlp():
    barf = ref_shim.import_reference("model.barf")
    opt = _opt("barf_llff", "barf", parent="nerf_inn_llff", barf_c2f=[0.1, 0.5])
    net = barf.NeRF(opt)
    p = syn.nerf_params(42)
    load_nerf(net, p)
    gen = torch.Generator().manual_seed(9)
    pts = (torch.rand(2, 3, 8, 3, generator=gen) * 2 - 1) * 1.5
    unit = torch.nn.functional.normalize(torch.randn(2, 3, 1, 3, generator=gen), dim=-1).expand_as(pts).contiguous()
    cases = {}
    for prog in (0.2, 0.3, 1.0):
        net.progress.data.fill_(prog)
        pts_g = pts.clone().requires_grad_(True)
        unit_g = unit.clone().requires_grad_(True)
        net.zero_grad()
        rgb, dens = net.forward(opt, pts_g, ray_unit=unit_g, mode="train")
        wr = syn.uniforms(31, *rgb.shape) - 0.5
        wd = syn.uniforms(32, *dens.shape) - 0.5
        ((rgb * wr).sum() + (dens * wd).sum()).backward()
        cases[prog] = dict(rgb=rgb.detach().clone(), density=dens.detach().clone(),
                           d_points=pts_g.grad.clone(), d_unit=unit_g.grad.clone(),
                           grads=grad_digest({k: v.grad for k, v in net.named_parameters() if v.grad is not None}))
        enc = net.positional_encoding(opt, pts, L=10)
        cases[prog]["enc_head"] = enc[0, 0, :2].clone()
    opt2 = copy.deepcopy(opt)
    opt2.barf_c2f = None
    rgb, dens = net.forward(opt2, pts, ray_unit=unit, mode="eval")
    cases["no_c2f"] = dict(rgb=rgb.detach().clone(), density=dens.detach().clone())
    save("nerf_mlp", dict(param_seed=42, points=pts, ray_unit=unit, c2f=[0.1, 0.5], cases=cases,
                          wr_seed=31, wd_seed=32))


def golden_composite():
    nerf_mod = ref_shim.import_reference("model.nerf")
    opt = _opt("nerf_inn_llff", "nerf")
    net = nerf_mod.NeRF.__new__(nerf_mod.NeRF)
    gen = torch.Generator().manual_seed(17)
    B, R, N = 2, 5, 32
    ray = torch.randn(B, R, 3, generator=gen)
    rgb_s = torch.rand(B, R, N, 3, generator=gen)
    sig = torch.rand(B, R, N, generator=gen) * 3
    sig[0, 0] = 0
    sig[0, 1, 3] = 500.0
    depth = (torch.rand(B, R, N, 1, generator=gen) + torch.arange(N)[None, None, :, None]) / N * 4 + 1
    ins = [t.clone().requires_grad_(True) for t in (ray, rgb_s, sig)]
    rgb, d, op, prob = nerf_mod.NeRF.composite(net, opt, ins[0], ins[1], ins[2], depth)
    w = [syn.uniforms(40 + i, *t.shape) - 0.5 for i, t in enumerate((rgb, d, op))]
    ((rgb * w[0]).sum() + (d * w[1]).sum() + (op * w[2]).sum()).backward()
    save("composite", dict(ray=ray, rgb_samples=rgb_s, sigma=sig, depth_samples=depth, rgb=rgb.detach(),
                           depth=d.detach(), opacity=op.detach(), prob=prob.detach(), w_seeds=[40, 41, 42],
                           d_ray=ins[0].grad, d_rgb_samples=ins[1].grad, d_sigma=ins[2].grad))


def golden_nvp():
    nvp = ref_shim.import_reference("model.nvp.nvp_ndr")
    net = nvp.DeformNetwork(d_feature=128, d_in=3, d_out_1=1, d_out_2=3, n_blocks=3, d_hidden=128,
                            n_layers=1, skip_in=[], multires=6, weight_norm=True, actfn="softplus")
    p = syn.nvp_params(77)
    sd = net.state_dict()
    assert set(sd.keys()) == set(p.keys()), set(sd.keys()) ^ set(p.keys())
    net.load_state_dict(p)
    B, P = 2, 40
    code = syn.latent_codes(78, B)
    gen = torch.Generator().manual_seed(79)
    pts = torch.randn(B, P, 1, 3, generator=gen) * 0.6
    pts[:, P // 2:] = 0.0  # the "camera centre" rows are the origin
    cases = {}
    for alpha in (0.05, 0.4, 1.0):
        net.zero_grad()
        code_g = code.clone().requires_grad_(True)
        out = net.forward(code_g, pts, alpha_ratio=alpha)
        w = syn.uniforms(80, *out.shape) - 0.5
        (out * w).sum().backward()
        grads = {k: v.grad for k, v in net.named_parameters()}
        cases[alpha] = dict(out=out.detach().clone(), d_code=code_g.grad.clone(), grads=grad_digest(grads),
                            d_a1w=grads["lin0_a_1.weight"].clone(), d_b1w=grads["lin2_b_1.weight"].clone())
    save("nvp", dict(param_seed=77, code_seed=78, pts=pts, w_seed=80, cases=cases))


def _synthetic_var(opt, B, seed, dtu=False):
    edict = ref_shim._AttrDict
    H, W = opt.H, opt.W
    var = edict(idx=torch.arange(B), image=syn.images(seed, B, H, W),
                intr=syn.intrinsics(B, H, W, 1.8 if dtu else 0.81),
                pose=syn.dtu_poses(seed + 1, B) if dtu else syn.llff_poses(seed + 1, B))
    if dtu:
        var.depth_range = torch.tensor([[1.2, 5.2]]).repeat(B, 1)
    return var


def _finish_graph(graph, opt, var, mode, it=None):
    torch.manual_seed(1000)
    with record_rng() as log:
        var = graph.forward(opt, var, mode=mode) if it is None else graph.forward(opt, var, mode=mode, iter=it)
        loss = graph.compute_loss(opt, var, mode=mode)
    total = sum(10 ** float(opt.loss_weight[k]) * loss[k] for k in loss if opt.loss_weight.get(k) is not None)
    graph.zero_grad()
    total.backward()
    return var, loss, total, log


def golden_graph_barf():
    barf = ref_shim.import_reference("model.barf")
    opt = _opt("barf_llff", "barf", parent="nerf_inn_llff", barf_c2f=[0.1, 0.5],
               data=dict(image_size=[24, 32]), nerf=dict(rand_rays=32, sample_intvs=16))
    B = 2
    graph = barf.Graph(opt)
    load_nerf(graph.nerf, syn.nerf_params(50))
    graph.nerf.progress.data.fill_(0.3)
    graph.se3_refine = torch.nn.Embedding(B, 6)
    graph.se3_refine.weight.data = torch.randn(B, 6, generator=torch.Generator().manual_seed(51)) * 0.05
    var = _synthetic_var(opt, B, 52)
    var, loss, total, log = _finish_graph(graph, opt, var, "train")
    save("graph_barf", dict(B=B, H=24, W=32, N=16, rand_rays=32, param_seed=50, se3_seed=51, var_seed=52,
                            progress=0.3, se3=graph.se3_refine.weight.data.clone(),
                            ray_idx=log["randperm"][0][:32 // B], u=log["rand"][0], rgb=var.rgb.detach(),
                            depth=var.depth.detach(), opacity=var.opacity.detach(), loss=total.detach(),
                            d_se3=graph.se3_refine.weight.grad.clone(),
                            grads=grad_digest({k: v.grad for k, v in graph.nerf.named_parameters() if v.grad is not None})))


def golden_graph_nerf():
    """Plain NeRF with given poses (model/nerf.py Graph :243-365, NeRF :367-483): no pose refinement, no c2f weighting."""
    nerf = ref_shim.import_reference("model.nerf")
    opt = _opt("nerf_inn_llff", "nerf", data=dict(image_size=[24, 32]), nerf=dict(rand_rays=40, sample_intvs=16))
    B = 2
    graph = nerf.Graph(opt)
    load_nerf(graph.nerf, syn.nerf_params(55))
    var = _synthetic_var(opt, B, 56)
    var, loss, total, log = _finish_graph(graph, opt, var, "train")
    save("graph_nerf", dict(B=B, H=24, W=32, N=16, rand_rays=40, param_seed=55, var_seed=56,
                            ray_idx=log["randperm"][0][:40 // B], u=log["rand"][0], rgb=var.rgb.detach(),
                            depth=var.depth.detach(), opacity=var.opacity.detach(), loss=total.detach(),
                            keys=sorted(graph.state_dict().keys()),
                            grads=grad_digest({k: v.grad for k, v in graph.nerf.named_parameters() if v.grad is not None})))


def golden_graph_inn_llff():
    mod = ref_shim.import_reference("model.barf_inn_llff")
    nvp = ref_shim.import_reference("model.nvp.nvp_ndr")
    out = {}
    for tag, rays_per_img, it in (("p16", 16, 5000), ("p40", 40, 40000)):
        B = 2
        opt = _opt("barf_inn_llff", "barf_inn_llff", barf_c2f=[0.1, 0.5], data=dict(image_size=[24, 32]),
                   nerf=dict(rand_rays=rays_per_img * B, sample_intvs=16),
                   loss_weight=dict(global_alignment=2))
        graph = mod.Graph(opt)
        load_nerf(graph.nerf, syn.nerf_params(60))
        graph.nerf.progress.data.fill_(0.3)
        graph.warp_latent = torch.nn.Embedding(B, 128)
        graph.warp_latent.weight.data = syn.latent_codes(61, B)
        graph.warp_mlp = nvp.DeformNetwork(d_feature=128, d_in=3, d_out_1=1, d_out_2=3, n_blocks=3, d_hidden=128,
                                           n_layers=1, skip_in=[], multires=6, weight_norm=True, actfn="softplus")
        graph.warp_mlp.load_state_dict(syn.nvp_params(62))
        graph.global_rigid = torch.nn.Embedding(B, 12)
        var = _synthetic_var(opt, B, 63)
        var, loss, total, log = _finish_graph(graph, opt, var, "train", it=it)
        out[tag] = dict(B=B, H=24, W=32, N=16, rays_per_img=rays_per_img, iter=it, nerf_seed=60, code_seed=61,
                        nvp_seed=62, var_seed=63, progress=0.3, ray_idx=log["randperm"][0][:rays_per_img],
                        u=log["rand"][0], rgb=var.rgb.detach(), depth=var.depth.detach(),
                        opacity=var.opacity.detach(), grid_3D=var.grid_3D.detach(), center=var.center.detach(),
                        alpha_ratio=var.inn_posenc_alpha, loss_render=loss.render.detach(),
                        loss_global_alignment=loss.global_alignment.detach(), loss=total.detach(),
                        global_rigid=graph.global_rigid.weight.data.clone(),
                        d_code=graph.warp_latent.weight.grad.clone(),
                        nvp_grads=grad_digest({k: v.grad for k, v in graph.warp_mlp.named_parameters()}),
                        grads=grad_digest({k: v.grad for k, v in graph.nerf.named_parameters() if v.grad is not None}))
    save("graph_inn_llff", out)


def golden_graph_inn_dtu():
    mod = ref_shim.import_reference("model.barf_inn_dtu")
    inn = ref_shim.import_reference("model.pose_models.inn")
    B = 2
    opt = _opt("barf_inn_dtu", "barf_inn_dtu", barf_c2f=[0.1, 0.5], data=dict(image_size=[18, 24]),
               nerf=dict(rand_rays=30 * B, sample_intvs=16, fine_sampling=True, sample_intvs_fine=32,
                         depth=dict(range=[1.2, 5.2])),
               loss_weight=dict(render_fine=0))
    var = _synthetic_var(opt, B, 73, dtu=True)
    pose_net = inn.INNPoseParams(opt, B, var.pose.clone(), device="cpu")
    pose_net.pose_latent.weight.data = syn.latent_codes(71, B)
    pose_net.pose_embedding.load_state_dict(syn.nvp_params(72))
    graph = mod.Graph(opt, pose_net)
    load_nerf(graph.nerf, syn.nerf_params(70))
    load_nerf(graph.nerf_fine, syn.nerf_params(74))
    graph.nerf.progress.data.fill_(0.3)
    graph.nerf_fine.progress.data.fill_(0.3)
    var, loss, total, log = _finish_graph(graph, opt, var, "train", it=30000)
    save("graph_inn_dtu", dict(B=B, H=18, W=24, N=16, Nf=32, rays_per_img=30, iter=30000, nerf_seed=70,
                               nerf_fine_seed=74, code_seed=71, nvp_seed=72, var_seed=73, progress=0.3,
                               ray_idx=log["randperm"][0][:30], u=log["rand"][0], rgb=var.rgb.detach(),
                               depth=var.depth.detach(), opacity=var.opacity.detach(),
                               rgb_fine=var.rgb_fine.detach(), depth_fine=var.depth_fine.detach(),
                               opacity_fine=var.opacity_fine.detach(), loss=total.detach(),
                               pose_global=pose_net.pose_global.weight.data.clone(),
                               d_code=pose_net.pose_latent.weight.grad.clone(),
                               nvp_grads=grad_digest({k: v.grad for k, v in pose_net.pose_embedding.named_parameters()}),
                               grads=grad_digest({k: v.grad for k, v in graph.nerf.named_parameters() if v.grad is not None}),
                               grads_fine=grad_digest({k: v.grad for k, v in graph.nerf_fine.named_parameters() if v.grad is not None})))


def golden_state_dicts():
    """Checkpoint inventory: ``graph.state_dict()`` keys and shapes of the reference Graphs as the reference's engine
    builds them (util.save_checkpoint stores exactly this dict, util.py:147-156)."""
    out = {}
    barf = ref_shim.import_reference("model.barf")
    opt = _opt("barf_llff", "barf", parent="nerf_inn_llff", barf_c2f=[0.1, 0.5], data=dict(image_size=[24, 32]))
    g = barf.Graph(opt)
    g.se3_refine = torch.nn.Embedding(3, 6)                      # model/barf.py:42
    out["barf"] = {k: list(v.shape) for k, v in g.state_dict().items()}
    mod = ref_shim.import_reference("model.barf_inn_llff")
    nvp = ref_shim.import_reference("model.nvp.nvp_ndr")
    opt = _opt("barf_inn_llff", "barf_inn_llff", barf_c2f=[0.1, 0.5], data=dict(image_size=[24, 32]))
    g = mod.Graph(opt)
    g.warp_latent = torch.nn.Embedding(3, 128)                   # model/barf_inn_llff.py:41-55
    g.warp_mlp = nvp.DeformNetwork(d_feature=128, d_in=3, d_out_1=1, d_out_2=3, n_blocks=3, d_hidden=128, n_layers=1,
                                   skip_in=[], multires=6, weight_norm=True, actfn="softplus")
    g.global_rigid = torch.nn.Embedding(3, 12)
    out["barf_inn_llff"] = {k: list(v.shape) for k, v in g.state_dict().items()}
    mod = ref_shim.import_reference("model.barf_inn_dtu")
    inn = ref_shim.import_reference("model.pose_models.inn")
    opt = _opt("barf_inn_dtu", "barf_inn_dtu", barf_c2f=[0.1, 0.5], data=dict(image_size=[18, 24]),
               nerf=dict(fine_sampling=True))
    var = _synthetic_var(opt, 3, 73, dtu=True)
    g = mod.Graph(opt, inn.INNPoseParams(opt, 3, var.pose.clone(), device="cpu"))
    out["barf_inn_dtu"] = {k: list(v.shape) for k, v in g.state_dict().items()}
    save("state_dicts", out)


def golden_test_optim():
    """Test-time photometric pose optimisation (model/barf.py:153-169): the reference's loop body, executed on the
    reference Graph for a few Adam steps on one held-out view; draws, per-step losses and the se3 trajectory recorded."""
    barf = ref_shim.import_reference("model.barf")
    camera = ref_shim.import_reference("camera")
    opt = _opt("barf_llff", "barf", parent="nerf_inn_llff", barf_c2f=[0.1, 0.5],
               data=dict(image_size=[24, 32]), nerf=dict(rand_rays=48, sample_intvs=16), optim=dict(test_photo=True))
    graph = barf.Graph(opt)
    load_nerf(graph.nerf, syn.nerf_params(80))
    graph.nerf.progress.data.fill_(1.0)
    gen = torch.Generator().manual_seed(81)
    Rm = camera.lie.so3_to_SO3(torch.randn(3, generator=gen) * 0.1)
    graph.sim3 = ref_shim._AttrDict(t0=torch.randn(1, 3, generator=gen) * 0.1, t1=torch.randn(1, 3, generator=gen) * 0.1,
                                    s0=torch.tensor(1.3), s1=torch.tensor(0.8), R=Rm)      # model/barf.py:113
    var = _synthetic_var(opt, 1, 82)
    lr = 1.e-2
    iters = 6
    var.se3_refine_test = torch.nn.Parameter(torch.zeros(1, 6))
    optim = torch.optim.Adam([dict(params=[var.se3_refine_test], lr=lr)])
    torch.manual_seed(1000)
    ridx, us, losses, traj, grads = [], [], [], [], []
    for it in range(iters):
        optim.zero_grad()
        var.pose_refine_test = camera.lie.se3_to_SE3(var.se3_refine_test)
        with record_rng() as log:
            var = graph.forward(opt, var, mode="test-optim")
            loss = graph.compute_loss(opt, var, mode="test-optim")
        total = sum(10 ** float(opt.loss_weight[k]) * loss[k] for k in loss if opt.loss_weight.get(k) is not None)
        total.backward()
        grads.append(var.se3_refine_test.grad.clone())
        optim.step()
        ridx.append(log["randperm"][0][:opt.nerf.rand_rays])
        us.append(log["rand"][0])
        losses.append(total.detach().clone())
        traj.append(var.se3_refine_test.data.clone())
    save("test_optim", dict(H=24, W=32, N=16, rand_rays=48, nerf_seed=80, var_seed=82, progress=1.0, lr=lr, iters=iters,
                            sim3=dict(t0=graph.sim3.t0, t1=graph.sim3.t1, s0=graph.sim3.s0, s1=graph.sim3.s1, R=graph.sim3.R),
                            ray_idx=torch.stack(ridx), u=torch.stack(us), losses=torch.stack(losses), se3=torch.stack(traj),
                            d_se3=torch.stack(grads)))


def golden_metrics():
    """PSNR / SSIM / depth error exactly as Model.evaluate_full computes them (model/nerf.py:176-183) from the
    reference's own pytorch_ssim and core/metrics.py, on a synthetic render / ground-truth pair."""
    pytorch_ssim = ref_shim.import_reference("external.pohsun_ssim.pytorch_ssim")
    metrics = ref_shim.import_reference("core.metrics")
    H, W, B = 37, 50, 2
    gen = torch.Generator().manual_seed(90)
    image = syn.images(91, B, H, W)
    rgb = (image.permute(0, 2, 3, 1).reshape(B, H * W, 3) + 0.1 * torch.randn(B, H * W, 3, generator=gen)).clamp(0, 1)
    out = dict(H=H, W=W, B=B, image_seed=91, rgb=rgb, psnr=[], ssim=[])
    for b in range(B):
        rgb_map = rgb[b:b + 1].view(-1, H, W, 3).permute(0, 3, 1, 2)
        out["psnr"].append(-10 * ((rgb_map.contiguous() - image[b:b + 1]) ** 2).mean().log10().item())
        out["ssim"].append(pytorch_ssim.ssim(rgb_map, image[b:b + 1]).item())
    depth_gt = torch.rand(1, H, W, generator=gen) * 3 + 1
    valid = torch.rand(1, H, W, generator=gen) > 0.3
    pred = (depth_gt.view(1, -1, 1) * 1.1 + 0.05 * torch.randn(1, H * W, 1, generator=gen))
    var = ref_shim._AttrDict(depth=pred, depth_gt=depth_gt, valid_depth_gt=valid)
    out.update(depth_gt=depth_gt, valid=valid, depth=pred, depth_scale=0.9,
               depth_err=metrics.compute_depth_error(var, 1.0), depth_err_scaled=metrics.compute_depth_error(var, 0.9))
    save("metrics", out)


def golden_eval_slices():
    """Full-frame evaluation render of the reference: ``Graph.forward(mode="eval")`` -> ``render_by_slices``
    (model/nerf.py:251-274,321-332) on a small frame with a ragged last slice, BARF model with the test poses mapped
    through a non-trivial sim3 (model/barf.py:235-246), un-stratified sampling (no draws), plus PSNR / SSIM of the frame."""
    barf = ref_shim.import_reference("model.barf")
    camera = ref_shim.import_reference("camera")
    pytorch_ssim = ref_shim.import_reference("external.pohsun_ssim.pytorch_ssim")
    H, W, B = 13, 18, 2
    opt = _opt("barf_llff", "barf", parent="nerf_inn_llff", barf_c2f=[0.1, 0.5], data=dict(image_size=[H, W]),
               nerf=dict(rand_rays=50, sample_intvs=16, sample_stratified=False), optim=dict(test_photo=False))
    graph = barf.Graph(opt)
    load_nerf(graph.nerf, syn.nerf_params(95))
    graph.nerf.progress.data.fill_(0.4)
    gen = torch.Generator().manual_seed(96)
    graph.sim3 = ref_shim._AttrDict(t0=torch.randn(1, 3, generator=gen) * 0.1, t1=torch.randn(1, 3, generator=gen) * 0.1,
                                    s0=torch.tensor(1.2), s1=torch.tensor(0.9),
                                    R=camera.lie.so3_to_SO3(torch.randn(3, generator=gen) * 0.1))
    var = _synthetic_var(opt, B, 97)
    with torch.no_grad():
        var = graph.forward(opt, var, mode="eval")
        rgb_map = var.rgb.view(-1, H, W, 3).permute(0, 3, 1, 2)
        psnr = [(-10 * ((rgb_map[b:b + 1].contiguous() - var.image[b:b + 1]) ** 2).mean().log10()).item() for b in range(B)]
        ssim = [pytorch_ssim.ssim(rgb_map[b:b + 1], var.image[b:b + 1]).item() for b in range(B)]
    save("eval_slices", dict(H=H, W=W, B=B, rand_rays=50, N=16, nerf_seed=95, var_seed=97, progress=0.4,
                             sim3=dict(t0=graph.sim3.t0, t1=graph.sim3.t1, s0=graph.sim3.s0, s1=graph.sim3.s1, R=graph.sim3.R),
                             rgb=var.rgb.clone(), depth=var.depth.clone(), opacity=var.opacity.clone(), psnr=psnr, ssim=ssim))


def golden_options():
    """Hot-path option fields of the YAMLs the target models use (checked against config.py)."""
    out = {}
    for name, parent in (("nerf_inn_llff", None), ("barf_inn_llff", None), ("barf_llff", "nerf_inn_llff"),
                         ("nerf_inn_dtu", None), ("barf_inn_dtu", None)):
        o = _opt(name, name, parent=parent)
        out[name] = {k: copy.deepcopy(dict(o[k])) if isinstance(o[k], dict) else o[k]
                     for k in ("arch", "nerf", "camera", "loss_weight", "barf_c2f", "max_iter", "inn", "warp_latent")
                     if k in o}
        out[name]["image_size"] = list(o.data.image_size)
    import json
    out = json.loads(json.dumps(out))
    save("options", out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["camera", "sampler", "nerf_mlp", "composite", "nvp", "graph_barf", "graph_nerf",
                             "graph_inn_llff", "graph_inn_dtu", "options", "state_dicts", "test_optim", "metrics", "eval_slices"]
    for w in which:
        globals()["golden_" + w]()
