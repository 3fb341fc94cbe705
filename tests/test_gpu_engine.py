"""GPU tests of the host orchestration in ``engine``: optimiser checkpoint / resume (reference util.py:124-163 persists
every optim* / sched* state), interchange with the reference's torch optimisers, the pose-LR warm-up scope
(model/barf_inn_llff.py:108-111), gradient delivery (autograd by default, in-kernel accumulation on opt-in, several
evaluations of one network per step) and the per-model evaluation gating (model/nerf.py:172, nerf_inn_dtu.py:217)."""
import copy

import pytest
import torch

from neural_invertible_warp_b200 import config as cfgmod
from neural_invertible_warp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, W, B, P, N = 48, 64, 4, 16, 32
MAX_ITER = 50


@pytest.fixture(scope="module")
def eng():
    from neural_invertible_warp_b200 import engine
    return engine


def load_nerf(module, p):
    sd = module.state_dict()
    module.load_state_dict({**{k: v for k, v in sd.items() if k not in p}, **{k: v.to(DEV) for k, v in p.items()}})


def make(eng, warmup=3, precision="fp32"):
    opt = cfgmod.builtin_options("barf_inn_llff", barf_c2f=[0.1, 0.5], device=DEV, data=dict(image_size=[H, W]),
                                 nerf=dict(rand_rays=B * P, sample_intvs=N), max_iter=MAX_ITER,
                                 optim=dict(warmup_pose=warmup), arch=dict(mlp_precision=precision))
    torch.manual_seed(0)
    graph = eng.build_graph(opt, B)
    load_nerf(graph.nerf, syn.nerf_params(21))
    graph.nerf.progress.data.fill_(0.3)
    graph.warp_latent.weight.data = syn.latent_codes(22, B).to(DEV)
    graph.warp_mlp.load_state_dict({k: v.to(DEV) for k, v in syn.nvp_params(23).items()})
    return opt, graph, eng.synthetic_var(opt, B, 24)


def draws(n, seed=9):
    gen = torch.Generator().manual_seed(seed)
    return [(torch.randperm(H * W, generator=gen)[:P].to(DEV), torch.rand(B, P, N, 1, generator=gen).to(DEV)) for _ in range(n)]


def flat_step(eng, opt, graph, var, fa, d, it):
    with eng.feed_draws(ray_idx=d[0], u=d[1]):
        loss = eng.train_step(opt, graph, cfgmod.AttrDict(var), it, bucket=fa)
    fa.step()
    return float(loss.all.detach())


def updates_close(g_a, g_b, g_0):
    """Adam normalises the update: an element whose gradient sits at the noise floor of the atomically accumulated sums may
    step +lr in one run and -lr in the other, so single elements may differ by a few lr while the update as a whole agrees."""
    pa, pb, p0 = dict(g_a.named_parameters()), dict(g_b.named_parameters()), dict(g_0.named_parameters())
    for n in pa:
        da, db = (pa[n] - p0[n]).detach().double(), (pb[n] - p0[n]).detach().double()
        assert (da - db).abs().max().item() <= 8e-3, n
        if da.norm().item() > 0:
            assert ((da - db).norm() / da.norm()).item() < 1e-1, (n, ((da - db).norm() / da.norm()).item())


def test_flat_adam_matches_reference_loop_with_pose_warmup(eng):
    """engine.FlatAdam + reference_optimizer_groups against the reference's loop written out with torch optimisers:
    Adam(nerf) + ExponentialLR, Adam([warp_mlp], [warp_latent]) + ExponentialLR with the linear warm-up applied to
    param_groups[0] ONLY (model/barf_inn_llff.py:84-120), progress.fill_(it / max_iter) after the pose step."""
    ds = draws(4)
    opt, g1, var1 = make(eng)
    o = opt.optim
    optim = torch.optim.Adam([dict(params=g1.nerf.parameters(), lr=o.lr)])
    optim_pose = torch.optim.Adam([dict(params=g1.warp_mlp.parameters(), lr=o.lr_pose)])
    optim_pose.add_param_group(dict(params=g1.warp_latent.parameters(), lr=o.lr_pose))
    sched = torch.optim.lr_scheduler.ExponentialLR(optim, gamma=(o.lr_end / o.lr) ** (1. / opt.max_iter))
    sched_pose = torch.optim.lr_scheduler.ExponentialLR(optim_pose, gamma=(o.lr_pose_end / o.lr_pose) ** (1. / opt.max_iter))
    losses1, latent_steps = [], []
    for it, d in enumerate(ds):
        optim.zero_grad(); optim_pose.zero_grad()
        pg0 = optim_pose.param_groups[0]
        pg0["lr_orig"] = pg0["lr"]
        pg0["lr"] *= min(1, it / o.warmup_pose)
        with eng.feed_draws(ray_idx=d[0], u=d[1]):
            loss = eng.train_step(opt, g1, cfgmod.AttrDict(var1), it)
        before = g1.warp_latent.weight.detach().clone()
        optim.step(); optim_pose.step()
        latent_steps.append((g1.warp_latent.weight.detach() - before).abs().max().item())
        pg0["lr"] = pg0["lr_orig"]
        sched_pose.step(); sched.step()
        g1.nerf.progress.data.fill_((it + 1) / opt.max_iter)
        losses1.append(float(loss.all.detach()))
    assert latent_steps[0] > 0          # the latent codes move in the very first step: no warm-up on param_groups[1]

    opt, g2, var2 = make(eng)
    g2.nerf.progress.data.fill_(0.3)
    groups = eng.reference_optimizer_groups(opt, g2)
    assert groups[1]["warmup_params"] == len(list(g2.warp_mlp.parameters())) and groups[1]["torch_groups"] == [36, 1]
    fa = eng.FlatAdam(groups, progress=[g2.nerf.progress.data], max_iter=opt.max_iter)
    w0 = g2.warp_mlp.lin0_a_1.weight.detach().clone()
    l0 = g2.warp_latent.weight.detach().clone()
    losses2 = [flat_step(eng, opt, g2, var2, fa, ds[0], 0)]
    assert torch.equal(g2.warp_mlp.lin0_a_1.weight.detach(), w0)            # warm-up factor 0 in the first step ...
    assert (g2.warp_latent.weight.detach() - l0).abs().max().item() > 0     # ... for the warp network only
    # NB progress is written by the kernel: the first step ran at 0.3 in both runs, later ones at it / max_iter
    for it, d in enumerate(ds[1:], start=1):
        losses2.append(flat_step(eng, opt, g2, var2, fa, d, it))
    torch.testing.assert_close(torch.tensor(losses2[:2]), torch.tensor(losses1[:2]), rtol=1e-3, atol=1e-6)
    assert abs(float(g2.nerf.progress) - 4 / opt.max_iter) < 1e-7
    _, g0, _ = make(eng)
    updates_close(g1, g2, g0)
    # decayed learning rates as the reference's schedulers hold them
    o1, s1 = fa.to_torch(0)
    assert abs(o1.param_groups[0]["lr"] - optim.param_groups[0]["lr"]) <= 1e-12 + 1e-9 * o.lr
    o2, s2 = fa.to_torch(1)
    assert len(o2.param_groups) == 2 and abs(o2.param_groups[1]["lr"] - optim_pose.param_groups[1]["lr"]) <= 1e-15 + 1e-9 * o.lr_pose
    assert s1.last_epoch == 4 and s2.last_epoch == 4


def test_flat_adam_checkpoint_resume_equals_uninterrupted_run(eng):
    """save after two steps (graph.state_dict + FlatAdam.state_dict), restore into freshly built objects, two
    more steps == four uninterrupted steps: parameters, both moments, step counters, learning rate and ``progress``."""
    ds = draws(4, seed=5)
    opt, ga, var = make(eng)
    fa = eng.FlatAdam(eng.reference_optimizer_groups(opt, ga), progress=[ga.nerf.progress.data], max_iter=opt.max_iter)
    for it, d in enumerate(ds):
        flat_step(eng, opt, ga, var, fa, d, it)

    opt, gb, var = make(eng)
    fb = eng.FlatAdam(eng.reference_optimizer_groups(opt, gb), progress=[gb.nerf.progress.data], max_iter=opt.max_iter)
    for it, d in enumerate(ds[:2]):
        flat_step(eng, opt, gb, var, fb, d, it)
    ckpt = copy.deepcopy(dict(graph={k: v.cpu() for k, v in gb.state_dict().items()},
                              optim={k: (v.cpu() if torch.is_tensor(v) else v) for k, v in fb.state_dict().items()}))
    assert float(ckpt["graph"]["nerf.progress"]) == pytest.approx(2 / opt.max_iter)

    opt, gc, var = make(eng)                                    # a new process would start here
    gc.load_state_dict({k: v.to(DEV) for k, v in ckpt["graph"].items()})
    fc = eng.FlatAdam(eng.reference_optimizer_groups(opt, gc), progress=[gc.nerf.progress.data], max_iter=opt.max_iter)
    fc.load_state_dict({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in ckpt["optim"].items()})
    assert fc.state[:, 0].tolist() == [2.0, 2.0]
    assert float(gc.nerf.progress) == pytest.approx(2 / opt.max_iter)       # not reset by the restored optimiser
    torch.testing.assert_close(fc.exp_avg, fb.exp_avg); torch.testing.assert_close(fc.exp_avg_sq, fb.exp_avg_sq)
    for it, d in enumerate(ds[2:], start=2):
        flat_step(eng, opt, gc, var, fc, d, it)
    assert fc.state[:, 0].tolist() == [4.0, 4.0] and float(gc.nerf.progress) == pytest.approx(4 / opt.max_iter)
    _, g0, _ = make(eng)
    updates_close(ga, gc, g0)
    assert fc.to_torch(0)[0].param_groups[0]["lr"] == pytest.approx(fa.to_torch(0)[0].param_groups[0]["lr"], rel=1e-12)


def test_flat_adam_interchange_with_torch_state_dicts(eng):
    """The optimiser state in the reference's own checkpoint format (``optim.state_dict()`` / ``sched.state_dict()`` as
    util.save_checkpoint stores them): FlatAdam -> torch objects -> state_dicts -> a fresh FlatAdam continues identically,
    and the torch optimiser continues to the same parameters."""
    ds = draws(3, seed=6)
    opt, ga, var = make(eng, warmup=None)
    fa = eng.FlatAdam(eng.reference_optimizer_groups(opt, ga))
    for it, d in enumerate(ds[:2]):
        flat_step(eng, opt, ga, var, fa, d, it)
    torch_objs = [fa.to_torch(gi) for gi in range(2)]
    sds = [(o.state_dict(), s.state_dict()) for o, s in torch_objs]
    graph_sd = {k: v.clone() for k, v in ga.state_dict().items()}
    assert sds[0][0]["state"][0]["step"].item() == 2 and sds[0][1]["last_epoch"] == 2

    opt, gb, var = make(eng, warmup=None)
    gb.load_state_dict(graph_sd)
    fb = eng.FlatAdam(eng.reference_optimizer_groups(opt, gb))
    for gi in range(2):
        fb.load_torch(gi, *sds[gi])
    torch.testing.assert_close(fb.exp_avg, fa.exp_avg); torch.testing.assert_close(fb.exp_avg_sq, fa.exp_avg_sq)
    assert fb.state[:, 0].tolist() == [2.0, 2.0] and fb.groups[0]["gamma"] == pytest.approx(fa.groups[0]["gamma"])
    flat_step(eng, opt, gb, var, fb, ds[2], 2)

    # the torch optimisers (they hold ga's parameter tensors) take the same third step
    with eng.feed_draws(ray_idx=ds[2][0], u=ds[2][1]):
        eng.train_step(opt, ga, cfgmod.AttrDict(var), 2, bucket=fa)
    for o, s in torch_objs:
        o.step(); s.step()
    _, g0, _ = make(eng, warmup=None)
    g0.load_state_dict(graph_sd)
    updates_close(ga, gb, g0)


def test_gradients_through_autograd_by_default_and_in_place_on_opt_in(eng):
    """without the engine's opt-in the MLP / warp parameter gradients travel through autograd
    (``torch.autograd.grad`` sees them); with ``engine.use_flat_gradients`` the kernels accumulate into ``.grad``.  A
    network evaluated twice in one step reports ``_grads_ready`` once, after its LAST backward node."""
    opt, graph, var = make(eng)
    d = draws(1)[0]
    with eng.feed_draws(ray_idx=d[0], u=d[1]):
        v = graph.forward(opt, cfgmod.AttrDict(var), mode="train", iter=5000)
    loss = eng.summarize_loss(opt, graph.compute_loss(opt, v, mode="train")).all
    w, wn = graph.nerf.mlp_feat[3].weight, graph.warp_mlp.lin1_b_1.weight
    gw, gn = torch.autograd.grad(loss, [w, wn])
    assert gw is not None and gn is not None and w.grad is None and wn.grad is None
    assert gw.abs().max().item() > 0 and gn.abs().max().item() > 0

    opt, g2, var2 = make(eng)
    bucket = eng.GradBucket(g2)
    with eng.feed_draws(ray_idx=d[0], u=d[1]):
        eng.train_step(opt, g2, cfgmod.AttrDict(var2), 5000, bucket=bucket)
    assert g2.nerf.accumulate_grads_in_place and g2.warp_mlp.accumulate_grads_in_place
    torch.testing.assert_close(g2.nerf.mlp_feat[3].weight.grad, gw, rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(g2.warp_mlp.lin1_b_1.weight.grad, gn, rtol=2e-2, atol=1e-6)

    # two evaluations of one network in a step
    bucket.zero()
    fired = []
    g2.nerf._grads_ready = lambda m: fired.append(m._pending_backward)
    gen = torch.Generator().manual_seed(3)
    center = (torch.randn(B, P, 3, generator=gen) * 0.1).to(DEV).requires_grad_(True)
    ray = torch.nn.functional.normalize(torch.randn(B, P, 3, generator=gen), dim=-1).to(DEV)
    d1, d2 = (torch.rand(B, P, N, 1, generator=gen) + 0.5).to(DEV), (torch.rand(B, P, N, 1, generator=gen) + 0.5).to(DEV)
    rgb1, _ = g2.nerf.forward_samples(opt, center, ray, d1, mode="train")
    rgb2, _ = g2.nerf.forward_samples(opt, center, ray, d2, mode="train")
    assert g2.nerf._pending_backward == 2
    rgb1.sum().backward()
    assert fired == [] and g2.nerf._pending_backward == 1
    g_first = g2.nerf.mlp_feat[0].weight.grad.clone()
    rgb2.sum().backward()
    assert fired == [0]
    g2.nerf._grads_ready = None
    assert (g2.nerf.mlp_feat[0].weight.grad - g_first).abs().max().item() > 0        # the second node accumulated on top


@pytest.mark.parametrize("model", ["barf_inn_llff", "barf_inn_dtu"])
def test_evaluate_view_defaults_for_inn_models(eng, model):
    """``evaluate_view(opt, graph, var)`` with default arguments on the INN models, whose YAMLs set
    optim.test_photo: the test-time refinement runs (the reference gates DTU with ``'barf' in opt.model``,
    nerf_inn_dtu.py:217) and the eval pose composes its result; skipping it explicitly composes the identity."""
    hw = [20, 28]
    kw = dict(barf_c2f=[0.1, 0.5], device=DEV, data=dict(image_size=hw), optim=dict(test_iter=2),
              nerf=dict(rand_rays=140, sample_intvs=16), arch=dict(mlp_precision="fp32"))
    if model == "barf_inn_dtu":
        kw["nerf"]["depth"] = dict(range=[1.2, 5.2])
    opt = cfgmod.builtin_options(model, **kw)
    assert opt.optim.test_photo
    var = eng.synthetic_var(opt, 1, 5, dtu=(model == "barf_inn_dtu"))
    graph = eng.build_graph(opt, 1, initial_poses_w2c=var.pose.clone()) if model == "barf_inn_dtu" else eng.build_graph(opt, 1)
    load_nerf(graph.nerf, syn.nerf_params(3))
    sim3 = cfgmod.AttrDict(t0=torch.zeros(1, 3, device=DEV), t1=torch.zeros(1, 3, device=DEV), s0=1.0, s1=1.0,
                           R=torch.eye(3, device=DEV))
    graph.sim3 = sim3
    if hasattr(graph, "pose_net"):
        graph.pose_net.sim3_est_to_gt_c2w = cfgmod.AttrDict(type="traj_align", s=1.0, R=torch.eye(3, device=DEV),
                                                            t=torch.zeros(3, 1, device=DEV))
    res = eng.evaluate_view(opt, graph, cfgmod.AttrDict(var))
    assert res.var.se3_refine_test.shape == (1, 6) and torch.isfinite(res.psnr).all() and torch.isfinite(res.ssim).all()
    res2 = eng.evaluate_view(opt, graph, cfgmod.AttrDict(var), test_optim=False)
    assert torch.equal(res2.var.pose_refine_test, torch.eye(3, 4, device=DEV)[None]) and torch.isfinite(res2.psnr).all()


def test_backward_overlap_matches_sequential_backward(eng):
    """engine.train_step with a flat bucket runs the MLP weight-gradient pass on a side stream under the pose / warp
    backward (functional.BackwardOverlap): same gradients as the single-stream backward (the dW slices are partitioned over
    fewer CTAs, so only the fp32 summation order differs), streams joined on return, also when captured in a CUDA graph."""
    d = draws(1, seed=12)[0]
    grads = {}
    for overlap in (False, True):
        opt, g, var = make(eng, precision="bf16")
        bucket = eng.GradBucket(g)
        with eng.feed_draws(ray_idx=d[0], u=d[1]):
            eng.train_step(opt, g, cfgmod.AttrDict(var), 5000, bucket=bucket, overlap_dw=overlap)
        torch.cuda.synchronize()
        grads[overlap] = bucket.flat.clone()
        if overlap:
            assert g.nerf.overlap_weight_gradients and not g._backward_overlap.used      # joined
            s_ridx, s_u = d[0].clone(), d[1].clone()

            def body():
                with eng.feed_draws(ray_idx=s_ridx, u=s_u):
                    eng.train_step(opt, g, cfgmod.AttrDict(var), 5000, bucket=bucket)
                return bucket.flat
            captured = eng.CapturedStep(body, warmup=1)
            captured()
            torch.cuda.synchronize()
            grads["graph"] = bucket.flat.clone()
    for k in (True, "graph"):
        rel = ((grads[k].double() - grads[False].double()).norm() / grads[False].double().norm()).item()
        assert rel < 1e-4, (k, rel)
        assert (grads[k] != 0).sum() == (grads[False] != 0).sum()


def test_train_step_with_optimizer_equals_step_then_update(eng):
    """engine.train_step(optimizer=FlatAdam) updates the pose / warp groups before the side stream is joined and the NeRF group
    after it.  The groups are independent: on the SAME gradients, FlatAdam.step(groups=...) group by group in that order is
    bit-identical to one FlatAdam.step(); and the fused call takes exactly one step in every group."""
    d = draws(1, seed=31)[0]
    opt, g, var = make(eng, precision="bf16")
    fa = eng.FlatAdam(eng.reference_optimizer_groups(opt, g))
    with eng.feed_draws(ray_idx=d[0], u=d[1]):
        eng.train_step(opt, g, cfgmod.AttrDict(var), 0, bucket=fa)
    torch.cuda.synchronize()
    grads, p0 = fa.flat.clone(), fa.flat_params.clone()
    fa.step()
    torch.cuda.synchronize()
    whole = (fa.flat_params.clone(), fa.exp_avg.clone(), fa.exp_avg_sq.clone(), fa.state.clone())
    # same start, same gradients, group by group in the order train_step(optimizer=) uses
    fa.flat_params.copy_(p0); fa.exp_avg.zero_(); fa.exp_avg_sq.zero_(); fa.state.zero_(); fa.flat.copy_(grads)
    fa.step(groups=range(1, len(fa.groups)))
    fa.step(groups=[0])
    torch.cuda.synchronize()
    for a, b in zip(whole, (fa.flat_params, fa.exp_avg, fa.exp_avg_sq, fa.state)):
        assert torch.equal(a, b)
    # the fused call: one more step in every group, parameters move
    before = fa.flat_params.clone()
    with eng.feed_draws(ray_idx=d[0], u=d[1]):
        eng.train_step(opt, g, cfgmod.AttrDict(var), 1, bucket=fa, optimizer=fa)
    torch.cuda.synchronize()
    assert torch.equal(fa.state[:, 0], torch.full_like(fa.state[:, 0], 2.0))
    for gr in fa.groups:
        seg = slice(gr["offset"], gr["offset"] + gr["n"])
        assert not torch.equal(fa.flat_params[seg], before[seg])


def test_warp_network_prepack_on_a_side_stream_is_the_same_forward(eng):
    """DeformNetwork.prepack (the weight pack launched early on a side stream, barf_inn_llff._prefetch_pose): the following
    forward picks it up -- same output bits and same gradients as the forward that packs for itself; a prepack for another
    code tensor is ignored."""
    opt, g, var = make(eng)
    net, code = g.warp_mlp, g.warp_latent.weight
    gen = torch.Generator().manual_seed(5)
    pts = (torch.randn(B, 40, 1, 3, generator=gen) * 0.5).to(DEV)
    w = torch.randn(B, 40, 1, 3, generator=gen).to(DEV)
    outs, grads = [], []
    side = torch.cuda.Stream()
    for mode in ("plain", "prepacked", "stale"):
        for p in list(net.parameters()) + [code]:
            p.grad = None
        if mode == "prepacked":
            side.wait_stream(torch.cuda.current_stream())
            net.prepack(code, side)
        elif mode == "stale":
            net.prepack(code.detach().clone(), side)          # another tensor: must not be used
        out = net.forward(code, pts, alpha_ratio=0.4)
        (out * w).sum().backward()
        torch.cuda.synchronize()
        assert getattr(net, "_prepacked", None) is None
        outs.append(out.detach().clone())
        grads.append(torch.cat([p.grad.reshape(-1) for p in net.parameters()] + [code.grad.reshape(-1)]).clone())
    for o, gr in zip(outs[1:], grads[1:]):
        assert torch.equal(o, outs[0])
        rel = ((gr.double() - grads[0].double()).norm() / grads[0].double().norm()).item()
        assert rel < 1e-5, rel


def test_captured_test_time_pose_refinement(eng):
    """the test-time photometric pose refinement (reference model/barf.py:153-169) as ONE captured CUDA graph
    replayed per iteration on engine.FlatAdam (device pixel draws, no host sync): from a perturbed test pose the loss of the
    fitted view falls and the refinement moves towards the perturbation's inverse; the networks are not touched (frozen:
    the MLP backward skips its weight-gradient pass) and the eager torch.optim loop, fed the same draws, takes the same
    first step."""
    from neural_invertible_warp_b200 import camera, _lib
    Hs, Ws = 24, 32
    opt = cfgmod.builtin_options("barf_llff", model="barf", barf_c2f=[0.1, 0.5], device=DEV, data=dict(image_size=[Hs, Ws]),
                                 nerf=dict(rand_rays=256, sample_intvs=32), optim=dict(test_iter=40, lr_pose=3e-3),
                                 arch=dict(mlp_precision="bf16"))
    graph = eng.build_graph(opt, 1)
    load_nerf(graph.nerf, syn.nerf_params(3))
    # early in the coarse-to-fine schedule (2.5 of 10 bands): the random-init field is smooth, so the photometric loss of a
    # shifted view is well above the stratified-sampling noise and has a usable slope (with all bands on it is not)
    graph.nerf.progress.data.fill_(0.2)
    graph.sim3 = cfgmod.AttrDict(t0=torch.zeros(1, 3, device=DEV), t1=torch.zeros(1, 3, device=DEV), s0=1.0, s1=1.0,
                                 R=torch.eye(3, device=DEV))
    var = eng.synthetic_var(opt, 1, 5)
    # ground truth image = the network's own render from the true pose (un-stratified would need another opt: use eval render)
    with torch.no_grad():
        var.pose_refine_test = torch.eye(3, 4, device=DEV)[None]
        opt.optim.test_photo = True
        full = graph.forward(opt, cfgmod.AttrDict(var), mode="eval")
        var.image = full.rgb.view(1, Hs, Ws, 3).permute(0, 3, 1, 2).contiguous()
    # perturb the test pose: the refinement has to undo it
    delta = torch.tensor([[0.0, 0.0, 0.0, 0.06, -0.045, 0.0]], device=DEV)
    var.pose = camera.pose.compose([camera.lie.se3_to_SE3(delta), var.pose])
    before = {k: v.detach().clone() for k, v in graph.named_parameters()}

    def view_loss(v):
        with torch.no_grad():
            out = graph.forward(opt, cfgmod.AttrDict(v), mode="eval")
            return float(((out.rgb.view(1, Hs, Ws, 3).permute(0, 3, 1, 2) - v.image) ** 2).mean())
    var.pose_refine_test = torch.eye(3, 4, device=DEV)[None]
    l0 = view_loss(var)
    n0 = _lib.launch_count()
    out = eng.test_time_photometric_optim(opt, graph, cfgmod.AttrDict(var))
    torch.cuda.synchronize()
    assert float(out.test_optim_steps) == opt.optim.test_iter
    assert _lib.launch_count() - n0 < 3 * 40                      # two eager warm-up iterations + capture: the replays launch nothing from the host
    l1 = view_loss(out)
    print("captured test-time refinement: view loss %.3e -> %.3e, se3 %s" % (l0, l1, out.se3_refine_test.detach().cpu().numpy().round(4)))
    assert l1 < 0.8 * l0
    for k, v in graph.named_parameters():
        assert torch.equal(v, before[k]) and v.requires_grad, k


@pytest.mark.parametrize("shard,dataset", [(None, "llff"), ((40, 200), "llff"), (None, "blender")])
def test_one_launch_warped_ray_generation_equals_the_three_launch_path(eng, shard, dataset):
    """csrc/nvp.cu niw_nvp_rays_fwd (pixels -> un-warped grid -> NVP warp -> ray / centre in one launch;
    model/barf_inn_llff.py:325-364) against raygen_unwarped -> warp -> rays_from_warp: same outputs, same gradients to the
    warp network and the latent codes, same un-warped points left in ``var``; also as a ray shard at its position in the
    global point list, and with an initial pose (blender branch, :309-323)."""
    from neural_invertible_warp_b200 import functional as F
    opt, g, var = make(eng)
    var = cfgmod.AttrDict(var)
    if dataset == "blender":
        opt.data.dataset, opt.camera.noise_type = "blender", None
        gen = torch.Generator().manual_seed(3)
        rot = torch.linalg.qr(torch.randn(B, 3, 3, generator=gen))[0]
        var.pose = torch.cat([rot, torch.randn(B, 3, 1, generator=gen)], dim=-1).to(DEV)
    Pn = 72
    gen = torch.Generator().manual_seed(9)
    var.ray_idx = torch.randperm(H * W, generator=gen)[:Pn].to(DEV)
    wr, wc, wg = (torch.randn(B, Pn, 3, generator=gen).to(DEV) for _ in range(3))
    params = list(g.warp_mlp.parameters()) + [g.warp_latent.weight]
    res = {}
    saved = F.fused_warped_rays, F.ray_shard
    try:
        F.ray_shard = shard
        for fused in (False, True):
            F.fused_warped_rays = fused
            for p in params:
                p.grad = None
            launches0 = _lib_launches()
            ray, center, grid, alpha = g.get_pose(opt, var, mode="train", iter=20)
            n_launch = _lib_launches() - launches0
            ((ray * wr).sum() + (center * wc).sum() + (grid * wg).sum()).backward()
            torch.cuda.synchronize()
            res[fused] = dict(out=[t.detach().clone() for t in (ray, center, grid, var.grid_cam, var.center_cam)], alpha=alpha,
                              grad=torch.cat([p.grad.reshape(-1) for p in params]).clone(), launches=n_launch)
    finally:
        F.fused_warped_rays, F.ray_shard = saved
    assert res[True]["alpha"] == res[False]["alpha"]
    for a, b in zip(res[True]["out"], res[False]["out"]):
        assert a.shape == b.shape
        torch.testing.assert_close(a, b, rtol=0, atol=1e-6)   # (one ulp: the two kernels contract the pixel -> world arithmetic differently)
    rel = ((res[True]["grad"].double() - res[False]["grad"].double()).norm() / res[False]["grad"].double().norm()).item()
    assert rel < 1e-5, rel
    assert res[True]["launches"] == res[False]["launches"] - 2, (res[True]["launches"], res[False]["launches"])


def _lib_launches():
    import ctypes
    from neural_invertible_warp_b200 import _lib
    return int(_lib.load().niw_launch_count())
