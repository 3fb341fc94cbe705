"""torch.profiler trace of ONE eager C2 train step: which ATen ops launch which kernels (finds stray launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from neural_invertible_warp_b200 import config as cfgmod, engine, synthetic as syn

dev = "cuda:0"
IMAGES = 16
opt = cfgmod.builtin_options("barf_inn_llff", barf_c2f=[0.1, 0.5], device=dev, nerf=dict(rand_rays=1024, sample_intvs=128),
                             arch=dict(mlp_precision="bf16"))
graph = engine.build_graph(opt, IMAGES)
graph.warp_mlp.load_state_dict({k: v.to(dev) for k, v in syn.nvp_params(1).items()})
graph.warp_latent.weight.data = syn.latent_codes(2, IMAGES).to(dev)
graph.nerf.progress.data.fill_(0.3)
var_dev = engine.synthetic_var(opt, IMAGES, seed=3)
bucket = engine.GradBucket(graph)
optim = torch.optim.Adam([dict(params=graph.nerf.parameters(), lr=1e-3)], fused=True, capturable=True)
optim_pose = torch.optim.Adam([dict(params=list(graph.warp_mlp.parameters()) + list(graph.warp_latent.parameters()), lr=5e-4)],
                              fused=True, capturable=True)

def step():
    v = cfgmod.AttrDict(var_dev)
    loss = engine.train_step(opt, graph, v, 5000, bucket=bucket)
    optim.step(); optim_pose.step()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events()], key=lambda e: e.time_range.start)
for e in evs:
    if e.device_type == torch.autograd.DeviceType.CUDA or e.name.startswith(("aten::", "cuda")) or "Backward" in e.name:
        depth = 0
        p = e.cpu_parent
        while p is not None:
            depth += 1; p = p.cpu_parent
        kind = "K" if e.device_type == torch.autograd.DeviceType.CUDA else "c"
        if kind == "K" or depth <= 1:
            print("%s %s%s" % (kind, "  " * depth, e.name[:100]))
