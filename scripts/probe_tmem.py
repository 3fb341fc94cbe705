"""Does the tensor pipe (tcgen05.mma accumulating into TMEM) overlap with draining TMEM (tcgen05.ld) on one SM?
One layer of the 256-wide MLP per iteration: 16 MMAs of M=128, N=256, K=16 (1.05 MFLOP each) and one 128 KB accumulator
read.  Prints cycles per iteration for the MMA stream alone, the readers alone, and both together."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_invertible_warp_b200 import _lib

lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 200
for what, name in ((1, "mma alone"), (2, "tmem reads alone"), (3, "both")):
    for _ in range(2):
        _lib.check(lib.niw_tc_probe(what, iters, ctypes.c_void_p(out.data_ptr()), st))
        torch.cuda.synchronize()
    m, r = [int(x) for x in out.tolist()]
    print("%-18s mma %8.1f clk/iter (%.0f FLOP/clk)   tmem read %8.1f clk/iter (%.1f B/clk)" % (
        name, m / iters, (16 * 2 * 128 * 256 * 16) * iters / m if m else 0, r / iters, 131072 * iters / r if r else 0))
