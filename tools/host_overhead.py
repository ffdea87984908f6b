"""Host-side cost of one eager call of the public API (the YOLO-ViL training step is launch-bound on the host at
32 img/GPU, so microseconds per call matter): wall time per call with the GPU kept far from the bottleneck
(tiny shape), and a cProfile breakdown."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import xlstm_yolo_clean_b200 as pkg
from oracle import mlstm_oracle as O

dev = torch.device("cuda:0")
inp = O.make_inputs(1, 2, 128, 64, 64, seed=0, dtype=torch.float32)
t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
h0 = torch.randn(1, 8, 128, 64, device=dev).to(torch.bfloat16).requires_grad_(True)
x0 = torch.randn(1, 128, 512, device=dev).to(torch.float16)
w0 = torch.randn(512, device=dev).requires_grad_(True)


def fw_only():
    with torch.no_grad():
        pkg.mlstm_chunkwise__b200(**leaves, chunk_size=64)


def fwbw():
    h = pkg.mlstm_chunkwise__b200(**leaves, chunk_size=64)
    h.backward(t["dh"])


def raw_fw():
    pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])


def cell():
    y = pkg.cell_out(h0, w0, None, w0, x0, out_dtype=torch.float16)
    y.backward(x0)


for name, fn in (("raw C-ABI forward wrapper", raw_fw), ("autograd forward (no_grad)", fw_only), ("autograd fwd+bwd", fwbw),
                 ("cell_out fwd+bwd", cell)):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:32s} {1e6 * (t1 - t0) / n:7.1f} us per call (host)")
pr = cProfile.Profile()
pr.enable()
for _ in range(1000):
    fwbw()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
