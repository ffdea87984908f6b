"""One warm-up and one measured forward + backward of the tensor-core kernels at the shapes the ncu captures under
profiles/ are taken on: config 2 (d=64), the base192 S=6400 call (d=32) and the base384 S=6400 call (d=128).
    python tools/ncu_shapes.py && ncu --set full --clock-control none --import-source on -k regex:tc_ -s 6 -c 6 -o gpurun_out/prof python tools/ncu_shapes.py
(-s 6 skips the warm-up launches: three shapes x (forward + backward))."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import xlstm_yolo_clean_b200 as pkg

SHAPES = [(32, 4, 1600, 64), (64, 12, 6400, 32), (16, 6, 6400, 128)]
ts = []
for B, NH, S, D in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(S + D)
    t = {k: (0.5 * torch.randn(B, NH, S, D, generator=g, device="cuda")).to(torch.bfloat16) for k in ("q", "k", "v", "dh")}
    t["i"] = torch.randn(B, NH, S, generator=g, device="cuda").to(torch.bfloat16)
    t["f"] = (3 + torch.randn(B, NH, S, generator=g, device="cuda")).to(torch.bfloat16)
    ts.append(t)
for rep in range(2):
    for t in ts:
        h, n, m, _, cst = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])
        pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n, m, t["dh"], c_states=cst)
    torch.cuda.synchronize()
print("ok")
