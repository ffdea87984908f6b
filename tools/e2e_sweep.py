"""Sweep the number of batch slices of HostFwBw at the bench workload (PCIe-bound end-to-end path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xlstm_yolo_clean_b200 as pkg
from oracle import mlstm_oracle as O

B, NH, S, D = 32, 4, 1600, 64
inp = O.make_inputs(B, NH, S, D, D, seed=0, dtype=torch.float32)
host = {k: v.to(torch.bfloat16).pin_memory() for k, v in inp.items()}
out = pkg.HostFwBw.alloc_host(B, NH, S, D, D)
for ns, tp, ln in ((5, True, 1), (1, False, 2), (2, False, 2), (5, True, 2), (1, False, 3), (1, False, 4), (2, False, 3)):
    pipe = pkg.HostFwBw(B, NH, S, D, D, n_slices=ns, taper=tp, lanes=ln)
    print(f"lanes={ln}", [sl.stop - sl.start for sl in pipe.slices], end=' ')
    for _ in range(4):
        pipe.run(host, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(12):
        pipe.run(host, out)
    pipe.flush()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    print(f"n_slices={ns:2d}  {ms:.3f} ms/step  {23.49e9 / (ms * 1e-3) / 1e12:.2f} TFLOP/s  ({(pipe.h2d_bytes + pipe.d2h_bytes) / ms / 1e6:.1f} GB/s both ways)")
