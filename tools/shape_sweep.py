"""Kernel time of the mLSTM calls the YAML models make (SURVEY.md §3.1 shapes), tensor path vs exact path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xlstm_yolo_clean_b200 as pkg
from oracle import mlstm_oracle as O

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

cfgs = [("base256 (cfg4/GPU)", 32, 8, 64), ("base384 (cfg5/GPU)", 16, 6, 128), ("base192 (cfg3)", 64, 12, 32)]
for name, B, NH, D in cfgs:
    for S in (6400, 1600, 448, 128):
        g = torch.Generator().manual_seed(0)
        t = {k: (0.3 * torch.randn(B, NH, S, D, generator=g)).to(torch.bfloat16).cuda() for k in ("q", "k", "v", "dh")}
        t["i"] = torch.full((B, NH, S), -8.73).to(torch.bfloat16).cuda()
        t["f"] = (3.0 + 3.0 * torch.rand(B, NH, S, generator=g)).to(torch.bfloat16).cuda()
        res = {}
        for impl in ("auto", "exact"):
            if impl == "exact" and S == 6400 and D != 32: continue
            pkg.set_default_impl(impl)
            saved = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])
            fw = timeit(lambda: pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"]))
            bw = timeit(lambda: pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], saved[1], saved[2], t["dh"], c_states=saved[4]))
            res[impl] = (fw, bw)
        pkg.set_default_impl("auto")
        tok = B * NH * S
        fwb = tok * (4 * D * 2 + 12); bwb = tok * (7 * D * 2 + 24)
        a = res["auto"]
        line = f"{name:20s} S={S:5d} d={D:3d} heads={B*NH:4d}  fw {a[0]*1e3:8.1f} us ({fwb/a[0]/1e6:6.0f} GB/s)  bw {a[1]*1e3:8.1f} us ({bwb/a[1]/1e6:6.0f} GB/s)"
        if "exact" in res: line += f"   exact fw {res['exact'][0]*1e3:9.1f} us bw {res['exact'][1]*1e3:9.1f} us"
        print(line)
