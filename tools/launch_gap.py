"""Marginal cost of one more fw / bw launch inside a CUDA graph (config 2): separates kernel period from launch gaps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xlstm_yolo_clean_b200 as pkg
from oracle import mlstm_oracle as O

B, NH, S, D = 32, 4, 1600, 64
sets = []
for r in range(4):
    inp = O.make_inputs(B, NH, S, D, D, seed=r, dtype=torch.float32)
    sets.append({k: v.to(torch.bfloat16).cuda() for k, v in inp.items()})
saved = [pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"]) for t in sets]
def fw(t): return pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])
def bw(t, s): return pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], s[1], s[2], t["dh"], c_states=s[4])
torch.cuda.synchronize()
def timed(fn, n):
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(n): fn(i)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(n): fn(i)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(30):
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
for name, fn in (("fw", lambda i: fw(sets[i % 4])), ("bw", lambda i: bw(sets[i % 4], saved[i % 4])),
                 ("fw+bw", lambda i: (fw(sets[i % 4]), bw(sets[i % 4], saved[i % 4])))):
    res = {n: timed(fn, n) for n in (1, 2, 4, 8, 16)}
    print(name, " ".join(f"n={n}: {t:7.1f} us" for n, t in res.items()), f"| marginal (16-8)/8 = {(res[16] - res[8]) / 8:.1f} us")
