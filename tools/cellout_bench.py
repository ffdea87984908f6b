"""Bandwidth of the fused cell-output kernels (mlstm_b200_cellout_fw / _bw) at the YOLO-ViL call shapes,
next to the torch composition the reference runs (group_norm + transposes + skip add, vision_lstm2.py:749-751,
928-944, 306).  Algorithmic bytes: fw reads h, x and writes y; bw reads dy, h, x and writes dh, dx."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import xlstm_yolo_clean_b200 as pkg

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(fn, reps=20):
    """Device time of fn's kernels alone: captured once, replayed with an L2 flush in between."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            keep = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    del keep
    return ts[len(ts) // 2]


class _Ctx:
    needs_input_grad = (True,) * 7

    def save_for_backward(self, *a):
        self.saved_tensors = a


CO = pkg.vil._CellOut


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def ref(h, w, b, sk, x):
    B, NH, S, D = h.shape
    hh = h.to(x.dtype)
    with torch.autocast("cuda", dtype=torch.float16):
        g = F.group_norm(hh.transpose(1, 2).reshape(B * S, NH * D), NH, w, b, 1e-6).view(B, S, NH, D).transpose(1, 2)
        return g.transpose(1, 2).reshape(B, S, -1) + sk * x


for (B, NH, S, D) in ((32, 8, 6400, 64), (32, 8, 1600, 64), (64, 12, 6400, 32), (16, 6, 6400, 128)):
    H = NH * D
    h = torch.randn(B, NH, S, D, device=dev).to(torch.bfloat16).requires_grad_(True)
    x = torch.randn(B, S, H, device=dev).to(torch.float16).requires_grad_(True)
    w, b, sk = (torch.randn(H, device=dev).requires_grad_(True) for _ in range(3))
    dy = torch.randn(B, S, H, device=dev).to(torch.float16)
    n = B * S * H
    fw_bytes, bw_bytes = n * 6, n * 10
    g_fw = graph_time(lambda: pkg.cell_out(h.detach(), w.detach(), b.detach(), sk.detach(), x.detach(), out_dtype=torch.float16))
    t_fw = timeit(lambda: pkg.cell_out(h.detach(), w.detach(), b.detach(), sk.detach(), x.detach(), out_dtype=torch.float16))
    y = pkg.cell_out(h, w, b, sk, x, out_dtype=torch.float16)
    t_bw = timeit(lambda: torch.autograd.grad(y, (h, w, b, sk, x), dy, retain_graph=True))
    ctx = _Ctx()
    CO.forward(ctx, h.detach(), w.detach(), b.detach(), sk.detach(), x.detach(), 1e-6, torch.float16)
    g_bw = graph_time(lambda: CO.backward(ctx, dy))  # the C-ABI backward alone (no autograd engine)
    t_rfw = timeit(lambda: ref(h.detach(), w.detach(), b.detach(), sk.detach(), x.detach()), 5)
    yr = ref(h, w, b, sk, x)
    t_rbw = timeit(lambda: torch.autograd.grad(yr, (h, w, b, sk, x), dy.float() if yr.dtype == torch.float32 else dy, retain_graph=True), 5)
    print(json.dumps({"shape": [B, NH, S, D], "fw_ms": t_fw, "fw_gbs": fw_bytes / t_fw / 1e6, "fw_frac": fw_bytes / t_fw / 1e6 / PEAK,
                      "bw_ms": t_bw, "bw_gbs": bw_bytes / t_bw / 1e6, "bw_frac": bw_bytes / t_bw / 1e6 / PEAK,
                      "graph_fw_ms": g_fw, "graph_fw_frac": fw_bytes / g_fw / 1e6 / PEAK, "graph_bw_ms": g_bw, "graph_bw_frac": bw_bytes / g_bw / 1e6 / PEAK,
                      "torch_fw_ms": t_rfw, "torch_bw_ms": t_rbw, "timing": "eager autograd call incl. host launch, L2 flushed"}), flush=True)
    del h, x, y, yr
