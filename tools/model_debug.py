#!/usr/bin/env python
"""Debug: run the reference model (staged under baseline/_ref) forward/backward with the B200 backend and
cross-check EVERY kernel call on the spot: tensor-core family vs a second run of itself (determinism) vs the
exact fp32 family (independent implementation).  Dumps the inputs of the first bad call to gpurun_out/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import model_bench as MB  # noqa: E402

MB._import_reference()
import xlstm_yolo_clean_b200 as pkg  # noqa: E402
from xlstm_yolo_clean_b200 import _cabi, backend  # noqa: E402

yaml_name = sys.argv[1] if len(sys.argv) > 1 else "640-base256.yaml"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
siging = (sys.argv[4] == "siging") if len(sys.argv) > 4 else False

orig_fw, orig_bw = backend.mlstm_chunkwise_fw, backend.mlstm_chunkwise_bw
state = {"n": 0, "dumped": False}


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def describe(t):
    return f"{tuple(t.shape)} stride {t.stride()} {t.dtype} finite={bool(torch.isfinite(t).all())} absmax={float(t.float().abs().nan_to_num(0, 0, 0).max()):.4g}"


def fw(q, k, v, i, f, *a, **kw):
    state["n"] += 1
    out = orig_fw(q, k, v, i, f, *a, **kw)
    out2 = orig_fw(q, k, v, i, f, *a, **kw)
    kw_e = dict(kw, impl=_cabi.IMPL_EXACT, save_states=False)
    out_e = orig_fw(q.float(), k.float(), v.float(), i.float(), f.float(), *a, **kw_e)
    d_self = rel(out[0], out2[0])
    d_exact = rel(out[0], out_e[0])
    bad = d_self != 0.0 or not d_exact < 3e-2
    print(f"fw call {state['n']:3d} S={q.shape[2]:5d} self-diff {d_self:.3g} vs-exact {d_exact:.3g} in-finite="
          f"{all(bool(torch.isfinite(t).all()) for t in (q, k, v, i, f))} {'BAD' if bad else ''}", flush=True)
    if bad and not state["dumped"]:
        state["dumped"] = True
        for n, t in (("q", q), ("k", k), ("v", v), ("i", i), ("f", f), ("h", out[0]), ("h_exact", out_e[0])):
            print("   ", n, describe(t))
        rows = (out[0].float() - out_e[0].float()).abs().amax(-1)  # (B, NH, S)
        idx = (rows > 0.03 * out_e[0].float().abs().max()).nonzero()
        print("    bad (b, h, s) rows:", idx[:20].tolist(), "count", idx.shape[0])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        torch.save({"q": q.cpu(), "k": k.cpu(), "v": v.cpu(), "i": i.cpu(), "f": f.cpu(), "kw": {x: y for x, y in kw.items() if not torch.is_tensor(y)}},
                   os.path.join(ROOT, "gpurun_out", "bad_call.pt"))
    return out


def bw(q, k, v, i, f, n_out, m_out, dh, *a, **kw):
    state["n"] += 1
    out = orig_bw(q, k, v, i, f, n_out, m_out, dh, *a, **kw)
    out2 = orig_bw(q, k, v, i, f, n_out, m_out, dh, *a, **kw)
    kw_e = dict(kw, impl=_cabi.IMPL_EXACT, c_states=None)
    fin = bool(torch.isfinite(dh).all())
    msg = f"bw call {state['n']:3d} S={q.shape[2]:5d} dh-finite={fin} self-diff " + " ".join(f"{rel(x, y):.3g}" for x, y in zip(out[:5], out2[:5]))
    if fin:
        out_e = orig_bw(q.float(), k.float(), v.float(), i.float(), f.float(), n_out, m_out, dh.float(), *a, **kw_e)
        msg += " vs-exact " + " ".join(f"{rel(x, y):.3g}" for x, y in zip(out[:5], out_e[:5]))
    print(msg, flush=True)
    return out


backend.mlstm_chunkwise_fw, backend.mlstm_chunkwise_bw = fw, bw
dev = torch.device("cuda", 0)
model = MB._build_model(yaml_name, dev)
print("cells patched:", pkg.patch_model(model, siging=siging))
batch = MB._batch(B, dev, seed=0)
model.train()
scaler = torch.amp.GradScaler("cuda")
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.937, nesterov=True)
for s in range(steps):
    with torch.autocast("cuda", dtype=torch.float16):
        loss, items = model(batch)
        loss = loss.sum()
    print(f"== step {s} loss {float(loss)} ==", flush=True)
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    scaler.step(opt)
    scaler.update()
    opt.zero_grad(set_to_none=True)
    print(f"== step {s} done, scale {scaler.get_scale()} ==", flush=True)
