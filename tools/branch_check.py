#!/usr/bin/env python
"""Model-level parity of the fused, flip-free branch: inside the unmodified reference model (staged under
baseline/_ref), every ViLLayer.mlstm_branch call is evaluated twice on the same input -- the reference's own
method (flips, MatrixLSTMCell with the chunkwise--b200 kernel, group_norm) and vil.mlstm_branch_b200 -- and the
outputs and input gradients are compared.  Prints one JSON line per YAML."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import model_bench as MB  # noqa: E402

MB._import_reference()
import xlstm_yolo_clean_b200 as pkg  # noqa: E402

dev = torch.device("cuda", 0)


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for yaml_name in sys.argv[1:] or ["640-base192.yaml", "640-base256.yaml", "640-base384.yaml"]:
    model = MB._build_model(yaml_name, dev)
    pkg.patch_model(model, siging=False)  # unfused: reference modules + chunkwise--b200 (exp gate, the oracle function)
    stats = []
    for mod in model.modules():
        if hasattr(mod, "mlstm_cell") and hasattr(mod, "proj_up"):
            orig = mod.mlstm_branch

            def both(self, x, _orig=orig):
                with torch.enable_grad():
                    xa = x.detach().clone().requires_grad_(True)
                    xb = x.detach().clone().requires_grad_(True)
                    ya = _orig(xa)
                    yb = pkg.mlstm_branch_b200(self, xb)
                    g = torch.randn_like(ya)
                    (ga,) = torch.autograd.grad(ya, xa, g)
                    (gb,) = torch.autograd.grad(yb, xb, g)
                if bool(torch.isfinite(ya).all()):
                    stats.append((x.shape[1], "rev" if pkg.vil._is_reverse(self) else "fwd", rel(yb, ya), rel(gb, ga)))
                return _orig(x)

            mod.mlstm_branch = types.MethodType(both, mod)
    x = MB._batch(2, dev, 0)["img"]
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        model(x)
    n_eval = len(stats)
    model.train()
    with torch.autocast("cuda", dtype=torch.float16):
        model(MB._batch(2, dev, 0))
    worst_y = max(s[2] for s in stats)
    worst_g = max(s[3] for s in stats)
    by = {}
    for S, d, ey, eg in stats:
        k = f"S={S} {d}"
        by[k] = [max(by.get(k, [0, 0])[0], ey), max(by.get(k, [0, 0])[1], eg)]
    print(json.dumps({"yaml": yaml_name, "branch_calls_checked": len(stats), "eval_calls": n_eval,
                      "max_rel_err_out": worst_y, "max_rel_err_dx": worst_g, "by_stage": by}), flush=True)
    del model
    torch.cuda.empty_cache()
