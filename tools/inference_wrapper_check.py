#!/usr/bin/env python
"""the B200 kernels behind the reference's inference wrapper
(wrap_chunkwise__arbitrary_sequence_length, mlstm_kernels/torch/kernel_wrappers.py:12-201), which splits an
arbitrary sequence length into chunk-64/32/16 kernel calls that hand (C, n) states to each other plus a
step-kernel remainder.  The wrapper unpacks 2-tuple states, i.e. it only fits the sigmoid-input-gate kernels
(SURVEY appendix B), so this drives chunkwise--b200_siging through mLSTMBackend(mode="inference") and compares
with the float64 oracle evaluated on the whole sequence at once.  Needs baseline/_ref (tools/stage_reference.sh)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import model_bench as MB  # noqa: E402

MB._import_reference()
from mlstm_kernels.torch.backend_module import mLSTMBackend, mLSTMBackendConfig  # noqa: E402

import xlstm_yolo_clean_b200 as pkg  # noqa: E402
from oracle import mlstm_oracle as O  # noqa: E402

pkg.register()
dev = torch.device("cuda", 0)
# multiples of 16: no step-kernel remainder (the reference's native step kernel is the exp-gate one)
for (B, NH, S, D, L) in ((2, 4, 1008, 64, 8), (1, 2, 336, 32, 8), (2, 3, 1616, 128, 8)):
    inp = O.make_inputs(B, NH, S, D, D, seed=S, dtype=torch.float32)
    t = {k: v.to(torch.bfloat16) for k, v in inp.items()}
    be = mLSTMBackend(mLSTMBackendConfig(chunkwise_kernel="chunkwise--b200_siging", sequence_kernel="native_sequence__native",
                                         step_kernel="native", mode="inference", return_last_states=True, chunk_size=64,
                                         eps=1e-6, autocast_kernel_dtype="bfloat16", inference_state_dtype="float32"))
    with torch.no_grad():
        out = be(q=t["q"].to(dev), k=t["k"].to(dev), v=t["v"].to(dev), i=t["i"].to(dev), f=t["f"].to(dev))
    h, states = (out if isinstance(out, tuple) else (out, None))
    d = {k: v.double() for k, v in t.items()}
    h_ref, _, _, last, _ = O.chunkwise_fw(d["q"], d["k"], d["v"], d["i"], d["f"], chunk_size=L, siging=True)
    line = {"shape": [B, NH, S, D], "calls": "chunk 64 / 32 / 16 kernel calls chained through (C, n) by the reference wrapper",
            "h_rel_err_vs_fp64_oracle": O.rel_err(h.cpu(), h_ref)}
    if states is not None:
        line["c_last_rel_err"] = O.rel_err(states[0].cpu(), last[0])
        line["n_last_rel_err"] = O.rel_err(states[1].cpu(), last[1])
    print(json.dumps(line), flush=True)
