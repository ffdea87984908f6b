"""Pin the oracle (oracle/mlstm_oracle.py) against vectors produced by the reference itself
(tests/golden/make_golden.py) and against its own step recurrence."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mlstm_oracle as O

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if not os.path.basename(p).startswith("vil_layer_"))  # those pin the layer around the kernel (test_cell_gpu.py)
TOL = 1e-11  # float64 vs float64, different summation order only


def _load(path):
    z = np.load(path)
    d = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    B, NH, S, DK, DV, L, with_states, seed = (int(x) for x in z["meta"])
    return d, dict(B=B, NH=NH, S=S, DK=DK, DV=DV, L=L, with_states=bool(with_states), seed=seed)


def test_golden_files_present():
    assert len(GOLD) >= 7


@pytest.mark.parametrize("path", [p for p in GOLD if "padded" not in p and "siging" not in p], ids=os.path.basename)
def test_oracle_matches_reference(path):
    d, meta = _load(path)
    st = meta["with_states"]
    c0, n0, m0 = (d.get("in_c0"), d.get("in_n0"), d.get("in_m0")) if st else (None, None, None)
    h, n_tok, m_tok, last, _ = O.chunkwise_fw(d["in_q"], d["in_k"], d["in_v"], d["in_i"], d["in_f"], c0, n0, m0,
                                              chunk_size=meta["L"])
    assert O.rel_err(h, d["h"]) < TOL
    assert O.rel_err(n_tok, d["n_out"]) < TOL
    assert O.rel_err(m_tok, d["m_out"]) < TOL
    grads = O.chunkwise_bw(d["in_q"], d["in_k"], d["in_v"], d["in_i"], d["in_f"], d["in_dh"], n_tok, m_tok, c0, n0, m0,
                           d.get("in_dc_last") if st else None, chunk_size=meta["L"])
    for name, g in zip(("dq", "dk", "dv", "di", "df"), grads):
        assert O.rel_err(g, d[name]) < TOL, name
    if st:
        assert O.rel_err(last[0], d["c_last"]) < TOL
        assert O.rel_err(last[1], d["n_last"]) < TOL
        assert O.rel_err(last[2], d["m_last"]) < TOL
        assert O.rel_err(grads[5], d["dc0"]) < TOL


def test_oracle_matches_reference_padded():
    """Reference pad wrapper (kernel_wrappers.py:227-264): zero-pad to a multiple of 64, slice back."""
    (path,) = [p for p in GOLD if "padded" in p]
    d, meta = _load(path)
    S, Sp = meta["S"], 128

    def pad(x, dim=2):
        shape = list(x.shape)
        shape[dim] = Sp
        out = x.new_zeros(shape)
        out.narrow(dim, 0, S).copy_(x)
        return out

    q, k, v, i, f, dh = (pad(d[f"in_{n}"]) for n in ("q", "k", "v", "i", "f", "dh"))
    h, n_tok, m_tok, _, _ = O.chunkwise_fw(q, k, v, i, f)
    assert O.rel_err(h[:, :, :S], d["h"]) < TOL
    grads = O.chunkwise_bw(q, k, v, i, f, dh, n_tok, m_tok)
    for name, g in zip(("dq", "dk", "dv", "di", "df"), grads):
        assert O.rel_err(g[:, :, :S], d[name]) < TOL, name


@pytest.mark.parametrize("L", [32, 64])
def test_oracle_matches_reference_siging(L):
    """Sigmoid-input-gate variant vs the reference's quadratic native_siging_custbw (any chunk size)."""
    (path,) = [p for p in GOLD if "siging" in p]
    d, meta = _load(path)
    h, n_tok, m_tok, _, _ = O.chunkwise_fw(d["in_q"], d["in_k"], d["in_v"], d["in_i"], d["in_f"], chunk_size=L, siging=True)
    assert O.rel_err(h, d["h"]) < TOL
    assert float(m_tok.abs().max()) == 0.0
    grads = O.chunkwise_bw(d["in_q"], d["in_k"], d["in_v"], d["in_i"], d["in_f"], d["in_dh"], n_tok, m_tok, chunk_size=L,
                           siging=True)
    for name, g in zip(("dq", "dk", "dv", "di", "df"), grads):
        assert O.rel_err(g, d[name]) < TOL, name


@pytest.mark.parametrize("L", [16, 32, 64])
def test_chunkwise_equals_step_recurrence(L):
    """Independent formulation: h and last states do not depend on the chunk size."""
    inp = O.make_inputs(2, 2, 128, 16, 24, seed=11, dtype=torch.float64, with_states=True)
    h, _, _, last, _ = O.chunkwise_fw(inp["q"], inp["k"], inp["v"], inp["i"], inp["f"], inp["c0"], inp["n0"], inp["m0"],
                                      chunk_size=L)
    hs, last_s = O.step_recurrence(inp["q"], inp["k"], inp["v"], inp["i"], inp["f"], inp["c0"], inp["n0"], inp["m0"])
    assert O.rel_err(h, hs) < 1e-11
    for a, b in zip(last, last_s):
        assert O.rel_err(a, b) < 1e-11


def test_split_sequence_continuation_is_exact():
    inp = O.make_inputs(1, 2, 256, 16, 16, seed=12, dtype=torch.float64)
    h, _, _, last, _ = O.chunkwise_fw(inp["q"], inp["k"], inp["v"], inp["i"], inp["f"])
    a = {k: v[:, :, :128] for k, v in inp.items()}
    b = {k: v[:, :, 128:] for k, v in inp.items()}
    h1, _, _, l1, _ = O.chunkwise_fw(a["q"], a["k"], a["v"], a["i"], a["f"])
    h2, _, _, l2, _ = O.chunkwise_fw(b["q"], b["k"], b["v"], b["i"], b["f"], *l1)
    assert O.rel_err(torch.cat([h1, h2], 2), h) < 1e-12
    assert O.rel_err(l2[0], last[0]) < 1e-12


def test_flop_count_config2():
    fwd, bwd = O.algorithmic_flops(32, 4, 1600, 64, 64, 64)
    assert abs((fwd + bwd) / 1e9 - 23.49) < 0.01
