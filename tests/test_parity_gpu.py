"""Parity of the CUDA path (through the C-ABI) against the oracle and the reference-generated
golden vectors.  Tolerances are north_star's: 1e-5 relative in fp32, 2e-2 in bf16/fp16, measured
as max|a-b| / max|b| per tensor against the float64 oracle on inputs pre-rounded to the test dtype
(SURVEY.md §8c)."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import mlstm_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 2e-2}
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if "siging" not in p and not os.path.basename(p).startswith("vil_layer_"))
IMPLS = ["exact", "auto"]


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as G

    G.build()
    import xlstm_yolo_clean_b200 as p

    return p


def _run(pkg, inp, dtype, L=64, states=False, strided=False, impl="auto"):
    """Run fwd+bwd through the public API; returns dict of outputs (on CPU, float64)."""
    pkg.set_default_impl(impl)
    dev = torch.device("cuda:0")
    t = {k: v.to(dtype).to(dev) for k, v in inp.items()}
    if strided:  # BSHD-strided views like MatrixLSTMCell.forward creates (vision_lstm2.py:718-727)
        B, NH, S, DK = t["q"].shape
        DV = t["v"].shape[-1]
        qk = torch.empty(B, S, NH, 2 * DK, dtype=dtype, device=dev)
        qk[..., :DK] = t["q"].transpose(1, 2)
        qk[..., DK:] = t["k"].transpose(1, 2)
        t["q"] = qk[..., :DK].transpose(1, 2)
        t["k"] = qk[..., DK:].transpose(1, 2)
        t["v"] = t["v"].transpose(1, 2).contiguous().transpose(1, 2)
        gates = torch.stack([t["i"], t["f"]], dim=-1).transpose(1, 2).contiguous()  # (B, S, NH, 2)
        t["i"] = gates[..., 0].transpose(1, 2)
        t["f"] = gates[..., 1].transpose(1, 2)
        assert not t["q"].is_contiguous() and not t["i"].is_contiguous()
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    kw = {}
    if states:
        c0 = t["c0"].detach().requires_grad_(True)
        kw = dict(c_initial=c0, n_initial=t["n0"], m_initial=t["m0"], return_last_states=True)
    out = pkg.mlstm_chunkwise__b200(**leaves, chunk_size=L, eps=1e-6, autocast_kernel_dtype=torch.float32, **kw)
    res = {}
    if states:
        h, (c_last, n_last, m_last) = out
        torch.autograd.backward([h, c_last], [t["dh"], t["dc_last"].to(c_last.dtype)])
        res.update(c_last=c_last, n_last=n_last, m_last=m_last, dc0=c0.grad)
    else:
        h = out
        h.backward(t["dh"])
    res.update(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad,
               df=leaves["f"].grad)
    torch.cuda.synchronize()
    pkg.set_default_impl("auto")
    return {k: v.detach().double().cpu() for k, v in res.items()}


def _oracle(inp, dtype, L=64, states=False):
    r = {k: v.to(dtype).double() for k, v in inp.items()}
    st = (r["c0"], r["n0"], r["m0"]) if states else (None, None, None)
    h, last, grads = O.fwbw(r["q"], r["k"], r["v"], r["i"], r["f"], r["dh"], *st,
                            dc_last=r.get("dc_last") if states else None, chunk_size=L)
    out = dict(h=h, dq=grads[0], dk=grads[1], dv=grads[2], di=grads[3], df=grads[4])
    if states:
        out.update(c_last=last[0], n_last=last[1], m_last=last[2], dc0=grads[5])
    return out


def _assert_close(got, want, tol, what=""):
    bad = {}
    for k, w in want.items():
        e = O.rel_err(got[k].reshape(w.shape), w)
        if not e < tol:
            bad[k] = e
    assert not bad, f"{what}: rel err above {tol}: {bad}"


@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_golden_fp32(pkg, path):
    """CUDA fp32 path vs vectors produced by the reference itself."""
    z = np.load(path)
    B, NH, S, DK, DV, L, st, _ = (int(x) for x in z["meta"])
    inp = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    want = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("in_") and k not in ("meta", "n_out", "m_out")}
    if "padded" in path:  # the reference pad wrapper zero-pads to a multiple of 64 (kernel_wrappers.py:227-264)
        Sp = 128
        inp = {k: torch.cat([v, v.new_zeros(*v.shape[:2], Sp - S, *v.shape[3:])], dim=2) for k, v in inp.items()}
        got = _run(pkg, inp, torch.float32, L=64)
        got = {k: v[:, :, :S] for k, v in got.items()}
    else:
        got = _run(pkg, inp, torch.float32, L=L, states=bool(st))
        # the vectors the forward saves for the backward are the reference's own (vecN_out = max(|q.n|, exp(-m)),
        # fw.py:208-210, and vecM_out, fw.py:178-184)
        t = {k: v.float().cuda() for k, v in inp.items()}
        kw = dict(c_initial=t["c0"], n_initial=t["n0"], m_initial=t["m0"]) if st else {}
        _, n_out, m_out, _, _ = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"], chunk_size=L, **kw)
        got.update(n_out=n_out.double().cpu(), m_out=m_out.double().cpu())
        want.update(n_out=torch.from_numpy(z["n_out"]), m_out=torch.from_numpy(z["m_out"]))
    _assert_close(got, want, TOL[torch.float32], os.path.basename(path))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("shape", [(2, 4, 256, 64, 64), (1, 3, 192, 32, 32), (1, 2, 128, 128, 128)],
                         ids=["d64", "d32", "d128"])
def test_oracle_parity(pkg, dtype, shape, impl):
    inp = O.make_inputs(*shape, seed=21, dtype=torch.float32)
    got = _run(pkg, inp, dtype, impl=impl)
    _assert_close(got, _oracle(inp, dtype), TOL[dtype], f"{shape} {dtype} {impl}")


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_states_and_dc_last(pkg, dtype, impl):
    inp = O.make_inputs(2, 2, 192, 64, 64, seed=22, dtype=torch.float32, with_states=True)
    got = _run(pkg, inp, dtype, states=True, impl=impl)
    _assert_close(got, _oracle(inp, dtype, states=True), TOL[dtype], f"states {dtype} {impl}")


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_strided_bshd_inputs(pkg, dtype, impl):
    inp = O.make_inputs(2, 4, 256, 64, 64, seed=23, dtype=torch.float32)
    got = _run(pkg, inp, dtype, strided=True, impl=impl)
    _assert_close(got, _oracle(inp, dtype), TOL[dtype], f"strided {dtype} {impl}")


@pytest.mark.parametrize("impl", IMPLS)
def test_model_like_gates(pkg, impl):
    """Random-init cell distribution: i = -8.7 constant, f in [2.96, 5.7] (vision_lstm2.py:755-769)."""
    inp = O.make_inputs(2, 8, 448, 64, 64, seed=24, dtype=torch.float32, dist="model")
    for dtype in (torch.float32, torch.bfloat16):
        got = _run(pkg, inp, dtype, impl=impl)
        _assert_close(got, _oracle(inp, dtype), TOL[dtype], f"model {dtype} {impl}")


@pytest.mark.parametrize("impl", IMPLS)
def test_extreme_gates_stay_finite(pkg, impl):
    """Soft-capped gate range is [-15, 15]; the stabiliser must keep everything finite."""
    inp = O.make_inputs(1, 2, 256, 64, 64, seed=25, dtype=torch.float32)
    inp["i"] = inp["i"] * 8.0
    inp["f"] = (inp["f"] - 3.0) * 8.0
    inp["i"].clamp_(-15, 15)
    inp["f"].clamp_(-15, 15)
    for dtype in (torch.float32, torch.bfloat16):
        got = _run(pkg, inp, dtype, impl=impl)
        assert all(torch.isfinite(v).all() for v in got.values())
        # |b| reaches ~15*64 here, where one fp32 ulp of the cumsum is 6e-5: the fp32 oracle itself is
        # ~5e-5 from float64 on these inputs, so the fp32 bound is relaxed for this stress case only.
        _assert_close(got, _oracle(inp, dtype), 3e-4 if dtype == torch.float32 else 3e-2, f"extreme {dtype} {impl}")


def test_tensor_path_is_one_launch_each_way(pkg):
    """bf16 d=64 must run the tcgen05 kernels: exactly one launch for forward and one for backward
    (the exact family needs 2 and 4), with and without the saved c_states."""
    inp = O.make_inputs(2, 4, 320, 64, 64, seed=31, dtype=torch.float32)
    t = {k: v.to(torch.bfloat16).cuda() for k, v in inp.items()}
    assert pkg.tensor_path_supported(2, 4, 320, 64, 64, torch.bfloat16)
    pkg.set_default_impl("tensor")
    try:
        h, n_out, m_out, _, cst = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])
        assert pkg.last_launch_count() == 1 and cst is not None
        g1 = pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"], c_states=cst)
        assert pkg.last_launch_count() == 1
        g2 = pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"], c_states=None)
        assert pkg.last_launch_count() == 2  # recompute pass + backward
        torch.cuda.synchronize()
        for a, b in zip(g1[:5], g2[:5]):
            assert torch.equal(a, b)
    finally:
        pkg.set_default_impl("auto")
    want = _oracle(inp, torch.bfloat16)
    got = dict(h=h, dq=g1[0], dk=g1[1], dv=g1[2], di=g1[3], df=g1[4])
    _assert_close({k: v.double().cpu() for k, v in got.items()}, want, 2e-2, "tensor path S=320 (ragged 128-tile)")


def test_chunk_size_error(pkg):
    q = torch.randn(1, 1, 100, 64, device="cuda")
    g = torch.randn(1, 1, 100, device="cuda")
    with pytest.raises(AssertionError, match="not divisible"):
        pkg.mlstm_chunkwise__b200(q=q, k=q, v=q, i=g, f=g, chunk_size=64)


def test_autocast_casts_to_kernel_dtype(pkg):
    """Under CUDA autocast inputs are cast to autocast_kernel_dtype (native/fwbw.py:37)."""
    inp = O.make_inputs(1, 2, 128, 64, 64, seed=26, dtype=torch.float32)
    t = {k: v.cuda() for k, v in inp.items()}
    with torch.autocast("cuda", dtype=torch.float16):
        h = pkg.mlstm_chunkwise__b200(q=t["q"], k=t["k"], v=t["v"], i=t["i"], f=t["f"],
                                      autocast_kernel_dtype=torch.bfloat16)
    assert h.dtype == torch.bfloat16
    assert O.rel_err(h, _oracle(inp, torch.bfloat16)["h"]) < 2e-2


@pytest.mark.parametrize("S", [128, 448, 1600, 6400])
@pytest.mark.parametrize("D", [32, 64, 128])
def test_model_call_shapes(pkg, S, D):
    """The four sequence lengths (after x64 padding) and three head dims the YAML models produce
    (SURVEY.md §3.1): 640-base192 (d=32), base256 (d=64), base384 (d=128); bf16, small batch."""
    inp = O.make_inputs(1, 2, S, D, D, seed=40 + D, dtype=torch.float32, dist="model" if S == 6400 else "normal")
    got = _run(pkg, inp, torch.bfloat16)
    _assert_close(got, _oracle(inp, torch.bfloat16), 2e-2, f"S={S} D={D}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("states", [False, True], ids=["plain", "states"])
def test_reverse_direction_equals_flipped_oracle(pkg, dtype, states):
    """reverse=True == flip(mLSTM(flip(inputs))) -- the bottom-right ViL direction (vision_lstm2.py:292-312)
    -- for outputs, every gradient and the boundary states, with no data movement."""
    inp = O.make_inputs(2, 3, 320, 64, 64, seed=60, dtype=torch.float32, with_states=states)
    dev = torch.device("cuda:0")
    t = {k: v.to(dtype).to(dev) for k, v in inp.items()}
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    kw = {}
    if states:
        c0 = t["c0"].detach().requires_grad_(True)
        kw = dict(c_initial=c0, n_initial=t["n0"], m_initial=t["m0"], return_last_states=True)
    out = pkg.mlstm_chunkwise__b200(**leaves, reverse=True, autocast_kernel_dtype=torch.float32, **kw)
    if states:
        h, (c_last, n_last, m_last) = out
        torch.autograd.backward([h, c_last], [t["dh"], t["dc_last"].to(c_last.dtype)])
    else:
        h = out
        h.backward(t["dh"])
    torch.cuda.synchronize()
    seq = ("q", "k", "v", "i", "f", "dh")
    flipped = {k: (v.flip(2) if k in seq else v) for k, v in inp.items()}
    want = _oracle(flipped, dtype, states=states)
    got = dict(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad, df=leaves["f"].grad)
    tol = TOL[dtype]
    for name in got:
        assert O.rel_err(got[name].double().cpu(), want[name].flip(2)) < tol, name
    if states:
        assert O.rel_err(c_last.double().cpu(), want["c_last"]) < tol
        assert O.rel_err(c0.grad.double().cpu(), want["dc0"]) < tol


@pytest.mark.parametrize("dtype,shape", [(torch.float32, (1, 2, 192, 16, 32)), (torch.float32, (2, 2, 256, 64, 64)),
                                         (torch.bfloat16, (2, 4, 320, 64, 64)), (torch.bfloat16, (1, 2, 256, 128, 128))],
                         ids=["fp32-rect", "fp32-d64", "bf16-d64-tensor", "bf16-d128"])
def test_siging_variant(pkg, dtype, shape):
    """Sigmoid-input-gate variant (the reference's CUDA model default) vs the oracle, which is pinned
    against parallel--native_siging_custbw (tests/golden/siging_S192.npz)."""
    inp = O.make_inputs(*shape, seed=70, dtype=torch.float32)
    t = {k: v.to(dtype).cuda() for k, v in inp.items()}
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    h, (c_last, n_last) = pkg.mlstm_siging_chunkwise__b200(**leaves, return_last_states=True,
                                                          autocast_kernel_dtype=torch.float32)
    h.backward(t["dh"])
    torch.cuda.synchronize()
    r = {k: v.to(dtype).double() for k, v in inp.items()}
    hw, last, grads = O.fwbw(r["q"], r["k"], r["v"], r["i"], r["f"], r["dh"], siging=True)
    got = dict(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad, df=leaves["f"].grad,
               c_last=c_last, n_last=n_last)
    want = dict(h=hw, dq=grads[0], dk=grads[1], dv=grads[2], di=grads[3], df=grads[4], c_last=last[0], n_last=last[1])
    _assert_close({k: v.detach().double().cpu() for k, v in got.items()}, want, TOL[dtype], f"siging {dtype} {shape}")


def test_siging_golden_fp32(pkg):
    """CUDA fp32 siging path vs the vector produced by the reference itself."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "siging_S192.npz"))
    t = {k[3:]: torch.from_numpy(z[k]).float().cuda() for k in z.files if k.startswith("in_")}
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    h = pkg.mlstm_siging_chunkwise__b200(**leaves, autocast_kernel_dtype=torch.float32)
    h.backward(t["dh"])
    torch.cuda.synchronize()
    got = dict(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad, df=leaves["f"].grad)
    _assert_close({k: v.detach().double().cpu() for k, v in got.items()}, {k: torch.from_numpy(z[k]) for k in got}, 1e-5,
                  "siging golden")


def test_host_pipeline_matches_device_path(pkg):
    """HostFwBw (pinned host buffers, batch-sliced 3-stream pipeline) == one device-resident call."""
    B, NH, S, D = 6, 4, 320, 64
    inp = O.make_inputs(B, NH, S, D, D, seed=50, dtype=torch.float32)
    host = {k: v.to(torch.bfloat16).pin_memory() for k, v in inp.items()}
    pipe = pkg.HostFwBw(B, NH, S, D, D, n_slices=4)
    out = pkg.HostFwBw.alloc_host(B, NH, S, D, D)
    pipe.run(host, out)
    torch.cuda.synchronize()
    d = {k: v.cuda() for k, v in host.items()}
    h, n_out, m_out, _, cst = pkg.mlstm_chunkwise_fw(d["q"], d["k"], d["v"], d["i"], d["f"])
    g = pkg.mlstm_chunkwise_bw(d["q"], d["k"], d["v"], d["i"], d["f"], n_out, m_out, d["dh"], c_states=cst)
    torch.cuda.synchronize()
    for name, ref in zip(("h", "dq", "dk", "dv", "di", "df"), (h,) + tuple(g[:5])):
        assert torch.equal(out[name], ref.cpu()), name


# ---- full-size (BASELINE config 2) checks through size-independent properties -----------------

CFG2 = (32, 4, 1600, 64, 64)


def test_config2_exact_vs_tensor_and_properties(pkg):
    inp = O.make_inputs(*CFG2, seed=0, dtype=torch.float32)
    bf = _run(pkg, inp, torch.bfloat16, impl="auto")
    ex = _run(pkg, inp, torch.bfloat16, impl="exact")
    for k in bf:  # two independent kernel families agree at bf16 tolerance
        assert O.rel_err(bf[k], ex[k]) < 2e-2, k
    # oracle on a slice of the batch (the op is independent per (b, h))
    sl = {k: v[:2] for k, v in inp.items()}
    want = _oracle(sl, torch.bfloat16)
    _assert_close({k: v[:2] for k, v in bf.items()}, want, 2e-2, "config2 slice")
    # linearity in v: h(2v) = 2 h(v)   (h is linear in v for fixed gates / q / k)
    inp2 = dict(inp)
    inp2["v"] = inp["v"] * 2.0
    bf2 = _run(pkg, inp2, torch.bfloat16, impl="auto")
    assert O.rel_err(bf2["h"], 2.0 * bf["h"]) < 1e-2


def test_split_sequence_continuation_gpu(pkg):
    """Running two halves with state passing equals the full sequence (fp32 exact path)."""
    inp = O.make_inputs(2, 2, 512, 64, 64, seed=27, dtype=torch.float32)
    t = {k: v.cuda() for k, v in inp.items()}
    full, last = pkg.mlstm_chunkwise__b200(q=t["q"], k=t["k"], v=t["v"], i=t["i"], f=t["f"], return_last_states=True,
                                           autocast_kernel_dtype=torch.float32)
    a = {k: v[:, :, :256] for k, v in t.items()}
    b = {k: v[:, :, 256:] for k, v in t.items()}
    h1, l1 = pkg.mlstm_chunkwise__b200(q=a["q"], k=a["k"], v=a["v"], i=a["i"], f=a["f"], return_last_states=True,
                                       autocast_kernel_dtype=torch.float32)
    h2, l2 = pkg.mlstm_chunkwise__b200(q=b["q"], k=b["k"], v=b["v"], i=b["i"], f=b["f"], c_initial=l1[0],
                                       n_initial=l1[1], m_initial=l1[2], return_last_states=True,
                                       autocast_kernel_dtype=torch.float32)
    assert O.rel_err(torch.cat([h1, h2], 2), full) < 1e-5
    assert O.rel_err(l2[0], last[0]) < 1e-5


# ---------------------------------------------------------------------------------------------------
# Head dim 32 (640-base192.yaml: NH=12, DH=32) on the tcgen05 kernels: 64-byte-swizzle operand tiles.
# ---------------------------------------------------------------------------------------------------
def test_d32_tensor_path_one_launch_and_states(pkg):
    """bf16 d=32 runs the tensor kernels (one launch each way), with initial states, dC_last and last states."""
    assert pkg.tensor_path_supported(2, 12, 320, 32, 32, torch.bfloat16)
    inp = O.make_inputs(2, 3, 320, 32, 32, seed=81, dtype=torch.float32, with_states=True)
    pkg.set_default_impl("tensor")
    try:
        got = _run(pkg, inp, torch.bfloat16, states=True, impl="tensor")
        t = {k: v.to(torch.bfloat16).cuda() for k, v in inp.items()}
        h, n_out, m_out, _, cst = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])
        assert pkg.last_launch_count() == 1 and cst is not None
        pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"], c_states=cst)
        assert pkg.last_launch_count() == 1
    finally:
        pkg.set_default_impl("auto")
    _assert_close(got, _oracle(inp, torch.bfloat16, states=True), 2e-2, "d32 states")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_d32_strided_bshd_inputs(pkg, dtype):
    """The BSHD-strided views MatrixLSTMCell.forward creates (vision_lstm2.py:718-727), base192 geometry."""
    inp = O.make_inputs(2, 12, 448, 32, 32, seed=82, dtype=torch.float32, dist="model")
    got = _run(pkg, inp, dtype, strided=True, impl="tensor")
    _assert_close(got, _oracle(inp, dtype), 2e-2, f"d32 strided {dtype}")


def test_d32_extreme_gates(pkg):
    inp = O.make_inputs(1, 2, 256, 32, 32, seed=83, dtype=torch.float32)
    inp["i"] = (inp["i"] * 8.0).clamp_(-15, 15)
    inp["f"] = ((inp["f"] - 3.0) * 8.0).clamp_(-15, 15)
    got = _run(pkg, inp, torch.bfloat16, impl="tensor")
    assert all(torch.isfinite(v).all() for v in got.values())
    _assert_close(got, _oracle(inp, torch.bfloat16), 3e-2, "d32 extreme")


@pytest.mark.parametrize("states", [False, True], ids=["plain", "states"])
def test_d32_reverse_direction(pkg, states):
    inp = O.make_inputs(2, 3, 320, 32, 32, seed=84, dtype=torch.float32, with_states=states)
    dtype = torch.bfloat16
    t = {k: v.to(dtype).cuda() for k, v in inp.items()}
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    kw = {}
    if states:
        c0 = t["c0"].detach().requires_grad_(True)
        kw = dict(c_initial=c0, n_initial=t["n0"], m_initial=t["m0"], return_last_states=True)
    out = pkg.mlstm_chunkwise__b200(**leaves, reverse=True, autocast_kernel_dtype=torch.float32, **kw)
    if states:
        h, (c_last, n_last, m_last) = out
        torch.autograd.backward([h, c_last], [t["dh"], t["dc_last"].to(c_last.dtype)])
    else:
        h = out
        h.backward(t["dh"])
    torch.cuda.synchronize()
    seq = ("q", "k", "v", "i", "f", "dh")
    flipped = {k: (v.flip(2) if k in seq else v) for k, v in inp.items()}
    want = _oracle(flipped, dtype, states=states)
    got = dict(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad, df=leaves["f"].grad)
    for name in got:
        assert O.rel_err(got[name].double().cpu(), want[name].flip(2)) < 2e-2, name
    if states:
        assert O.rel_err(c_last.double().cpu(), want["c_last"]) < 2e-2
        assert O.rel_err(c0.grad.double().cpu(), want["dc0"]) < 2e-2


def test_d32_siging_variant(pkg):
    inp = O.make_inputs(2, 4, 320, 32, 32, seed=85, dtype=torch.float32)
    dtype = torch.bfloat16
    t = {k: v.to(dtype).cuda() for k, v in inp.items()}
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    h, (c_last, n_last) = pkg.mlstm_siging_chunkwise__b200(**leaves, return_last_states=True,
                                                          autocast_kernel_dtype=torch.float32)
    h.backward(t["dh"])
    torch.cuda.synchronize()
    r = {k: v.to(dtype).double() for k, v in inp.items()}
    hw, last, grads = O.fwbw(r["q"], r["k"], r["v"], r["i"], r["f"], r["dh"], siging=True)
    got = dict(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad, df=leaves["f"].grad,
               c_last=c_last, n_last=n_last)
    want = dict(h=hw, dq=grads[0], dk=grads[1], dv=grads[2], di=grads[3], df=grads[4], c_last=last[0], n_last=last[1])
    for name in got:
        assert O.rel_err(got[name].double().cpu(), want[name]) < 2e-2, name


def test_d32_base192_call_exact_vs_tensor(pkg):
    """A base192-sized call (B=4 of 64, NH=12, S=1600): tensor path agrees with the exact fp32-accumulate route."""
    inp = O.make_inputs(4, 12, 1600, 32, 32, seed=86, dtype=torch.float32, dist="model")
    a = _run(pkg, inp, torch.bfloat16, impl="tensor")
    b = _run(pkg, inp, torch.bfloat16, impl="exact")
    for name in a:
        assert O.rel_err(a[name], b[name]) < 2e-2, name


@pytest.mark.parametrize("S", [100, 400, 52, 1004])
@pytest.mark.parametrize("D", [32, 64])
@pytest.mark.parametrize("reverse", [False, True], ids=["causal", "anticausal"])
def test_ragged_sequence_lengths_on_tensor_path(pkg, S, D, reverse):
    """Any S % 4 == 0 runs on the tcgen05 kernels (ragged 128-token tail tiles are handled in-kernel: TMA
    zero-fill, gates masked at scan time), so the model's S = 400 / 100 stages need no zero-padding copies
    (kernel_wrappers.py:227-247).  chunk_size = 4 divides every case; the result does not depend on it."""
    assert pkg.tensor_path_supported(2, 3, S, D, D, torch.bfloat16, chunk_size=4)
    inp = O.make_inputs(2, 3, S, D, D, seed=70 + S, dtype=torch.float32)
    dev = torch.device("cuda:0")
    pkg.set_default_impl("tensor")
    try:
        t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
        leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
        h = pkg.mlstm_chunkwise__b200(**leaves, chunk_size=4, reverse=reverse, autocast_kernel_dtype=torch.bfloat16)
        h.backward(t["dh"])
        torch.cuda.synchronize()
    finally:
        pkg.set_default_impl("auto")
    seq = ("q", "k", "v", "i", "f", "dh")
    src = {k: (v.flip(2) if (reverse and k in seq) else v) for k, v in inp.items()}
    want = _oracle(src, torch.bfloat16, L=4)
    got = dict(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad, df=leaves["f"].grad)
    for name in got:
        w = want[name].flip(2) if reverse else want[name]
        assert O.rel_err(got[name].double().cpu(), w) < 2e-2, (name, O.rel_err(got[name].double().cpu(), w))


def test_ragged_d128_forward_and_siging(pkg):
    inp = O.make_inputs(1, 2, 100, 128, 128, seed=5, dtype=torch.float32)
    dev = torch.device("cuda:0")
    t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
    with torch.no_grad():
        h = pkg.mlstm_chunkwise__b200(q=t["q"], k=t["k"], v=t["v"], i=t["i"], f=t["f"], chunk_size=4)
    assert O.rel_err(h.double().cpu(), _oracle(inp, torch.bfloat16, L=4)["h"]) < 2e-2
    inp = O.make_inputs(2, 2, 400, 64, 64, seed=6, dtype=torch.float32)
    t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    h = pkg.mlstm_siging_chunkwise__b200(**leaves, chunk_size=16)
    h.backward(t["dh"])
    r = {k: v.to(torch.bfloat16).double() for k, v in inp.items()}
    hw, _, g = O.fwbw(r["q"], r["k"], r["v"], r["i"], r["f"], r["dh"], chunk_size=16, siging=True)
    assert O.rel_err(h.double().cpu(), hw) < 2e-2
    for got, want in zip((leaves["q"].grad, leaves["k"].grad, leaves["v"].grad, leaves["i"].grad, leaves["f"].grad), g[:5]):
        assert O.rel_err(got.double().cpu(), want) < 2e-2


@pytest.mark.parametrize("B,NH,S,D", [(1, 1, 4, 64), (1, 3, 8, 32), (1, 1, 60, 64), (7, 1, 124, 32), (1, 2, 132, 64),
                                     (3, 1, 260, 32), (1, 5, 388, 64)])
def test_tiny_and_odd_shapes_on_tensor_path(pkg, B, NH, S, D):
    """Edge shapes of the tcgen05 route: sequences shorter than one 128-token tile, tails of 4 tokens, odd
    batch / head counts; forward, every gradient, fp16 this time."""
    inp = O.make_inputs(B, NH, S, D, D, seed=S, dtype=torch.float32)
    pkg.set_default_impl("tensor")
    try:
        got = _run(pkg, inp, torch.float16, L=4, impl="tensor")
    finally:
        pkg.set_default_impl("auto")
    _assert_close(got, _oracle(inp, torch.float16, L=4), 2e-2, f"{(B, NH, S, D)}")


@pytest.mark.parametrize("states", [False, True], ids=["plain", "states"])
@pytest.mark.parametrize("reverse", [False, True], ids=["causal", "anticausal"])
def test_d128_backward_by_blocks(pkg, states, reverse):
    """Head dim 128 (base384): forward on tc_fw_d128, backward as four 64 x 64 block problems on tc_bw<64>
    (bw128_by_blocks in csrc/tensor_kernels.cu, inside the C-ABI call).  Every gradient, incl. dC_initial with dC_last, both scan directions, a ragged S."""
    S = 324
    inp = O.make_inputs(2, 3, S, 128, 128, seed=128 + S, dtype=torch.float32, with_states=states)
    dev = torch.device("cuda:0")
    t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
    leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    kw = {}
    if states:
        c0 = t["c0"].detach().requires_grad_(True)
        kw = dict(c_initial=c0, n_initial=t["n0"], m_initial=t["m0"], return_last_states=True)
    out = pkg.mlstm_chunkwise__b200(**leaves, chunk_size=4, reverse=reverse, autocast_kernel_dtype=torch.bfloat16, **kw)
    n_before = pkg.last_launch_count()
    if states:
        h, (c_last, n_last, m_last) = out
        torch.autograd.backward([h, c_last], [t["dh"], t["dc_last"].to(c_last.dtype)])
    else:
        h = out
        h.backward(t["dh"])
    torch.cuda.synchronize()
    seq = ("q", "k", "v", "i", "f", "dh")
    src = {k: (v.flip(2) if (reverse and k in seq) else v) for k, v in inp.items()}
    want = _oracle(src, torch.bfloat16, L=4, states=states)
    got = dict(h=h, dq=leaves["q"].grad, dk=leaves["k"].grad, dv=leaves["v"].grad, di=leaves["i"].grad, df=leaves["f"].grad)
    for name in got:
        w = want[name].flip(2) if reverse else want[name]
        assert O.rel_err(got[name].double().cpu(), w) < 2e-2, (name, O.rel_err(got[name].double().cpu(), w))
    if states:
        assert O.rel_err(c0.grad.double().cpu(), want["dc0"]) < 2e-2
    del n_before


def test_d128_backward_matches_exact_family(pkg):
    """Same call through the exact fp32 FFMA kernels (the route d = 128 backward took before) and the block route."""
    inp = O.make_inputs(1, 2, 256, 128, 128, seed=3, dtype=torch.float32)
    a = _run(pkg, inp, torch.bfloat16, impl="exact")
    b = _run(pkg, inp, torch.bfloat16, impl="auto")
    for k in a:
        assert O.rel_err(b[k], a[k]) < 2e-2, k


@pytest.mark.parametrize("reverse", [False, True], ids=["causal", "anticausal"])
def test_d128_backward_recompute_equals_saved_states(pkg, reverse):
    """The head-dim-128 backward with the forward's saved block states and without them (each block problem then
    recomputes its slice, bw.py:251-266) must agree bit for bit: the saved states ARE what the recompute produces."""
    dev = torch.device("cuda:0")
    inp = O.make_inputs(2, 2, 388, 128, 128, seed=9, dtype=torch.float32, with_states=True)
    t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
    kw = dict(c_initial=t["c0"], n_initial=t["n0"], m_initial=t["m0"])
    h, n_out, m_out, _, cst = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"], chunk_size=4, reverse=reverse, **kw)
    assert cst is not None
    a = pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"], chunk_size=4, c_states=cst,
                               reverse=reverse, want_dc_initial=True, dc_last=t["dc_last"], **kw)
    b = pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"], chunk_size=4, c_states=None,
                               reverse=reverse, want_dc_initial=True, dc_last=t["dc_last"], **kw)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_views_the_tma_cannot_describe_are_accepted(pkg):
    """The reference takes any strides (its @contiguous decorator copies every argument, mlstm_kernels/torch/utils.py:30-42).
    Views a TMA tensor map cannot describe -- a base pointer at an odd offset, a token stride that is not a multiple of
    16 bytes, a broadcast batch -- are copied by the host side (AUTO route) instead of raising, and give bit-identical
    results to the contiguous call; at the C-ABI the same views fall through to the any-stride exact kernels."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    B, NH, S, D = 2, 3, 256, 64
    big = torch.randn(B, NH, S, D + 16, generator=g).to(torch.bfloat16).to(dev)
    q = big[..., 3:3 + D]                                     # base pointer 6 bytes off
    k = torch.randn(1, NH, S, D, generator=g).to(torch.bfloat16).to(dev).expand(B, NH, S, D)  # stride-0 batch
    odd = torch.randn(B, NH, S, D + 4, generator=g).to(torch.bfloat16).to(dev)
    v = odd[..., :D]                                          # token stride 68 elements = 136 bytes
    i = torch.randn(B, NH, S, generator=g).to(torch.bfloat16).to(dev)
    f = (3.0 + torch.randn(B, NH, S, generator=g)).to(torch.bfloat16).to(dev)
    dh = torch.randn(B, NH, S, D + 4, generator=g).to(torch.bfloat16).to(dev)[..., 2:2 + D]
    assert q.data_ptr() % 16 and v.stride(2) % 8 and k.stride(0) == 0

    def run(q, k, v, dh):
        leaves = [t.detach().requires_grad_(True) for t in (q, k, v, i, f)]
        h = pkg.mlstm_chunkwise__b200(*leaves, chunk_size=64)
        grads = torch.autograd.grad(h, leaves, dh)
        torch.cuda.synchronize()
        return [h] + list(grads)

    got = run(q, k, v, dh)
    want = run(q.contiguous(), k.contiguous(), v.contiguous(), dh.contiguous())
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)
    # C-ABI level: AUTO + misaligned q -> exact kernels (needs the exact family's workspace)
    import ctypes as C

    from xlstm_yolo_clean_b200 import _cabi, backend
    lib = _cabi.load_library()
    a = _cabi.FwArgs()
    a.shape = backend._shape(q, v, 64, 1e-6, _cabi.IMPL_AUTO)
    ex = backend._shape(q, v, 64, 1e-6, _cabi.IMPL_EXACT)
    ws = torch.empty(lib.mlstm_b200_workspace_bytes(C.byref(ex), 0), dtype=torch.uint8, device=dev)
    h = torch.empty(B, NH, S, D, dtype=torch.bfloat16, device=dev)
    nm = torch.empty(2, B, NH, S, dtype=torch.float32, device=dev)
    kc, vc = k.contiguous(), v.contiguous()
    a.q, a.k, a.v, a.i, a.f, a.h = (backend._tensor(t) for t in (q, kc, vc, i, f, h))
    a.n_out, a.m_out = nm[0].data_ptr(), nm[1].data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    st = lib.mlstm_b200_chunkwise_fw(C.byref(a), None)
    torch.cuda.synchronize()
    assert st == 0, lib.mlstm_b200_last_error()
    assert lib.mlstm_b200_last_launch_count() == 2  # the exact family's two forward kernels, not the tensor path's one
    assert O.rel_err(h.double().cpu(), want[0].double().cpu()) < 2e-2


@pytest.mark.parametrize("reverse", [False, True], ids=["causal", "anticausal"])
@pytest.mark.parametrize("B,NH,S,D", [(2, 3, 384, 64), (2, 4, 200, 32), (1, 2, 256, 128)], ids=["d64", "d32_ragged", "d128_blocks"])
@pytest.mark.parametrize("kdt,gdt", [(torch.bfloat16, torch.float16), (torch.float16, torch.bfloat16)], ids=["bf16_kernel_fp16_grads", "fp16_kernel_bf16_grads"])
def test_backward_writes_gradients_in_the_callers_dtype(pkg, kdt, gdt, B, NH, S, D, reverse):
    """shape.grad_dtype: a kernel that computes in one 16-bit dtype rounds its fp32 accumulators straight to the other
    (what autograd's cast of the gradients does under fp16 autocast with the bf16 kernels, native/fwbw.py:37, minus the
    pass over dq / dk / dv): same values as the kernel-dtype gradients up to one rounding, and within the 16-bit bar of
    the fp64 oracle -- for freshly allocated gradients and for caller-owned strided ones."""
    inp = O.make_inputs(B, NH, S, D, D, seed=71, dtype=torch.float32)
    t = {k: v.to(kdt).cuda() for k, v in inp.items()}
    L = math.gcd(S, 64)
    _, n_out, m_out, _, cs = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"], chunk_size=L, reverse=reverse)
    args = (t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"])
    base = pkg.mlstm_chunkwise_bw(*args, chunk_size=L, c_states=cs, reverse=reverse)[:5]
    got = pkg.mlstm_chunkwise_bw(*args, chunk_size=L, c_states=cs, reverse=reverse, grad_dtype=gdt)[:5]
    # caller-owned gradients in the layer layout (B, S, NH, 2D | D | 2), other dtype
    d_qk = torch.empty(B, S, NH, 2 * D, dtype=gdt, device="cuda")
    d_v = torch.empty(B, S, NH, D, dtype=gdt, device="cuda")
    d_g = torch.empty(B, S, NH, 2, dtype=gdt, device="cuda")
    out = (d_qk[..., :D].transpose(1, 2), d_qk[..., D:].transpose(1, 2), d_v.transpose(1, 2), d_g[..., 0].transpose(1, 2),
           d_g[..., 1].transpose(1, 2))
    pkg.mlstm_chunkwise_bw(*args, chunk_size=L, c_states=cs, reverse=reverse, out=out)
    torch.cuda.synchronize()
    r = {k: v.to(kdt).double() for k, v in inp.items()}
    flip = (lambda x: x.flip(2)) if reverse else (lambda x: x)
    _, _, grads = O.fwbw(*(flip(r[n]) for n in ("q", "k", "v", "i", "f", "dh")), None, None, None, chunk_size=L)
    for name, b, g, o, w in zip(("dq", "dk", "dv", "di", "df"), base, got, out, grads):
        assert g.dtype == gdt and o.dtype == gdt and b.dtype == kdt
        assert torch.equal(g, o.contiguous()), name
        assert O.rel_err(g.double().cpu(), b.double().cpu()) < 6e-3, name  # one bf16 rounding apart
        # (dF is a suffix sum: its bf16 bar at long S is documented in test_model_shapes_gpu; S <= 384 here)
        assert O.rel_err(g.double().cpu(), flip(w)) < 2e-2, name


def test_registry_function_under_fp16_autocast_returns_fp16_gradients_without_casts(pkg):
    """fp16 leaves under CUDA autocast with the bf16 kernels (the reference trainer's situation): gradients arrive in
    fp16, written by the backward kernel itself, equal to what the kernel-dtype gradients round to."""
    inp = O.make_inputs(2, 4, 256, 64, 64, seed=72, dtype=torch.float32)
    leaves = {k: inp[k].to(torch.float16).cuda().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
    dh = inp["dh"].cuda()
    with torch.autocast("cuda", dtype=torch.float16):
        h = pkg.mlstm_chunkwise__b200(**leaves, autocast_kernel_dtype=torch.bfloat16)
    assert h.dtype == torch.bfloat16
    h.backward(dh.to(h.dtype))
    ref = _oracle({k: (v.to(torch.float16) if k in leaves else v) for k, v in inp.items()}, torch.bfloat16)
    for n, key in (("q", "dq"), ("k", "dk"), ("v", "dv"), ("i", "di"), ("f", "df")):
        assert leaves[n].grad.dtype == torch.float16
        assert O.rel_err(leaves[n].grad.double().cpu(), ref[key]) < 2e-2, key
