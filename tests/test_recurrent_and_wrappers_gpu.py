"""SURVEY.md section 8(f) #3 (gate soft cap inside the kernels) and #4 (inference / state API): the recurrent step and
sequence kernels against the oracle's token-by-token recurrence (native_step.py:8-101), and -- where the unmodified
reference is staged under baseline/_ref (tools/stage_reference.sh; it travels to the GPU box with the snapshot) -- the
B200 kernels BEHIND the reference's own wrappers:

  * mLSTMBackend + wrap_chunkwise__pad_zeros            (backend_module.py:131-231, kernel_wrappers.py:204-265)
  * wrap_chunkwise__arbitrary_sequence_length           (kernel_wrappers.py:12-201): chained chunk-64/32/16 kernel
    calls handing (C, n) to each other plus a step-kernel remainder, with chunkwise--b200_siging, step_kernel="b200"
  * the fused flip-free branch vs the reference's own ViLLayer.mlstm_branch inside the reference model
"""
import os
import sys

import pytest
import torch

from oracle import mlstm_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mlstm_kernels")),
                               reason="reference not staged under baseline/_ref (tools/stage_reference.sh)")


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as G

    G.build()
    import xlstm_yolo_clean_b200 as p

    return p


@pytest.fixture(scope="module")
def ref(pkg):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import model_bench as MB

    MB._import_reference()
    pkg.register()
    return MB


# ------------------------------------------------------------------------------------------------ recurrent kernels
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2), (torch.float16, 2e-2)],
                         ids=["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("D", [32, 64, 128])
def test_recurrent_sequence_matches_step_recurrence(pkg, D, dtype, tol):
    inp = O.make_inputs(2, 3, 40, D, D, seed=11 + D, dtype=torch.float32, with_states=True)
    r = {k: v.to(dtype).double() for k, v in inp.items()}
    h_ref, (c_ref, n_ref, m_ref) = O.step_recurrence(r["q"], r["k"], r["v"], r["i"], r["f"], r["c0"], r["n0"], r["m0"])
    t = {k: v.to(dtype).cuda() for k, v in inp.items()}
    h, (c, n, m) = pkg.mlstm_recurrent_sequence__b200(t["q"], t["k"], t["v"], t["i"], t["f"], t["c0"], t["n0"], t["m0"],
                                                      return_last_states=True)
    torch.cuda.synchronize()
    assert c.dtype == torch.float32 and m.shape == (2, 3, 1)
    for name, a, b in (("h", h, h_ref), ("c", c, c_ref), ("n", n, n_ref), ("m", m, m_ref)):
        assert O.rel_err(a.double().cpu().reshape(b.shape), b) < tol, name
    # no initial states, no last states: returns h alone (the reference's convention)
    h2 = pkg.mlstm_recurrent_sequence__b200(t["q"], t["k"], t["v"], t["i"], t["f"])
    h2_ref, _ = O.step_recurrence(r["q"], r["k"], r["v"], r["i"], r["f"])
    assert isinstance(h2, torch.Tensor) and O.rel_err(h2.double().cpu(), h2_ref) < tol


def test_step_kernel_chains_like_the_sequence(pkg):
    """Single steps fed with their own states reproduce the in-kernel loop bit for bit (fp32), and strided views of a
    (B, S, NH, D) layout are taken as they are."""
    B, NH, S, D = 2, 4, 9, 64
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(3, B, S, NH, D, generator=g).cuda()
    q, k, v = (x.transpose(1, 2) for x in qkv)  # (B, NH, S, D) views, token stride NH * D
    i, f = torch.randn(B, NH, S, generator=g).cuda(), (2 + torch.randn(B, NH, S, generator=g)).cuda()
    h_seq, (c_seq, n_seq, m_seq) = pkg.mlstm_recurrent_sequence__b200(q, k, v, i, f, return_last_states=True)
    c = torch.zeros(B, NH, D, D, device="cuda")
    n = torch.zeros(B, NH, D, device="cuda")
    m = torch.zeros(B, NH, 1, device="cuda")
    hs = []
    for t in range(S):
        h, (c, n, m) = pkg.mlstm_recurrent_step__b200(q[:, :, t], k[:, :, t], v[:, :, t], i[:, :, t, None], f[:, :, t, None], c, n, m)
        hs.append(h)
    torch.cuda.synchronize()
    assert torch.equal(torch.stack(hs, 2), h_seq) and torch.equal(c, c_seq) and torch.equal(n, n_seq) and torch.equal(m, m_seq)


def test_recurrent_continues_a_chunkwise_call(pkg):
    """Chunkwise kernel over the first 128 tokens, its last states into the recurrent kernel for the remaining 7:
    equals the oracle on all 135 tokens (what the reference's inference wrapper does with the remainder)."""
    B, NH, D = 2, 2, 64
    inp = O.make_inputs(B, NH, 135, D, D, seed=3, dtype=torch.float32)
    r = {k: v.to(torch.bfloat16).double() for k, v in inp.items()}
    h_ref, _ = O.step_recurrence(r["q"], r["k"], r["v"], r["i"], r["f"])
    t = {k: v.to(torch.bfloat16).cuda() for k, v in inp.items()}
    a = {k: v[:, :, :128] for k, v in t.items()}
    b = {k: v[:, :, 128:] for k, v in t.items()}
    h1, (c, n, m) = pkg.mlstm_chunkwise__b200(a["q"], a["k"], a["v"], a["i"], a["f"], return_last_states=True)
    h2 = pkg.mlstm_recurrent_sequence__b200(b["q"], b["k"], b["v"], b["i"], b["f"], c, n, m)
    torch.cuda.synchronize()
    assert O.rel_err(torch.cat([h1, h2], 2).double().cpu(), h_ref) < 2e-2


def test_recurrent_refuses_cpu(pkg):
    x = torch.randn(1, 1, 4, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pkg.mlstm_recurrent_sequence__b200(x, x, x, x[..., 0], x[..., 0])


# ------------------------------------------------------------------------------------------------ gate soft cap
@pytest.mark.parametrize("D,S", [(64, 384), (32, 200), (128, 256)])
@pytest.mark.parametrize("reverse", [False, True], ids=["causal", "anticausal"])
@pytest.mark.parametrize("siging", [False, True], ids=["exp", "siging"])
def test_in_kernel_soft_cap_equals_capping_first(pkg, D, S, reverse, siging):
    """gate_soft_cap: the scan warp applies cap * tanh(x / cap) (vision_lstm2.py:714-715, 755-756) to the pre-activations
    and the backward stores dI / dF w.r.t. the pre-activations.  Same forward bit for bit as capping in torch first (fp32
    tanh of the same 16-bit values, rounded once more only in the torch composition -- hence a tolerance, not equality),
    and the same gate gradients as autograd through the torch cap."""
    cap = 15.0
    g = torch.Generator().manual_seed(D + S)
    dt = torch.bfloat16
    q, k, v, dh = (torch.randn(2, 3, S, D, generator=g).to(dt).cuda() for _ in range(4))
    # pre-activations in the non-linear part of the cap (|x| ~ 4 .. 12 of 15: tanh deviates 2 .. 17 % from the identity)
    # with the well-conditioned statistics of a random-init cell (input gate around -6, forget gate around +6); widely
    # spread input gates make h = num / max(|den|, e^-m) ill-conditioned (|h| > 100) whatever computes it
    pre_i = (-6.0 + torch.randn(2, 3, S, generator=g)).to(dt).cuda()
    pre_f = (6.0 + 2.0 * torch.randn(2, 3, S, generator=g)).to(dt).cuda()
    h, n_out, m_out, _, cst = pkg.mlstm_chunkwise_fw(q, k, v, pre_i, pre_f, chunk_size=4, reverse=reverse, siging=siging,
                                                     gate_soft_cap=cap)
    dq, dk, dv, di, df, _ = pkg.mlstm_chunkwise_bw(q, k, v, pre_i, pre_f, n_out, m_out, dh, chunk_size=4, c_states=cst,
                                                   reverse=reverse, siging=siging, gate_soft_cap=cap)
    # reference composition: cap in fp32 torch (no extra rounding), kernel without cap on fp32 gates is not possible on
    # the tensor route, so compare against the fp64 oracle fed with the capped gates
    ci = cap * torch.tanh(pre_i.double().cpu() / cap)
    cf = cap * torch.tanh(pre_f.double().cpu() / cap)
    seq = lambda x: x.flip(2) if reverse else x  # noqa: E731
    r = [seq(x.double().cpu()) for x in (q, k, v)] + [seq(ci), seq(cf), seq(dh.double().cpu())]
    h_ref, _, grads = O.fwbw(*r, chunk_size=4, siging=siging)
    want = dict(h=h_ref, dq=grads[0], dk=grads[1], dv=grads[2],
                di=grads[3] * seq(1 - torch.tanh(pre_i.double().cpu() / cap) ** 2),
                df=grads[4] * seq(1 - torch.tanh(pre_f.double().cpu() / cap) ** 2))
    got = dict(h=h, dq=dq, dk=dk, dv=dv, di=di, df=df)
    torch.cuda.synchronize()
    for name in want:
        assert O.rel_err(seq(got[name].double().cpu()), want[name]) < 2e-2, name


def test_soft_cap_is_refused_on_the_exact_route(pkg):
    q = torch.randn(1, 2, 64, 64, device="cuda")
    g = torch.randn(1, 2, 64, device="cuda")
    with pytest.raises(RuntimeError, match="tensor-core route"):
        pkg.mlstm_chunkwise_fw(q, q, q, g, g, gate_soft_cap=15.0)


# ------------------------------------------------------------------------------------------------ reference wrappers
@needs_ref
@pytest.mark.parametrize("shape", [(2, 4, 100, 64), (1, 12, 400, 32), (2, 6, 448, 128)], ids=["S100_d64", "S400_d32", "S448_d128"])
def test_drop_in_through_backend_and_pad_wrapper(pkg, ref, shape):
    """mLSTMBackend(chunkwise--b200, mode='train_with_padding') -- exactly what patch_model installs -- on BSHD-strided
    fp16 views like MatrixLSTMCell.forward creates, forward and input gradients vs the oracle on the zero-padded problem."""
    from mlstm_kernels.torch.backend_module import mLSTMBackend, mLSTMBackendConfig

    B, NH, S, D = shape
    be = mLSTMBackend(mLSTMBackendConfig(chunkwise_kernel="chunkwise--b200", sequence_kernel="native_sequence__native",
                                         step_kernel="native", mode="train_with_padding", return_last_states=False,
                                         chunk_size=64, eps=1e-6, autocast_kernel_dtype="bfloat16"))
    g = torch.Generator().manual_seed(S)
    qk = (0.5 * torch.randn(B, S, 2 * NH * D, generator=g)).half().cuda().requires_grad_(True)
    vv = (0.5 * torch.randn(B, S, NH * D, generator=g)).half().cuda().requires_grad_(True)
    gates = torch.cat([-8.7 + torch.randn(B, S, NH, generator=g), 4.0 + torch.randn(B, S, NH, generator=g)], -1).half().cuda().requires_grad_(True)
    q = qk[..., :NH * D].view(B, S, NH, D).transpose(1, 2)
    k = qk[..., NH * D:].view(B, S, NH, D).transpose(1, 2)
    v = vv.view(B, S, NH, D).transpose(1, 2)
    i, f = gates[..., :NH].transpose(1, 2), gates[..., NH:].transpose(1, 2)
    dh = torch.randn(B, NH, S, D, generator=g).half().cuda()
    with torch.autocast("cuda", dtype=torch.float16):
        h = be(q=q, k=k, v=v, i=i, f=f)
    assert h.shape == (B, NH, S, D)
    h.backward(dh.to(h.dtype))
    torch.cuda.synchronize()
    # oracle on the bf16-rounded inputs the pad wrapper hands to the kernel (kernel_wrappers.py:249-262)
    r = {n: t.detach().to(torch.bfloat16).double().cpu() for n, t in dict(q=q, k=k, v=v, i=i, f=f, dh=dh).items()}
    pad = (-S) % 64
    rp = {n: torch.nn.functional.pad(t, (0, 0, 0, pad) if t.dim() == 4 else (0, pad)) for n, t in r.items()}
    h_ref, _, grads = O.fwbw(rp["q"], rp["k"], rp["v"], rp["i"], rp["f"], rp["dh"], chunk_size=64)
    assert O.rel_err(h.double().cpu(), h_ref[:, :, :S]) < 2e-2
    dq_ref = grads[0][:, :, :S].transpose(1, 2).reshape(B, S, NH * D)
    dk_ref = grads[1][:, :, :S].transpose(1, 2).reshape(B, S, NH * D)
    dv_ref = grads[2][:, :, :S].transpose(1, 2).reshape(B, S, NH * D)
    assert O.rel_err(qk.grad[..., :NH * D].double().cpu(), dq_ref) < 2e-2
    assert O.rel_err(qk.grad[..., NH * D:].double().cpu(), dk_ref) < 2e-2
    assert O.rel_err(vv.grad.double().cpu(), dv_ref) < 2e-2
    di_ref, df_ref = grads[3][:, :, :S].transpose(1, 2), grads[4][:, :, :S].transpose(1, 2)
    assert O.rel_err(gates.grad[..., :NH].double().cpu(), di_ref) < 2e-2
    assert O.rel_err(gates.grad[..., NH:].double().cpu(), df_ref) < 2e-2


@needs_ref
@pytest.mark.parametrize("shape,step", [((2, 4, 1008, 64), "native"), ((1, 2, 336, 32), "native"), ((2, 3, 1616, 128), "native"),
                                        ((2, 4, 1003, 64), "b200")], ids=["d64", "d32", "d128", "d64_remainder_b200_step"])
def test_behind_the_arbitrary_sequence_length_wrapper(pkg, ref, shape, step):
    """mode='inference': wrap_chunkwise__arbitrary_sequence_length splits S into chunk-64 / 32 / 16 kernel calls that hand
    (C, n) to each other (it unpacks 2-tuple states: only the sigmoid-input-gate kernels fit) and runs
    the remainder through the sequence / step kernel.  The reference's native step kernel is the exp-gate one, so a
    remainder is only exact with the B200 sequence kernel in siging mode: sequence_kernel='native_sequence__b200' is
    registered with the siging flag bound for that purpose."""
    from functools import partial

    from mlstm_kernels.torch.backend_module import mLSTMBackend, mLSTMBackendConfig
    from mlstm_kernels.torch.recurrent import registry_sequence, registry_step

    B, NH, S, D = shape
    registry_sequence["native_sequence__b200_siging"] = partial(pkg.mlstm_recurrent_sequence__b200, siging=True)
    registry_step["b200_siging"] = partial(pkg.mlstm_recurrent_step__b200, siging=True)
    seq_kernel = "native_sequence__native" if step == "native" else "native_sequence__b200_siging"
    be = mLSTMBackend(mLSTMBackendConfig(chunkwise_kernel="chunkwise--b200_siging", sequence_kernel=seq_kernel,
                                         step_kernel="native" if step == "native" else "b200_siging", mode="inference",
                                         return_last_states=True, chunk_size=64, eps=1e-6, autocast_kernel_dtype="bfloat16",
                                         inference_state_dtype="float32"))
    inp = O.make_inputs(B, NH, S, D, D, seed=S, dtype=torch.float32)
    t = {k: v.to(torch.bfloat16) for k, v in inp.items()}
    with torch.no_grad():
        out = be(q=t["q"].cuda(), k=t["k"].cuda(), v=t["v"].cuda(), i=t["i"].cuda(), f=t["f"].cuda())
    torch.cuda.synchronize()
    # this fork binds return_last_states=False into the backend's inference function (backend_module.py:103-117): h only
    assert isinstance(out, torch.Tensor)
    d = {k: v.double() for k, v in t.items()}
    L = 8 if S % 8 == 0 else 1
    h_ref, _, _, last, _ = O.chunkwise_fw(d["q"], d["k"], d["v"], d["i"], d["f"], chunk_size=L, siging=True)
    assert O.rel_err(out.double().cpu(), h_ref) < 2e-2
    # the wrapper itself, asked for the last states
    from mlstm_kernels.torch.kernel_wrappers import wrap_chunkwise__arbitrary_sequence_length

    h, states = wrap_chunkwise__arbitrary_sequence_length(
        mlstm_chunkwise_kernel=pkg.mlstm_siging_chunkwise__b200,
        mlstm_sequence_kernel=partial(pkg.mlstm_recurrent_sequence__b200, siging=True),
        mlstm_step_kernel=partial(pkg.mlstm_recurrent_step__b200, siging=True),
        q=t["q"].cuda(), k=t["k"].cuda(), v=t["v"].cuda(), i=t["i"].cuda(), f=t["f"].cuda(), return_last_states=True,
        chunk_size=64, eps=1e-6, autocast_kernel_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(h, out)
    assert O.rel_err(states[0].double().cpu(), last[0]) < 2e-2 and O.rel_err(states[1].double().cpu(), last[1]) < 2e-2


@needs_ref
def test_fused_branch_inside_the_reference_model(pkg, ref):
    """The fused flip-free branch (vil.mlstm_branch_b200: anti-causal kernel, rotated conv, in-kernel soft cap, fused cell
    output) against the reference's OWN ViLLayer.mlstm_branch inside the unmodified 640-base256 model: every branch call
    of one forward (eval) and one training forward, output and input gradient."""
    import types

    dev = torch.device("cuda", 0)
    model = ref._build_model("640-base256.yaml", dev)
    pkg.patch_model(model, siging=False)  # reference modules + chunkwise--b200 (exp gate: the oracle's function)
    stats = []

    def rel(a, b):
        a, b = a.float(), b.float()
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

    for mod in model.modules():
        if hasattr(mod, "mlstm_cell") and hasattr(mod, "proj_up"):
            orig = mod.mlstm_branch

            def both(self, x, _orig=orig):
                with torch.enable_grad():
                    xa = x.detach().clone().requires_grad_(True)
                    xb = x.detach().clone().requires_grad_(True)
                    ya = _orig(xa)
                    yb = pkg.mlstm_branch_b200(self, xb, siging=False)
                    gr = torch.randn_like(ya)
                    (ga,) = torch.autograd.grad(ya, xa, gr)
                    (gb,) = torch.autograd.grad(yb, xb, gr)
                if bool(torch.isfinite(ya).all()):
                    stats.append((x.shape[1], rel(yb, ya), rel(gb, ga)))
                return _orig(x)

            mod.mlstm_branch = types.MethodType(both, mod)
    x = ref._batch(2, dev, 0)["img"]
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        model(x)
    assert len(stats) == 20  # 10 block pairs x 2 directions
    worst_y, worst_g = max(s[1] for s in stats), max(s[2] for s in stats)
    assert worst_y < 2e-2 and worst_g < 2e-2, (worst_y, worst_g, stats)
