"""CPU oracle for the mLSTM chunkwise forward / backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; it
is used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the *checker* and the CPU timing
baseline, never as a compute path of the shipped backend.

It is a restatement (not a copy) of the reference's native-torch chunkwise
mLSTM, written from the math in SURVEY.md Appendix A.  Each function names the
reference lines it follows (paths relative to the reference root):

  gates            mlstm_kernels/torch/chunkwise/native/fw.py:257-262, 81-83
  inter-chunk      mlstm_kernels/torch/chunkwise/native/fw.py:29-128
  intra + combine  mlstm_kernels/torch/chunkwise/native/fw.py:131-221
  backward dC      mlstm_kernels/torch/chunkwise/native/bw.py:31-103
  backward dQKV    mlstm_kernels/torch/chunkwise/native/bw.py:106-203
  gate grads       mlstm_kernels/torch/chunkwise/native/bw.py:319-337
  step recurrence  mlstm_kernels/torch/recurrent/native_step.py:8-101

Pinning: the reference holds no golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` and committed as
``tests/golden/*.npz`` (``tests/test_oracle_golden.py`` checks every one), and
against the step recurrence below, an independent formulation.

All tensors are laid out (B, NH, S, D) / (B, NH, S) like the reference API.
The arithmetic runs in the dtype of the inputs (use float64 for goldens).
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class ChunkGates:
    """Per-chunk gate vectors, all shaped (B, NH, NC, L)."""

    b: torch.Tensor  # inclusive cumsum of logsigmoid(f) inside the chunk
    a: torch.Tensor  # g - b + i
    g: torch.Tensor  # (B, NH, NC) total log-forget of the chunk
    i: torch.Tensor


def chunk_gates(i: torch.Tensor, f: torch.Tensor, L: int, siging: bool = False) -> ChunkGates:
    """fw.py:257-262 (b) and fw.py:81-83 (a, g).  ``siging``: the input gate is logsigmoid(i)
    (triton_xl_chunk_siging/chunkwise_gates.py:15-47, parallel/native_siging/fw.py:50-52)."""
    if siging:
        i = F.logsigmoid(i)
    B, NH, S = f.shape
    assert S % L == 0, f"Sequence length {S} is not divisible by chunk size {L}."
    lf = F.logsigmoid(f).reshape(B, NH, S // L, L)
    b = torch.cumsum(lf, dim=-1)
    g = b[..., -1]
    ic = i.reshape(B, NH, S // L, L)
    a = g[..., None] - b + ic
    return ChunkGates(b=b, a=a, g=g, i=ic)


def inter_chunk_states(k, v, gates: ChunkGates, c0=None, n0=None, m0=None, siging=False):
    """State recurrence over chunk boundaries, fw.py:29-128.

    Returns C (B,NH,NC+1,DK,DV), n (B,NH,NC+1,DK), m (B,NH,NC+1); index j holds
    the state *entering* chunk j, index NC the final state.
    """
    B, NH, S, DK = k.shape
    DV = v.shape[-1]
    NC, L = gates.b.shape[2], gates.b.shape[3]
    dt, dev = k.dtype, k.device
    C = torch.zeros(B, NH, NC + 1, DK, DV, dtype=dt, device=dev)
    n = torch.zeros(B, NH, NC + 1, DK, dtype=dt, device=dev)
    m = torch.zeros(B, NH, NC + 1, dtype=dt, device=dev)
    if c0 is not None:
        C[:, :, 0] = c0
    if n0 is not None:
        n[:, :, 0] = n0
    if m0 is not None:
        m[:, :, 0] = m0.reshape(B, NH)
    kc = k.reshape(B, NH, NC, L, DK)
    vc = v.reshape(B, NH, NC, L, DV)
    a_max = gates.a.amax(dim=-1)
    for j in range(NC):
        m_next = torch.maximum(gates.g[:, :, j] + m[:, :, j], a_max[:, :, j])  # fw.py:96-98
        if siging:  # no max state: every gate factor is <= 1 already
            m_next = torch.zeros_like(m_next)
        decay = torch.exp(gates.g[:, :, j] + m[:, :, j] - m_next)  # fw.py:106
        w = torch.exp(gates.a[:, :, j] - m_next[..., None])  # fw.py:102
        kw = kc[:, :, j] * w[..., None]  # K is not scaled, fw.py:100
        C[:, :, j + 1] = decay[..., None, None] * C[:, :, j] + torch.einsum("bhld,bhle->bhde", kw, vc[:, :, j])
        n[:, :, j + 1] = decay[..., None] * n[:, :, j] + kw.sum(dim=2)
        m[:, :, j + 1] = m_next
    return C, n, m


def _log_decay_matrix(gates: ChunkGates):
    """logD[t, s] = b_t - b_s + i_s for s <= t, -inf above the diagonal (fw.py:171-175)."""
    L = gates.b.shape[-1]
    logd = gates.b[..., :, None] - gates.b[..., None, :] + gates.i[..., None, :]
    keep = torch.ones(L, L, dtype=torch.bool, device=logd.device).tril()
    return logd.masked_fill(~keep, float("-inf"))


def intra_chunk_outputs(q, k, v, gates: ChunkGates, C, n, m, scale: float, eps: float, siging=False):
    """fw.py:131-221.  C/n/m are the states entering each chunk (first NC entries)."""
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    NC, L = gates.b.shape[2], gates.b.shape[3]
    qc = q.reshape(B, NH, NC, L, DK)
    kc = k.reshape(B, NH, NC, L, DK)
    vc = v.reshape(B, NH, NC, L, DV)
    logd = _log_decay_matrix(gates)
    m_intra = logd.amax(dim=-1)  # fw.py:178-180
    m_inter = gates.b + m[:, :, :NC, None]  # fw.py:183
    m_tok = torch.maximum(m_inter, m_intra)  # fw.py:184
    if siging:  # denominator becomes max(|n|, 1): parallel/native_siging/fw.py:62-66
        m_tok = torch.zeros_like(m_tok)
    d = torch.exp(logd - m_tok[..., None])  # fw.py:189-190
    s = torch.einsum("bhctd,bhcsd->bhcts", qc, kc) * scale  # fw.py:192
    p = s * d  # fw.py:194
    qbar = qc * (torch.exp(m_inter - m_tok) * scale)[..., None]  # fw.py:197-198
    num = torch.einsum("bhctd,bhcde->bhcte", qbar, C[:, :, :NC]) + torch.einsum("bhcts,bhcse->bhcte", p, vc)
    den = torch.einsum("bhctd,bhcd->bhct", qbar, n[:, :, :NC]) + p.sum(dim=-1)  # fw.py:204-206
    n_tok = torch.maximum(den.abs(), torch.exp(-m_tok))  # fw.py:208-210
    h = num / (n_tok[..., None] + eps)  # fw.py:212
    return h.reshape(B, NH, S, DV), n_tok.reshape(B, NH, S), m_tok.reshape(B, NH, S)


def chunkwise_fw(q, k, v, i, f, c0=None, n0=None, m0=None, chunk_size=64, eps=1e-6, scale=None, siging=False):
    """mlstm_chunkwise_fw, fw.py:224-318.

    Returns h, n_out, m_out, (C_last, n_last, m_last), (C_all, n_all, m_all).
    """
    B, NH, S, DK = q.shape
    scale = DK ** -0.5 if scale is None else scale
    gates = chunk_gates(i, f, chunk_size, siging)
    C, n, m = inter_chunk_states(k, v, gates, c0, n0, m0, siging)
    h, n_tok, m_tok = intra_chunk_outputs(q, k, v, gates, C, n, m, scale, eps, siging)
    last = (C[:, :, -1], n[:, :, -1], m[:, :, -1:])
    return h, n_tok, m_tok, last, (C, n, m)


def chunkwise_bw(q, k, v, i, f, dh, n_tok, m_tok, c0=None, n0=None, m0=None, dc_last=None,
                 chunk_size=64, eps=1e-6, scale=None, siging=False):
    """mlstm_chunkwise_bw, bw.py:206-348 (n_tok and every m are constants).

    Returns dq, dk, dv, di, df, dc0 (dc0 only meaningful when c0 was given).
    """
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    L = chunk_size
    NC = S // L
    scale = DK ** -0.5 if scale is None else scale
    gates = chunk_gates(i, f, L, siging)
    C, _, m = inter_chunk_states(k, v, gates, c0, n0, m0, siging)  # bw.py:251-266 (recompute)

    qc = q.reshape(B, NH, NC, L, DK)
    kc = k.reshape(B, NH, NC, L, DK)
    vc = v.reshape(B, NH, NC, L, DV)
    mt = m_tok.reshape(B, NH, NC, L)
    dht = (dh / (n_tok[..., None] + eps)).reshape(B, NH, NC, L, DV)  # bw.py:88-90,135

    bbar = torch.exp(gates.b + m[:, :, :NC, None] - mt)  # bw.py:79-82,186
    abar = torch.exp(gates.a - m[:, :, 1:, None])  # bw.py:187
    qbar = qc * (bbar * scale)[..., None]

    # state gradients, bw.py:67-98.  dC[j] is the gradient w.r.t. the state entering chunk j.
    dC = torch.zeros(B, NH, NC + 1, DK, DV, dtype=q.dtype, device=q.device)
    if dc_last is not None:
        dC[:, :, NC] = dc_last
    for j in range(NC, 0, -1):
        decay = torch.exp(gates.g[:, :, j - 1] + m[:, :, j - 1] - m[:, :, j])
        dC[:, :, j - 1] = decay[..., None, None] * dC[:, :, j] + torch.einsum(
            "bhtd,bhte->bhde", qbar[:, :, j - 1], dht[:, :, j - 1])

    # intra-chunk, bw.py:153-170
    dbar = torch.exp(_log_decay_matrix(gates) - mt[..., None])
    sbar = torch.einsum("bhctd,bhcsd->bhcts", qc, kc) * scale * dbar
    ds = torch.einsum("bhcte,bhcse->bhcts", dht, vc) * dbar
    dv_ = torch.einsum("bhcts,bhcte->bhcse", sbar, dht)
    dq_ = torch.einsum("bhcts,bhcsd->bhctd", ds, kc) * scale
    dk_ = torch.einsum("bhcts,bhctd->bhcsd", ds, qc) * scale
    # inter-chunk, bw.py:181-193
    dv_ = dv_ + torch.einsum("bhcsd,bhcde->bhcse", kc * abar[..., None], dC[:, :, 1:])
    dk_ = dk_ + torch.einsum("bhcse,bhcde->bhcsd", vc * abar[..., None], dC[:, :, 1:])
    dq_ = dq_ + torch.einsum("bhcte,bhcde->bhctd", dht * bbar[..., None], C[:, :, :NC] * scale)
    dq_ = dq_.reshape(B, NH, S, DK)
    dk_ = dk_.reshape(B, NH, S, DK)
    dv_ = dv_.reshape(B, NH, S, DV)

    # gate gradients, bw.py:319-327 (reverse cumulative sum spans the whole sequence)
    acc = (q * dq_ - k * dk_).sum(dim=-1)
    dfbar = acc.flip(-1).cumsum(-1).flip(-1)
    df = dfbar * torch.sigmoid(-f)
    di = (v * dv_).sum(dim=-1)
    if siging:  # chain rule through logsigmoid(i): chunkwise_gates.py:97-98, parallel/native_siging/bw.py
        di = di * torch.sigmoid(-i)
    return dq_, dk_, dv_, di, df, dC[:, :, 0]


def step_recurrence(q, k, v, i, f, c0=None, n0=None, m0=None, eps=1e-6):
    """Token-by-token mLSTM (native_step.py:8-101); independent check of chunkwise_fw."""
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    dt = q.dtype
    C = torch.zeros(B, NH, DK, DV, dtype=dt) if c0 is None else c0.clone()
    n = torch.zeros(B, NH, DK, dtype=dt) if n0 is None else n0.clone()
    m = torch.zeros(B, NH, dtype=dt) if m0 is None else m0.reshape(B, NH).clone()
    lf = F.logsigmoid(f)
    hs = []
    scale = DK ** -0.5
    for t in range(S):
        m_new = torch.maximum(lf[:, :, t] + m, i[:, :, t])
        fa = torch.exp(lf[:, :, t] + m - m_new)
        ia = torch.exp(i[:, :, t] - m_new)
        C = fa[..., None, None] * C + ia[..., None, None] * (k[:, :, t, :, None] * v[:, :, t, None, :])
        n = fa[..., None] * n + ia[..., None] * k[:, :, t]
        qs = q[:, :, t] * scale
        num = torch.einsum("bhd,bhde->bhe", qs, C)
        den = torch.maximum((qs * n).sum(-1).abs(), torch.exp(-m_new)) + eps
        hs.append(num / den[..., None])
        m = m_new
    return torch.stack(hs, dim=2), (C, n, m[..., None])


def fwbw(q, k, v, i, f, dh, c0=None, n0=None, m0=None, dc_last=None, chunk_size=64, eps=1e-6, siging=False):
    """Forward followed by the hand-written backward: what one bench 'step' computes."""
    h, n_tok, m_tok, last, _ = chunkwise_fw(q, k, v, i, f, c0, n0, m0, chunk_size, eps, siging=siging)
    grads = chunkwise_bw(q, k, v, i, f, dh, n_tok, m_tok, c0, n0, m0, dc_last, chunk_size, eps, siging=siging)
    return h, last, grads


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the per-tensor metric of SURVEY.md §8(c)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def algorithmic_flops(B, NH, S, DK, DV, L=64):
    """Dense-contraction FLOPs of fwd+bwd per SURVEY.md §8(d): fwd 4·L·dk·dv + 2·L²·(dk+dv),
    bwd 10·L·dk·dv + 2·L²·(3·dk+2·dv), per chunk."""
    nc = S // L
    fwd = 4 * L * DK * DV + 2 * L * L * (DK + DV)
    bwd = 10 * L * DK * DV + 2 * L * L * (3 * DK + 2 * DV)
    return B * NH * nc * fwd, B * NH * nc * bwd


def make_inputs(B, NH, S, DK, DV, seed=0, dtype=torch.float32, dist="normal", with_states=False):
    """Seeded synthetic inputs of SURVEY.md §8(d).  ``dist='model'`` mimics random-init cells."""
    g = torch.Generator().manual_seed(seed)
    if dist == "normal":
        q, k = (torch.randn(B, NH, S, DK, generator=g) for _ in range(2))
        v = torch.randn(B, NH, S, DV, generator=g)
        i = torch.randn(B, NH, S, generator=g)
        f = torch.randn(B, NH, S, generator=g) + 3.0
    elif dist == "model":
        q, k = (0.3 * torch.randn(B, NH, S, DK, generator=g) for _ in range(2))
        v = 0.3 * torch.randn(B, NH, S, DV, generator=g)
        i = torch.full((B, NH, S), 15.0 * math.tanh(-10.0 / 15.0))
        fh = 15.0 * torch.tanh(torch.linspace(3.0, 6.0, NH) / 15.0)
        f = fh[None, :, None].expand(B, NH, S).contiguous()
    else:
        raise ValueError(dist)
    dh = torch.randn(B, NH, S, DV, generator=g)
    out = dict(q=q, k=k, v=v, i=i, f=f, dh=dh)
    if with_states:
        out["c0"] = torch.randn(B, NH, DK, DV, generator=g)
        out["n0"] = torch.randn(B, NH, DK, generator=g).abs() + 1.0
        out["m0"] = torch.randn(B, NH, 1, generator=g)
        out["dc_last"] = torch.randn(B, NH, DK, DV, generator=g)
    return {k_: t.to(dtype) for k_, t in out.items()}
