"""Host side of the B200 mLSTM chunkwise backend (Python/PyTorch above the C-ABI).

Mirrors the reference's operator interface for this path so it is a drop-in:

  * ``mlstm_chunkwise__b200`` has the signature, return convention, dtype rule and error
    behaviour of ``mlstm_chunkwise__native_custbw`` (mlstm_kernels/torch/chunkwise/native/fwbw.py:228-263);
  * ``mlstm_chunkwise_fw`` / ``mlstm_chunkwise_bw`` correspond to native/fw.py:224-318 and
    native/bw.py:206-348 and are thin wrappers over the two C-ABI calls;
  * ``register`` / ``patch_model`` hook it into mlstm_kernels.torch.chunkwise.registry and
    ``MatrixLSTMCell`` (ultralytics/nn/modules/vision_lstm/vision_lstm2.py:623-770).

PyTorch is used for device memory, streams and autograd plumbing only.  There is no CPU or
eager fallback: CPU tensors or a missing library raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _cabi

KERNEL_NAME = "b200"  # registry key -> "chunkwise--b200"

_DTYPES = {torch.float32: _cabi.F32, torch.bfloat16: _cabi.BF16, torch.float16: _cabi.F16}
_default_impl = _cabi.IMPL_AUTO
_last_launches = 0


def set_default_impl(name: str) -> None:
    """'auto' | 'exact' (fp32 FFMA kernels) | 'tensor' (tcgen05 kernels, error if unsupported)."""
    global _default_impl
    _default_impl = {"auto": _cabi.IMPL_AUTO, "exact": _cabi.IMPL_EXACT, "tensor": _cabi.IMPL_TENSOR}[name]


def last_launch_count() -> int:
    """Kernels launched by the most recent forward or backward call on this thread."""
    return _cabi.load_library().mlstm_b200_last_launch_count()


def _tensor(t: Optional[torch.Tensor]) -> _cabi.Tensor:
    out = _cabi.Tensor()
    if t is None:
        out.ptr = None
        return out
    out.ptr = t.data_ptr()
    st = t.stride()
    out.stride[:len(st)] = st
    return out


class _on_device:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the common case costs nothing)."""

    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


_BYTES_CACHE: dict = {}


def _scratch_bytes(lib, shape: "_cabi.Shape", backward: int, want_states: bool):
    """(workspace bytes, states bytes) of a shape; the two C-ABI queries are cached per shape key."""
    key = (shape.B, shape.NH, shape.S, shape.DHQK, shape.DHHV, shape.chunk_size, shape.dtype, shape.impl, backward)
    r = _BYTES_CACHE.get(key)
    if r is None:
        r = (lib.mlstm_b200_workspace_bytes(C.byref(shape), backward), lib.mlstm_b200_states_bytes(C.byref(shape)))
        _BYTES_CACHE[key] = r
    return r[0], (r[1] if want_states else 0)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _shape(q, v, chunk_size, eps, impl, qk_scale=None, reverse=False, siging=False) -> _cabi.Shape:
    B, NH, S, DK = q.shape
    s = _cabi.Shape()
    s.B, s.NH, s.S, s.DHQK, s.DHHV = B, NH, S, DK, v.shape[-1]
    s.chunk_size = int(chunk_size)
    s.dtype = _DTYPES[q.dtype]
    s.impl = _default_impl if impl is None else impl
    s.reverse = 1 if reverse else 0
    s.siging = 1 if siging else 0
    s.eps = float(eps)
    s.qk_scale = -1.0 if qk_scale is None else float(qk_scale)
    return s


def tensor_path_supported(B, NH, S, DK, DV, dtype=torch.bfloat16, chunk_size=64) -> bool:
    s = _cabi.Shape()
    s.B, s.NH, s.S, s.DHQK, s.DHHV, s.chunk_size, s.dtype = B, NH, S, DK, DV, chunk_size, _DTYPES[dtype]
    return bool(_cabi.load_library().mlstm_b200_tensor_path_supported(C.byref(s)))


def _rowmajor_last(t: torch.Tensor) -> torch.Tensor:
    """The kernels take any batch/head/token strides but need a unit innermost stride."""
    return t if t.stride(-1) == 1 else t.contiguous()


def _check_inputs(q, k, v, i, f):
    for name, t in (("q", q), ("k", k), ("v", v), ("i", i), ("f", f)):
        if not t.is_cuda:
            raise RuntimeError(f"mlstm_chunkwise__b200: {name} is on {t.device}; this backend has no CPU path")
    if q.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported dtype {q.dtype}")
    B, NH, S, DK = q.shape
    assert k.shape == (B, NH, S, DK), f"k has wrong shape {tuple(k.shape)}"
    assert v.shape[:3] == (B, NH, S), f"v has wrong shape {tuple(v.shape)}"
    assert i.shape == (B, NH, S) and f.shape == (B, NH, S), "i / f must be (B, NH, S)"


def _state_f32(t, shape):
    if t is None:
        return None
    return t.detach().to(torch.float32).reshape(shape).contiguous()


def mlstm_chunkwise_fw(q, k, v, i, f, c_initial=None, n_initial=None, m_initial=None, qk_scale=None,
                       return_last_states=False, chunk_size=64, eps=1e-6, impl=None, save_states=True, reverse=False, siging=False):
    """C-ABI forward.  Returns h, n_out, m_out, last_states-or-None (fp32), c_states-or-None.

    ``c_states`` is the opaque per-tile state buffer the tensor-core backward consumes (the
    reference's return_all_states mode, native/fwbw.py:73-101); None when the kernels recompute."""
    global _last_launches
    lib = _cabi.load_library()
    _check_inputs(q, k, v, i, f)
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    assert S % chunk_size == 0, f"Sequence length {S} is not divisible by chunk size {chunk_size}."
    q, k, v = (_rowmajor_last(t) for t in (q, k, v))
    i = i if i.dtype == q.dtype else i.to(q.dtype)
    f = f if f.dtype == q.dtype else f.to(q.dtype)
    k = k if k.dtype == q.dtype else k.to(q.dtype)
    v = v if v.dtype == q.dtype else v.to(q.dtype)
    dev = q.device
    c0, n0, m0 = _state_f32(c_initial, (B, NH, DK, DV)), _state_f32(n_initial, (B, NH, DK)), _state_f32(m_initial, (B, NH))
    if c0 is not None or n0 is not None or m0 is not None:
        c0 = torch.zeros(B, NH, DK, DV, device=dev) if c0 is None else c0
        n0 = torch.zeros(B, NH, DK, device=dev) if n0 is None else n0
        m0 = torch.zeros(B, NH, device=dev) if m0 is None else m0
    with _on_device(dev):
        h = torch.empty(B, NH, S, DV, dtype=q.dtype, device=dev)
        n_out = torch.empty(B, NH, S, dtype=torch.float32, device=dev)
        m_out = torch.empty(B, NH, S, dtype=torch.float32, device=dev)
        last = None
        if return_last_states:
            last = (torch.empty(B, NH, DK, DV, dtype=torch.float32, device=dev),
                    torch.empty(B, NH, DK, dtype=torch.float32, device=dev),
                    torch.empty(B, NH, 1, dtype=torch.float32, device=dev))
        a = _cabi.FwArgs()
        a.shape = _shape(q, v, chunk_size, eps, impl, qk_scale, reverse, siging)
        ws_bytes, st_bytes = _scratch_bytes(lib, a.shape, 0, save_states)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        c_states = torch.empty(st_bytes, dtype=torch.uint8, device=dev) if st_bytes else None
        a.c_states = _ptr(c_states)
        a.q, a.k, a.v, a.i, a.f, a.h = (_tensor(t) for t in (q, k, v, i, f, h))
        a.c_initial, a.n_initial, a.m_initial = _ptr(c0), _ptr(n0), _ptr(m0)
        a.n_out, a.m_out = n_out.data_ptr(), m_out.data_ptr()
        if last is not None:
            a.c_last, a.n_last, a.m_last = (t.data_ptr() for t in last)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
        st = lib.mlstm_b200_chunkwise_fw(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _cabi.check(st, "mlstm_b200_chunkwise_fw")
    return h, n_out, m_out, last, c_states


def mlstm_chunkwise_bw(q, k, v, i, f, n_out, m_out, dh, c_initial=None, n_initial=None, m_initial=None,
                       dc_last=None, qk_scale=None, chunk_size=64, eps=1e-6, impl=None, want_dc_initial=False,
                       c_states=None, reverse=False, siging=False, out=None):
    """C-ABI backward.  Returns dq, dk, dv, di, df, dc_initial-or-None (fp32).

    ``out`` = (dq, dk, dv, di, df) lets the caller provide the gradient tensors (any batch/head/token strides,
    unit innermost stride for dq/dk/dv), e.g. views into a fused (B, S, 2H) qk gradient."""
    lib = _cabi.load_library()
    _check_inputs(q, k, v, i, f)
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    q, k, v = (_rowmajor_last(t) for t in (q, k, v))
    dh = _rowmajor_last(dh if dh.dtype == q.dtype else dh.to(q.dtype))
    i = i if i.dtype == q.dtype else i.to(q.dtype)
    f = f if f.dtype == q.dtype else f.to(q.dtype)
    dev = q.device
    c0, n0, m0 = _state_f32(c_initial, (B, NH, DK, DV)), _state_f32(n_initial, (B, NH, DK)), _state_f32(m_initial, (B, NH))
    if c0 is not None or n0 is not None or m0 is not None:
        c0 = torch.zeros(B, NH, DK, DV, device=dev) if c0 is None else c0
        n0 = torch.zeros(B, NH, DK, device=dev) if n0 is None else n0
        m0 = torch.zeros(B, NH, device=dev) if m0 is None else m0
    dcl = _state_f32(dc_last, (B, NH, DK, DV))
    with _on_device(dev):
        if out is not None:
            dq, dk, dv, di, df = out
            assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == v.shape and di.shape == i.shape
            assert all(t.dtype == q.dtype and t.device == dev for t in out)
        else:
            dq = torch.empty(B, NH, S, DK, dtype=q.dtype, device=dev)
            dk = torch.empty(B, NH, S, DK, dtype=q.dtype, device=dev)
            dv = torch.empty(B, NH, S, DV, dtype=q.dtype, device=dev)
            di = torch.empty(B, NH, S, dtype=q.dtype, device=dev)
            df = torch.empty(B, NH, S, dtype=q.dtype, device=dev)
        dc0 = torch.empty(B, NH, DK, DV, dtype=torch.float32, device=dev) if want_dc_initial else None
        a = _cabi.BwArgs()
        a.shape = _shape(q, v, chunk_size, eps, impl, qk_scale, reverse, siging)
        ws_bytes, _ = _scratch_bytes(lib, a.shape, 1, False)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a.q, a.k, a.v, a.i, a.f, a.dh = (_tensor(t) for t in (q, k, v, i, f, dh))
        a.c_initial, a.n_initial, a.m_initial = _ptr(c0), _ptr(n0), _ptr(m0)
        a.n_out, a.m_out = n_out.data_ptr(), m_out.data_ptr()
        a.c_states = _ptr(c_states)
        a.dc_last = _ptr(dcl)
        a.dq, a.dk, a.dv, a.di, a.df = (_tensor(t) for t in (dq, dk, dv, di, df))
        a.dc_initial = _ptr(dc0)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
        st = lib.mlstm_b200_chunkwise_bw(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _cabi.check(st, "mlstm_b200_chunkwise_bw")
    return dq, dk, dv, di, df, dc0


def _make_function(autocast_kernel_dtype: torch.dtype):
    class _MlstmChunkwiseB200(torch.autograd.Function):
        """autograd.Function with the contract of _mlstm_chunkwise_fwbw (native/fwbw.py:35-171)."""

        @staticmethod
        @custom_fwd(device_type="cuda", cast_inputs=autocast_kernel_dtype)
        def forward(ctx, q, k, v, i, f, c_initial, n_initial, m_initial, return_last_states, chunk_size, eps, reverse,
                    siging):
            need_bw = any(ctx.needs_input_grad[:6])
            h, n_out, m_out, last, c_states = mlstm_chunkwise_fw(
                q, k, v, i, f, c_initial, n_initial, m_initial, return_last_states=return_last_states,
                chunk_size=chunk_size, eps=eps, save_states=need_bw, reverse=reverse, siging=siging)
            ctx.save_for_backward(q, k, v, i, f, c_initial, n_initial, m_initial, n_out, m_out, c_states)
            ctx.chunk_size, ctx.eps, ctx.reverse, ctx.siging = chunk_size, eps, reverse, siging
            if last is None:
                return h, None, None, None
            # native kernels hand states back in the input dtype (native/fw.py:53-62)
            return h, last[0].to(q.dtype), last[1].to(q.dtype), last[2].to(q.dtype)

        @staticmethod
        @custom_bwd(device_type="cuda")
        def backward(ctx, dh, dc_last, dn_last, dm_last):
            q, k, v, i, f, c0, n0, m0, n_out, m_out, c_states = ctx.saved_tensors
            dq, dk, dv, di, df, dc0 = mlstm_chunkwise_bw(
                q, k, v, i, f, n_out, m_out, dh, c0, n0, m0, dc_last=dc_last, chunk_size=ctx.chunk_size, eps=ctx.eps,
                want_dc_initial=c0 is not None, c_states=c_states, reverse=ctx.reverse, siging=ctx.siging)
            # dn_last / dm_last are ignored and dN/dM_initial are zeros, as in native/bw.py:329-337
            return (dq, dk, dv, di, df,
                    None if c0 is None else dc0.to(c0.dtype),
                    None if n0 is None else torch.zeros_like(n0),
                    None if m0 is None else torch.zeros_like(m0),
                    None, None, None, None, None)

    return _MlstmChunkwiseB200


_FUNCTIONS = {dt: _make_function(dt) for dt in (torch.float32, torch.float16, torch.bfloat16)}


def mlstm_chunkwise__b200(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    i: torch.Tensor,
    f: torch.Tensor,
    c_initial: torch.Tensor = None,
    n_initial: torch.Tensor = None,
    m_initial: torch.Tensor = None,
    return_last_states: bool = False,
    eps: float = 1e-6,
    chunk_size: int = 64,
    autocast_kernel_dtype: torch.dtype = torch.bfloat16,
    reverse: bool = False,
    **kwargs,
):
    """Drop-in for ``mlstm_chunkwise__native_custbw`` (native/fwbw.py:228-263).

    ``reverse=True`` (extension) runs the anti-causal scan: the result equals
    ``flip(f(flip(q), flip(k), ...))`` along S -- what ViLLayer obtains with two ``x.flip`` copies for
    its bottom-right direction (vision_lstm2.py:292-312) -- without moving any data.

    Returns h (B, NH, S, DHHV) or (h, (C_last, n_last, m_last)) when ``return_last_states``.
    Extra keyword arguments are ignored like ``native_autograd`` does (fwbw.py:204).
    """
    if autocast_kernel_dtype not in _FUNCTIONS:
        raise ValueError(f"Unsupported kernel dtype {autocast_kernel_dtype}.")
    fn = _FUNCTIONS[autocast_kernel_dtype]
    h, c_last, n_last, m_last = fn.apply(q, k, v, i, f, c_initial, n_initial, m_initial, bool(return_last_states),
                                         int(chunk_size), float(eps), bool(reverse), False)
    if return_last_states:
        return h, (c_last, n_last, m_last)
    return h


def mlstm_siging_chunkwise__b200(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    i: torch.Tensor,
    f: torch.Tensor,
    c_initial: torch.Tensor = None,
    n_initial: torch.Tensor = None,
    return_last_states: bool = False,
    eps: float = 1e-6,
    normalize: bool = True,
    chunk_size: int = 64,
    autocast_kernel_dtype: torch.dtype = torch.bfloat16,
    reverse: bool = False,
    **kwargs,
):
    """Sigmoid-input-gate variant: drop-in for ``mlstm_siging_chunkwise__xl_chunk``
    (mlstm_kernels/torch/chunkwise/triton_xl_chunk_siging/fwbw.py:211-268), the kernel the reference's
    CUDA model path selects (vision_lstm2.py:685-697).  No max state: last states are (C, n).
    The Triton tile-size kwargs are accepted and ignored.
    """
    if not normalize:
        raise NotImplementedError("normalize=False is not supported by the B200 siging kernel")
    if autocast_kernel_dtype not in _FUNCTIONS:
        raise ValueError(f"Unsupported kernel dtype {autocast_kernel_dtype}.")
    fn = _FUNCTIONS[autocast_kernel_dtype]
    need_states = c_initial is not None or n_initial is not None
    m_initial = None
    if need_states:
        B, NH = q.shape[:2]
        m_initial = torch.zeros(B, NH, 1, dtype=q.dtype, device=q.device)
        if c_initial is None:
            c_initial = torch.zeros(B, NH, q.shape[-1], v.shape[-1], dtype=q.dtype, device=q.device)
        if n_initial is None:
            n_initial = torch.zeros(B, NH, q.shape[-1], dtype=q.dtype, device=q.device)
    h, c_last, n_last, _ = fn.apply(q, k, v, i, f, c_initial, n_initial, m_initial, bool(return_last_states),
                                    int(chunk_size), float(eps), bool(reverse), True)
    if return_last_states:
        return h, (c_last, n_last)
    return h


def register(name: str = KERNEL_NAME) -> str:
    """Insert the kernels into the reference registry (mlstm_kernels/torch/chunkwise/__init__.py:9-15):
    ``<name>`` (exp input gate, max-state stabilised) and ``<name>_siging`` (sigmoid input gate).

    Returns the full kernel name usable as ``mLSTMBackendConfig(chunkwise_kernel=...)``.
    ``mlstm_kernels`` must be importable (it is the reference's package, not part of this repo).
    """
    from mlstm_kernels.torch.chunkwise import registry  # noqa: WPS433 (reference package)

    registry[name] = mlstm_chunkwise__b200
    registry[name + "_siging"] = mlstm_siging_chunkwise__b200
    return f"chunkwise--{name}"


def _cell_uses_siging(cell) -> bool:
    """True if the cell's current CUDA backend is a sigmoid-input-gate kernel (its ``chunkwise_kernel`` name
    carries ``siging``: ``chunkwise--triton_xl_chunk_siging`` in the reference, vision_lstm2.py:685-697)."""
    cfg = getattr(getattr(cell, "gpu_backend", None), "config", None)
    return "siging" in str(getattr(cfg, "chunkwise_kernel", ""))


def patch_model(model: torch.nn.Module, name: str = KERNEL_NAME, mode: str = "train_with_padding",
                siging: Optional[bool] = None, fused: bool = False, kernel_dtype: str = "bfloat16",
                keep_activations: bool = False) -> int:
    """Point ``gpu_backend`` of every MatrixLSTMCell (vision_lstm2.py:685-697) at the B200 kernel.

    ``siging=None`` (default) keeps the FUNCTION each cell computes on CUDA: a cell whose ``gpu_backend`` is a
    sigmoid-input-gate kernel -- the reference's CUDA default ``chunkwise--triton_xl_chunk_siging`` (sigmoid
    input gate, denominator max(|n|, 1)) -- gets ``chunkwise--b200_siging``, so released weights keep their
    meaning; any other cell gets the exp-gate / max-state kernel.  ``siging=True`` / ``False`` force one variant
    for every cell (``False`` = the function the reference runs on CPU, SURVEY.md finding 6; the parity oracle).
    ``fused=True`` additionally rebinds ``ViLLayer.mlstm_branch`` (vision_lstm2.py:292-312) to
    ``vil.mlstm_branch_b200``: same parameters and function, but the cell's output stage (MultiHeadLayerNorm +
    relayout + learnable skip) runs as one fused CUDA pass each way, the bottom-right direction uses the
    kernel's anti-causal scan instead of two ``x.flip`` copies, and q/k/v are consumed as strided views
    (SURVEY.md section 8(f) #2, #3).

    ``keep_activations=True`` raises ``ViLBlockPair.ckpt_thresh`` (vision_lstm2.py:1030, 1071-1078) so that the
    two S=6400 block pairs stop re-running their forward inside the backward: the reference checkpoints them to
    fit 40-80 GB parts; a B=32 base256 step peaks at ~22 GB of the B200's 180 GB with checkpointing on.
    Returns the number of cells patched.
    """
    from mlstm_kernels.torch.backend_module import mLSTMBackend, mLSTMBackendConfig

    base = register(name)
    n = n_sig = 0
    for mod in model.modules():
        if hasattr(mod, "gpu_backend") and hasattr(mod, "cpu_backend"):
            sig = _cell_uses_siging(mod) if siging is None else bool(siging)
            mod.gpu_backend = mLSTMBackend(mLSTMBackendConfig(
                chunkwise_kernel=base + ("_siging" if sig else ""), sequence_kernel="native_sequence__native",
                step_kernel="native", mode=mode, return_last_states=False, chunk_size=64, eps=1e-6,
                autocast_kernel_dtype="bfloat16"))
            n += 1
            n_sig += int(sig)
    if keep_activations:
        for mod in model.modules():
            if hasattr(mod, "ckpt_thresh"):
                mod.ckpt_thresh = 1 << 62
    if fused:
        from . import vil

        vil.patch_layers(model, siging=siging, kernel_dtype=kernel_dtype)
    return n
