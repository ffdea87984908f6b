"""Out-of-bounds writes: outputs are carved out of buffers with sentinel-filled guard bands on both sides and the
guards are checked after the kernels ran (compute-sanitizer is not available on the GPU pool).  Covers the kernels
that take caller-provided output tensors: the mLSTM backward (strided outputs, d = 32 / 64 natively, d = 128 through
the four-block route with its scratch accumulation) and the cell output stage through the raw C-ABI."""
import ctypes as C

import pytest
import torch

from oracle import mlstm_oracle as O

pytestmark = pytest.mark.gpu
GUARD = 4096  # elements on each side
SENT = -12345.0


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as G

    G.build()
    import xlstm_yolo_clean_b200 as p

    return p


def guarded(shape, dtype, dev):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * GUARD,), SENT, dtype=dtype, device=dev)
    return buf, buf[GUARD:GUARD + n].view(*shape)


def intact(buf):
    return bool((buf[:GUARD] == SENT).all()) and bool((buf[-GUARD:] == SENT).all())


@pytest.mark.parametrize("D,S", [(64, 100), (32, 52), (128, 324), (64, 1600), (128, 128)])
@pytest.mark.parametrize("reverse", [False, True], ids=["causal", "anticausal"])
def test_backward_outputs_stay_inside_their_tensors(pkg, D, S, reverse):
    dev = torch.device("cuda:0")
    B, NH = 2, 3
    inp = O.make_inputs(B, NH, S, D, D, seed=D + S, dtype=torch.float32)
    t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
    h, n_out, m_out, _, cst = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"], chunk_size=4, reverse=reverse)
    bufs, outs = [], []
    for shape in ((B, NH, S, D),) * 3 + ((B, NH, S),) * 2:
        b, v = guarded(shape, torch.bfloat16, dev)
        bufs.append(b)
        outs.append(v)
    ref = pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"], chunk_size=4, c_states=cst,
                                 reverse=reverse)
    pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], n_out, m_out, t["dh"], chunk_size=4, c_states=cst,
                           reverse=reverse, out=tuple(outs))
    torch.cuda.synchronize()
    for b in bufs:
        assert intact(b)
    for a, b in zip(outs, ref[:5]):
        assert torch.equal(a, b)  # same kernels, same inputs: bit-identical, wherever the outputs live


@pytest.mark.parametrize("NH,D", [(8, 64), (12, 32), (6, 128)])
def test_cell_output_stage_stays_inside_its_tensors(pkg, NH, D):
    from xlstm_yolo_clean_b200 import _cabi
    from xlstm_yolo_clean_b200.backend import _tensor

    lib = pkg.load_library()
    dev = torch.device("cuda:0")
    B, S, H = 3, 333, NH * D
    h = torch.randn(B, NH, S, D, device=dev).to(torch.bfloat16)
    x = torch.randn(B, S, H, device=dev).to(torch.float16)
    dy = torch.randn(B, S, H, device=dev).to(torch.float16)
    w, bias, sk = (torch.randn(H, device=dev) for _ in range(3))
    by, y = guarded((B, S, H), torch.float16, dev)
    bdh, dh = guarded((B, NH, S, D), torch.bfloat16, dev)
    bdx, dx = guarded((B, S, H), torch.float16, dev)
    bpar, dpar = guarded((3, H), torch.float32, dev)
    a = _cabi.CellOutArgs()
    a.B, a.NH, a.S, a.D = B, NH, S, D
    a.h_dtype, a.x_dtype, a.y_dtype = _cabi.BF16, _cabi.F16, _cabi.F16
    a.eps = 1e-6
    a.h, a.x, a.y = _tensor(h), _tensor(x), _tensor(y)
    a.weight, a.bias, a.skip = w.data_ptr(), bias.data_ptr(), sk.data_ptr()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    assert lib.mlstm_b200_cellout_fw(C.byref(a), st) == 0
    b = _cabi.CellOutBwArgs()
    b.fw = a
    b.dy, b.dh, b.dx = _tensor(dy), _tensor(dh), _tensor(dx)
    b.dweight, b.dbias, b.dskip = dpar[0].data_ptr(), dpar[1].data_ptr(), dpar[2].data_ptr()
    nws = lib.mlstm_b200_cellout_workspace_bytes(C.byref(a))
    bws = torch.full((nws + 2 * GUARD,), 0x5A, dtype=torch.uint8, device=dev)
    b.workspace, b.workspace_bytes = bws[GUARD:].data_ptr(), nws
    assert lib.mlstm_b200_cellout_bw(C.byref(b), st) == 0
    torch.cuda.synchronize()
    for buf in (by, bdh, bdx, bpar):
        assert intact(buf)
    assert bool((bws[:GUARD] == 0x5A).all()) and bool((bws[-GUARD:] == 0x5A).all())
    assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(dh.float()).all()) and bool(torch.isfinite(dpar).all())
