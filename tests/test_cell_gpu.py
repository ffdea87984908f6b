"""Parity of the cell's fused output stage and of the flip-free ViLLayer branch.

The checker is a plain PyTorch fp32 restatement of what the reference composes
(vision_lstm2.py:292-312, 701-753, 928-944): flips, F.group_norm over heads, transposes, skip add.
Tolerances: 1e-5 relative for fp32 tensors, 2e-2 where a 16-bit rounding is part of the function."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as G

    G.build()
    import xlstm_yolo_clean_b200 as p

    return p


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def ref_cell_out(h, weight, bias, skip, x, eps):
    """MultiHeadLayerNorm (vision_lstm2.py:928-944) + skip (vision_lstm2.py:306), float64."""
    B, NH, S, D = h.shape
    g = F.group_norm(h.double().transpose(1, 2).reshape(B * S, NH * D), NH,
                     None if weight is None else weight.double(), None if bias is None else bias.double(), eps)
    y = g.view(B, S, NH * D)
    if x is not None:
        y = y + skip.double() * x.double()
    return y


@pytest.mark.parametrize("NH,D", [(8, 64), (12, 32), (6, 128), (4, 64)])
@pytest.mark.parametrize("h_dtype,x_dtype", [(torch.bfloat16, torch.float16), (torch.float32, torch.float32),
                                             (torch.float16, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_cell_out_fw_bw(pkg, NH, D, h_dtype, x_dtype):
    torch.manual_seed(NH * 1000 + D)
    dev = torch.device("cuda:0")
    B, S = 3, 333  # rows not a multiple of anything the kernel tiles by
    H = NH * D
    h = (torch.randn(B, NH, S, D, device=dev) * 1.7 + 0.3).to(h_dtype).requires_grad_(True)
    x = torch.randn(B, S, H, device=dev).to(x_dtype).requires_grad_(True)
    w = (1.0 + 0.2 * torch.randn(H, device=dev)).requires_grad_(True)
    b = (0.1 * torch.randn(H, device=dev)).requires_grad_(True)
    sk = (1.0 + 0.1 * torch.randn(H, device=dev)).requires_grad_(True)
    dy = torch.randn(B, S, H, device=dev).to(x_dtype)
    y = pkg.cell_out(h, w, b, sk, x, eps=1e-6, out_dtype=x_dtype)
    assert y.shape == (B, S, H) and y.dtype == x_dtype
    y.backward(dy)
    got = dict(y=y, dh=h.grad, dx=x.grad, dw=w.grad, db=b.grad, dsk=sk.grad)
    leaves = [t.detach().double().requires_grad_(True) for t in (h, x, w, b, sk)]
    yr = ref_cell_out(leaves[0], leaves[2], leaves[3], leaves[4], leaves[1], 1e-6)
    yr.backward(dy.double())
    want = dict(y=yr, dh=leaves[0].grad, dx=leaves[1].grad, dw=leaves[2].grad, db=leaves[3].grad, dsk=leaves[4].grad)
    # outputs are rounded to their storage dtype once: half an ulp of the largest element
    tol_of = {torch.float32: 1e-5, torch.float16: 1e-3, torch.bfloat16: 8e-3}
    for k in got:
        tol = tol_of[h_dtype] if k == "dh" else tol_of[x_dtype] if k in ("y", "dx") else 1e-5
        assert rel(got[k], want[k]) < tol, (k, rel(got[k], want[k]), tol)


def test_cell_out_no_skip_strided_h(pkg):
    """h as a slice of a larger buffer (the padded-call case), no skip / bias."""
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    B, NH, S, D = 2, 8, 100, 64
    big = torch.randn(B, NH, S + 28, D, device=dev).to(torch.bfloat16)
    h = big[:, :, 28:]
    w = 1.0 + 0.2 * torch.randn(NH * D, device=dev)
    y = pkg.cell_out(h, w, None, None, None, eps=1e-6, out_dtype=torch.float32)
    assert rel(y, ref_cell_out(h, w, None, None, None, 1e-6)) < 1e-5


def test_cell_out_deterministic(pkg):
    torch.manual_seed(2)
    dev = torch.device("cuda:0")
    B, NH, S, D = 4, 8, 1600, 64
    h = torch.randn(B, NH, S, D, device=dev).to(torch.bfloat16).requires_grad_(True)
    x = torch.randn(B, S, NH * D, device=dev).to(torch.float16)
    w, sk = (torch.randn(NH * D, device=dev).requires_grad_(True) for _ in range(2))
    dy = torch.randn(B, S, NH * D, device=dev).to(torch.float16)
    outs = []
    for _ in range(2):
        h.grad = w.grad = sk.grad = None
        pkg.cell_out(h, w, None, sk, x, out_dtype=torch.float16).backward(dy)
        outs.append((h.grad.clone(), w.grad.clone(), sk.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)  # two-stage reduction in a fixed order: bit-identical parameter gradients


def test_cell_out_rejects_cpu_and_bad_shape(pkg):
    with pytest.raises(RuntimeError):
        pkg.cell_out(torch.randn(1, 4, 8, 64))
    with pytest.raises(RuntimeError):
        pkg.cell_out(torch.randn(1, 3, 8, 48, device="cuda:0"))  # D = 48 unsupported


# ---------------------------------------------------------------------------------------------
class _Norm(torch.nn.Module):
    def __init__(self, H):
        super().__init__()
        self.weight = torch.nn.Parameter(0.1 * torch.randn(H))
        self.bias = torch.nn.Parameter(0.1 * torch.randn(H))
        self.eps = 1e-6

    @property
    def weight_proxy(self):
        return 1.0 + self.weight


class _Conv(torch.nn.Conv2d):
    seqlens = None


class _Cell(torch.nn.Module):
    def __init__(self, H, NH):
        super().__init__()
        self.dim, self.num_heads, self.gate_soft_cap = H, NH, 15.0
        self.use_autocast, self.autocast_dtype = True, torch.float16
        self.ifgate = torch.nn.Linear(3 * H, 2 * NH)
        self.outnorm = _Norm(H)
        with torch.no_grad():
            self.ifgate.bias[:NH] = -2.0
            self.ifgate.bias[NH:] = torch.linspace(3.0, 6.0, NH)


class _Layer(torch.nn.Module):
    """Attribute-compatible stand-in for ViLLayer (vision_lstm2.py:218-290), test-local."""

    def __init__(self, dim, NH, direction):
        super().__init__()
        inner = 2 * dim
        self.direction = direction
        self.proj_up = torch.nn.Linear(dim, 2 * inner)
        self.conv = _Conv(inner, inner, 3, padding=1, groups=inner)
        self.qk_proj = torch.nn.Linear(inner, 2 * inner)
        self.v_proj = torch.nn.Linear(inner, inner)
        self.mlstm_cell = _Cell(inner, NH)
        self.learnable_skip = torch.nn.Parameter(1.0 + 0.1 * torch.randn(inner))
        self.proj_down = torch.nn.Linear(inner, dim)
        self.norm = torch.nn.RMSNorm(dim, eps=1e-6)
        with torch.no_grad():
            self.norm.weight.add_(0.1 * torch.randn(dim))


def _reference_branch(pkg, layer, x, reverse):
    """The composition the reference runs (flip, conv, cell with group_norm, skip, flip back), plain torch +
    the already parity-tested causal kernel."""
    if reverse:
        x = x.flip(dims=[1])
    B, S, _ = x.shape
    x_qk, x_v = torch.chunk(layer.proj_up(x), 2, dim=-1)
    hh = int(S ** 0.5)
    img = x_qk.reshape(B, hh, hh, -1).permute(0, 3, 1, 2)
    x_act = F.silu(layer.conv(img).permute(0, 2, 3, 1).reshape(B, S, -1))
    q, k = torch.chunk(layer.qk_proj(x_act), 2, dim=-1)
    v = layer.v_proj(x_v)
    cell = layer.mlstm_cell
    NH = cell.num_heads
    pre = cell.ifgate(torch.cat([q, k, v], -1))
    pre = 15.0 * torch.tanh(pre / 15.0)
    i, f = (t.transpose(-1, -2) for t in torch.chunk(pre, 2, -1))
    qh, kh, vh = (t.view(B, S, NH, -1).transpose(1, 2).to(torch.float16) for t in (q, k, v))
    i, f = i.to(torch.float16), f.to(torch.float16)
    pad = (-S) % 64
    if pad:
        qh, kh, vh = (F.pad(t, (0, 0, 0, pad)) for t in (qh, kh, vh))
        i, f = F.pad(i, (0, pad)), F.pad(f, (0, pad))
    h = pkg.mlstm_chunkwise__b200(q=qh, k=kh, v=vh, i=i, f=f, chunk_size=64, eps=1e-6)[:, :, :S].to(x.dtype)
    D = h.shape[-1]
    g = F.group_norm(h.transpose(1, 2).reshape(B * S, NH * D), NH, cell.outnorm.weight_proxy, cell.outnorm.bias, 1e-6)
    y = g.view(B, S, NH * D) + layer.learnable_skip * x_act
    out = layer.proj_down(y)
    return out.flip(dims=[1]) if reverse else out


class _Dir:
    def __init__(self, name):
        self.name = name


@pytest.mark.parametrize("S", [256, 100])  # 100 -> padded to 128 like the model's smallest stage
@pytest.mark.parametrize("reverse", [False, True])
def test_flip_free_branch_matches_reference_composition(pkg, S, reverse):
    torch.manual_seed(5)
    dev = torch.device("cuda:0")
    dim, NH, B = 128, 4, 2  # inner 256, D = 64
    layer = _Layer(dim, NH, _Dir("ROWWISE_FROM_BOT_RIGHT" if reverse else "ROWWISE_FROM_TOP_LEFT")).to(dev)
    x = torch.randn(B, S, dim, device=dev)
    dout = torch.randn(B, S, dim, device=dev)

    def run(fn):
        layer.zero_grad()
        xi = x.clone().requires_grad_(True)
        out = fn(xi)
        out.backward(dout)
        grads = {n: p.grad.clone() for n, p in layer.named_parameters() if p.grad is not None}  # (norm is not on the branch)
        return out.detach(), xi.grad.clone(), grads

    o_ref, dx_ref, g_ref = run(lambda xi: _reference_branch(pkg, layer, xi, reverse))
    o_new, dx_new, g_new = run(lambda xi: pkg.mlstm_branch_b200(layer, xi))
    # both sides run the same fp16 kernel arithmetic; what differs is summation order and one 16-bit rounding
    assert rel(o_new, o_ref) < 5e-3, rel(o_new, o_ref)
    assert rel(dx_new, dx_ref) < 2e-2, rel(dx_new, dx_ref)
    for n in g_ref:
        assert rel(g_new[n], g_ref[n]) < 2e-2, (n, rel(g_new[n], g_ref[n]))


def test_patch_layers_rebinds_and_trains_under_amp(pkg):
    """patch_layers on a two-direction stack: the rebound branch is what runs, under fp16 autocast + GradScaler
    like the ultralytics trainer (engine/trainer.py:382-392), gradients reach every parameter and match the
    un-rebound composition."""
    torch.manual_seed(9)
    dev = torch.device("cuda:0")

    class Stack(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = _Layer(128, 4, _Dir("ROWWISE_FROM_TOP_LEFT"))
            self.b = _Layer(128, 4, _Dir("ROWWISE_FROM_BOT_RIGHT"))
            for l in (self.a, self.b):
                l.mlstm_branch = lambda x, _l=l: _reference_branch(pkg, _l, x, pkg.vil._is_reverse(_l))

        def forward(self, x):
            x = x + self.a.mlstm_branch(self.a.norm(x))
            return x + self.b.mlstm_branch(self.b.norm(x))

    model = Stack().to(dev)
    x = torch.randn(2, 400, 128, device=dev)

    def step():
        model.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            out = model(x)
            loss = out.float().pow(2).mean()
        loss.backward()
        return float(loss), {n: p.grad.clone() for n, p in model.named_parameters()}

    loss_ref, g_ref = step()
    assert pkg.patch_layers(model) == 2
    loss_new, g_new = step()
    assert abs(loss_new - loss_ref) < 5e-3 * abs(loss_ref)
    assert set(g_new) == set(g_ref)
    for n in g_ref:
        assert rel(g_new[n], g_ref[n]) < 3e-2, (n, rel(g_new[n], g_ref[n]))


@pytest.mark.parametrize("C", [192, 256, 384, 512])
@pytest.mark.parametrize("x_dtype,y_dtype", [(torch.float16, torch.float16), (torch.float32, torch.float16),
                                             (torch.float32, torch.float32), (torch.bfloat16, torch.float32)])
def test_rms_norm_fw_bw(pkg, C, x_dtype, y_dtype):
    """Fused RMSNorm (ViLLayer.norm / .ffn_norm, vision_lstm2.py:277-278) against torch.rms_norm itself -- the
    composite path it replaces under autocast -- and against float64 autograd for the gradients."""
    torch.manual_seed(C)
    dev = torch.device("cuda:0")
    x = (torch.randn(5, 77, C, device=dev) * 1.3 + 0.2).to(x_dtype).requires_grad_(True)
    w = (1.0 + 0.2 * torch.randn(C, device=dev)).requires_grad_(True)
    dy = torch.randn(5, 77, C, device=dev).to(y_dtype)
    y = pkg.rms_norm_b200(x, w, 1e-6, out_dtype=y_dtype)
    assert y.shape == x.shape and y.dtype == y_dtype
    y.backward(dy)
    want = torch.rms_norm(x.detach(), (C,), w.detach(), 1e-6)  # fp32 result of the composite / fused torch path
    tol_y = {torch.float32: 2e-6, torch.float16: 1e-3, torch.bfloat16: 8e-3}
    assert rel(y, want) < max(tol_y[y_dtype], tol_y[x_dtype]), rel(y, want)
    x64, w64 = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    (x64 * torch.rsqrt(x64.pow(2).mean(-1, keepdim=True) + 1e-6) * w64).backward(dy.double())
    assert rel(x.grad, x64.grad) < max(tol_y[x_dtype], 2e-3), rel(x.grad, x64.grad)
    assert rel(w.grad, w64.grad) < max(tol_y[x_dtype], 2e-3), rel(w.grad, w64.grad)


def test_rms_norm_autocast_output_feeds_linear_identically(pkg):
    """Under fp16 autocast the consumers are Linear layers: what they read from the fused norm must equal what
    they read from torch's composite (y = half((x rstd) w), one rounding) -- same arithmetic, so the only
    freedom is the summation order of mean(x^2): at most one fp16 ulp, on a small fraction of the elements."""
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    norm = torch.nn.RMSNorm(256, eps=1e-6).to(dev)
    with torch.no_grad():
        norm.weight.add_(0.3 * torch.randn(256, device=dev))
    x = torch.randn(4, 400, 256, device=dev).half()
    with torch.autocast("cuda", dtype=torch.float16):
        a = norm(x).to(torch.float16)
        b = pkg.rms_norm_b200(x, norm.weight, norm.eps)
    assert b.dtype == torch.float16
    diff = (a.float() - b.float()).abs()
    assert float((diff / a.float().abs().clamp_min(1e-3)).max()) <= 2.0 ** -9  # one ulp of fp16
    assert float((diff > 0).float().mean()) < 0.02


def _layer_golden_runner(pkg, name, one_launch="auto"):
    import os

    import numpy as np

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"vil_layer_{name}.npz"))
    dim, NH, side, B = (int(v) for v in z["meta"])
    dev = torch.device("cuda:0")
    layer = _Layer(dim, NH, _Dir("ROWWISE_FROM_BOT_RIGHT" if name.endswith("rev") else "ROWWISE_FROM_TOP_LEFT"))
    layer.conv.seqlens = [side, side]
    sd = {k[2:]: torch.from_numpy(z[k]).float() for k in z.files if k.startswith("p_")}
    missing = layer.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and set(missing.missing_keys) <= {"norm.weight"}
    layer = layer.to(dev)
    gy, gdx = torch.from_numpy(z["y"]), torch.from_numpy(z["dx"])
    dout = torch.from_numpy(z["dout"]).float().to(dev)

    def run(use_fp16_cast):
        layer.mlstm_cell.use_autocast = use_fp16_cast
        x = torch.from_numpy(z["x"]).float().to(dev).requires_grad_(True)
        y = pkg.mlstm_branch_b200(layer, x, one_launch=one_launch)
        (dx,) = torch.autograd.grad(y, x, dout)
        with torch.no_grad():  # the inference composition (one launch per cell under "auto") computes the same function
            y_ng = pkg.mlstm_branch_b200(layer, x.detach(), one_launch=one_launch)
        assert rel(y_ng.cpu(), y.detach().cpu()) < (5e-3 if use_fp16_cast else 1e-5)
        return rel(y.detach().cpu(), gy), rel(dx.cpu(), gdx)

    return run


@pytest.mark.parametrize("tag", ["fwd", "rev"])
def test_branch_matches_reference_vil_layer_golden(pkg, tag):
    """mlstm_branch_b200 against the UNMODIFIED reference ViLLayer.mlstm_branch (vision_lstm2.py:292-312), run on CPU
    in float64 by tests/golden/make_golden_vil.py for both scan directions at S = 100 (the padded stage): same
    parameters, same input -> same output and same input gradient, although this side has no flips, an
    anti-causal kernel, a rotated conv, no padding, layer-layout gradient writes and a fused cell output.

    The tight comparison runs the kernels in fp32 (exact family; ``use_autocast=False`` skips the fp16 cast
    MatrixLSTMCell applies on CUDA, vision_lstm2.py:730-745).  With that cast -- the reference's own CUDA rule --
    the 16-bit rounding of h is amplified by the LayerNorm that follows (h has a small per-head variance here), in
    the fused branch and in the plain torch composition alike (measured: identical 1.2e-2 / 1.9e-1 on this vector),
    so for fp16 only the forward is held to the 16-bit bar."""
    run = _layer_golden_runner(pkg, tag)
    ey, ex = run(False)
    assert ey < 1e-4 and ex < 1e-3, (ey, ex)
    ey16, _ = run(True)
    assert ey16 < 2e-2, ey16


@pytest.mark.parametrize("one_launch", ["auto", "always", "never"])
@pytest.mark.parametrize("tag", ["fwd", "rev"])
def test_branch_matches_well_conditioned_layer_golden_in_fp16(pkg, tag, one_launch):
    """Same comparison on vectors where the 16-bit rounding is NOT amplified (tests/golden/make_golden_vil.py, the
    ``wc_`` pair: q/k/v of order 1-10, gates spread like a trained model's, S = 144 = one full 128-token tile + a ragged one): here the
    reference's own CUDA rule -- q/k/v/i/f cast to fp16, vision_lstm2.py:730-745 -- runs the tcgen05 kernels (causal
    and anti-causal, ragged last tile, fused LayerNorm epilogue) and BOTH the output and the input gradient of the
    unmodified float64 reference layer are held to the 16-bit bar.  Rounding q/k/v/h to fp16 inside the reference
    itself moves y and dx by 7e-4 on these vectors."""
    run = _layer_golden_runner(pkg, "wc_" + tag, one_launch)  # training: "always" = fused epilogue, else two launches
    ey, ex = run(False)
    assert ey < 1e-4 and ex < 1e-3, (ey, ex)
    ey16, ex16 = run(True)
    assert ey16 < 5e-3 and ex16 < 5e-3, (ey16, ex16)  # measured 8e-4 / 7e-4 (both directions)


@pytest.mark.parametrize("kdt,odt", [(torch.bfloat16, torch.float16), (torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16)],
                         ids=["bf16_kernel_fp16_out", "fp16", "bf16"])
@pytest.mark.parametrize("NH,D,S", [(4, 64, 384), (12, 32, 200), (3, 128, 256), (8, 64, 100)])
@pytest.mark.parametrize("reverse", [False, True], ids=["causal", "anticausal"])
def test_fused_forward_epilogue_equals_kernel_plus_cell_out(pkg, NH, D, S, reverse, kdt, odt):
    """mlstm_b200_fw_epilogue: MultiHeadLayerNorm + relayout + learnable skip inside the
    forward kernel's epilogue.  The un-normalised h it can still write is bit-identical to the plain forward's, and y equals
    the stand-alone cell-output kernel applied to that h (itself pinned to a float64 group_norm above) up to one rounding
    of the output dtype; without want_h no h is written at all."""
    from xlstm_yolo_clean_b200 import backend

    B, H = 2, NH * D
    g = torch.Generator().manual_seed(S + D)
    dev = torch.device("cuda:0")
    qk = torch.randn(B, S, 2 * H, generator=g).to(kdt).to(dev)
    v = torch.randn(B, S, H, generator=g).to(kdt).to(dev)
    gates = torch.cat([torch.randn(B, S, NH, generator=g), 3 + torch.randn(B, S, NH, generator=g)], -1).to(kdt).to(dev)
    x = torch.randn(B, S, H, generator=g).to(odt).to(dev)
    w = (1 + 0.1 * torch.randn(H, generator=g)).to(dev)
    b = (0.1 * torch.randn(H, generator=g)).to(dev)
    sk = torch.randn(H, generator=g).to(dev)
    q, k, vv, i, f = pkg.vil._heads(qk, v, gates, NH)
    L = 4
    h0 = pkg.mlstm_chunkwise_fw(q, k, vv, i, f, chunk_size=L, reverse=reverse)[0]
    heads = lambda t: t.view(B, S, NH, D).transpose(1, 2)  # noqa: E731
    y = torch.empty(B, S, H, dtype=odt, device=dev)
    epi = backend.FwEpilogue(heads(y), heads(x), w, b, sk, 1e-5, True)
    h1 = backend._fw_launch(q, k, vv, i, f, None, None, None, None, False, L, 1e-6, None, True, reverse, False, 0.0, epi)[0]
    torch.cuda.synchronize()
    assert torch.equal(h1, h0)
    want = pkg.cell_out(h0, w, b, sk, x, eps=1e-5, out_dtype=odt)
    assert rel(y, want) < (2e-3 if odt == torch.float16 else 1.2e-2)
    # inference flavour: no skip input, no bias, h not written
    y2 = torch.empty(B, S, H, dtype=odt, device=dev)
    epi2 = backend.FwEpilogue(heads(y2), None, w, None, None, 1e-5, False)
    h2 = backend._fw_launch(q, k, vv, i, f, None, None, None, None, False, L, 1e-6, None, False, reverse, False, 0.0, epi2)[0]
    torch.cuda.synchronize()
    assert h2 is None
    assert rel(y2, pkg.cell_out(h0, w, None, None, None, eps=1e-5, out_dtype=odt)) < (2e-3 if odt == torch.float16 else 1.2e-2)


@pytest.mark.parametrize("src,dst", [(torch.float16, torch.bfloat16), (torch.bfloat16, torch.float16)], ids=["fp16_to_bf16", "bf16_to_fp16"])
def test_convert16_is_bit_identical_to_torch(pkg, src, dst):
    """mlstm_b200_convert16 = Tensor.to() for the autocast re-rounding of q / k / v (native/fwbw.py:37): every size
    class of the vector loop, a misaligned view, a dense permuted tensor, non-finite and subnormal values."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    for n in (1, 7, 8, 9, 8 * 256 * 4 - 1, 1_000_003, 32 * 1600 * 512 + 24):
        x = (torch.randn(n, generator=g) * 50).to(src).to(dev)
        y = pkg.convert16(x, dst)
        assert y.dtype == dst and torch.equal(y.view(torch.int16), x.to(dst).view(torch.int16)), n
    special = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), 65504.0, -65504.0, 6e-8, 1e-7, 3e38, -3e38,
                            1.0009765625, 0.99951171875, 5.9e-8, 1e-40] * 3, dtype=torch.float32).to(src).to(dev)
    a, b = pkg.convert16(special, dst), special.to(dst)
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    assert torch.equal(a.nan_to_num(7.0).view(torch.int16), b.nan_to_num(7.0).view(torch.int16))
    base = (torch.randn(4099, generator=g)).to(src).to(dev)
    off = base[3:]  # starts 6 bytes into a vector
    assert torch.equal(pkg.convert16(off, dst), off.to(dst))
    perm = (torch.randn(6, 40, 24, generator=g)).to(src).to(dev).permute(1, 0, 2)  # dense, not contiguous: strides are kept
    yp = pkg.convert16(perm, dst)
    assert yp.stride() == perm.stride() and torch.equal(yp, perm.to(dst))
    sliced = perm[:, :, :8]  # not dense: torch's path
    assert torch.equal(pkg.convert16(sliced, dst), sliced.to(dst))
    assert pkg.convert16(base, src) is base


@pytest.mark.parametrize("S,reverse", [(400, False), (100, True)])
def test_graphed_branch_is_bit_identical_to_eager_over_optimizer_steps(pkg, S, reverse):
    """patch_layers(graphs=True): the branch's forward / backward replay as CUDA graphs (vil._GraphedBranch).  Same
    kernels on the live parameters: losses, input gradients and the parameters after every SGD step equal the eager
    branch's bit for bit -- in particular the fp16 weight copies of the autocast rule are re-made on every replay."""
    import copy

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    side = int(S ** 0.5)
    eager = _Layer(128, 4, _Dir("ROWWISE_FROM_BOT_RIGHT" if reverse else "ROWWISE_FROM_TOP_LEFT")).to(dev)
    eager.conv.seqlens = [side, side]
    graphed = copy.deepcopy(eager)
    graphed.conv.seqlens = [side, side]
    assert pkg.patch_layers(eager) == 1 and pkg.patch_layers(graphed, graphs=True) == 1
    opts = [torch.optim.SGD(m.parameters(), lr=0.05) for m in (eager, graphed)]
    for it in range(4):
        x = torch.randn(8, S, 128, device=dev).half()
        res = []
        for m, opt in zip((eager, graphed), opts):
            xi = x.clone().requires_grad_(True)
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.float16):
                y = m.mlstm_branch(xi)
            loss = (y.float() ** 2).mean()
            loss.backward()
            opt.step()
            res.append((loss.detach(), xi.grad))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]), it
        for a, b in zip(eager.parameters(), graphed.parameters()):
            assert torch.equal(a, b), it
    assert len(graphed.mlstm_branch._graphs) == 1
    for it in range(3):  # no gradients: a forward-only graph, same function, live parameters
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            x = torch.randn(8, S, 128, device=dev).half()
            a, b = graphed.mlstm_branch(x), eager.mlstm_branch(x)
            assert torch.equal(a, b), it
            assert a.data_ptr() != graphed.mlstm_branch(x).data_ptr()  # results are copies, not the static buffer
        with torch.no_grad():
            for pa, pb in zip(eager.parameters(), graphed.parameters()):
                d = 0.01 * torch.randn_like(pa)
                pa.add_(d)
                pb.add_(d)
    assert len(graphed.mlstm_branch._graphs) == 2
    clone = copy.deepcopy(graphed)  # graphs do not travel with copies; the copy builds its own
    assert clone.mlstm_branch._graphs == {} and clone.mlstm_branch.layer is clone


class _NoDrop(torch.nn.Module):
    drop_prob = 0.0


class _FullLayer(_Layer):
    """_Layer plus what ViLLayer.forward wraps around the branch (vision_lstm2.py:331-341)."""

    def __init__(self, dim, NH, direction):
        super().__init__(dim, NH, direction)
        self.ffn_norm = torch.nn.RMSNorm(dim, eps=1e-6)
        self.ffn = torch.nn.Sequential(torch.nn.Linear(dim, 2 * dim), torch.nn.GELU(), torch.nn.Linear(2 * dim, dim))
        self.drop_path = _NoDrop()

    def mlstm_branch(self, x):  # replaced by patch_layers
        raise AssertionError("unpatched")

    def forward(self, x):
        x = x + self.mlstm_branch(self.norm(x))
        return x + self.ffn(self.ffn_norm(x))


def test_graphed_whole_layer_is_bit_identical_to_eager_over_optimizer_steps(pkg):
    """patch_layers(graphs=True) on a full ViLLayer-shaped module: norm -> fused branch -> residual -> ffn_norm -> ffn ->
    residual replay as one forward and one backward CUDA graph (vil._GraphedLayer); bit-identical training."""
    import copy

    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    eager = _FullLayer(256, 8, _Dir("ROWWISE_FROM_BOT_RIGHT")).to(dev)
    eager.conv.seqlens = [20, 20]
    graphed = copy.deepcopy(eager)
    graphed.conv.seqlens = [20, 20]
    assert pkg.patch_layers(eager) == 1 and pkg.patch_layers(graphed, graphs=True) == 1
    assert "forward" in graphed.__dict__ and "forward" not in eager.__dict__
    opts = [torch.optim.SGD(m.parameters(), lr=0.05) for m in (eager, graphed)]
    for it in range(3):
        x = torch.randn(4, 400, 256, device=dev)
        res = []
        for m, opt in zip((eager, graphed), opts):
            xi = x.clone().requires_grad_(True)
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.float16):
                y = m(xi)
            loss = (y.float() ** 2).mean()
            loss.backward()
            opt.step()
            res.append((loss.detach(), xi.grad))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]), it
        for a, b in zip(eager.parameters(), graphed.parameters()):
            assert torch.equal(a, b), it
    assert len(graphed.forward._graphs) == 1
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):  # inference: forward-only graph of the whole layer
        graphed.eval(), eager.eval()
        xe = torch.randn(1, 400, 256, device=dev)
        assert torch.equal(graphed(xe), eager(xe)) and torch.equal(graphed(xe), eager(xe))
        graphed.train(), eager.train()
    assert len(graphed.forward._graphs) == 2
    graphed.drop_path.drop_prob = 0.1  # stochastic depth selects samples by value: eager
    with torch.autocast("cuda", dtype=torch.float16):
        graphed(x.clone().requires_grad_(True))
    assert len(graphed.forward._graphs) == 2


@pytest.mark.parametrize("reentrant", [True, False])
def test_graphed_layer_inside_a_checkpointed_block(pkg, reentrant):
    """The reference checkpoints its S = 6400 block pairs (vision_lstm2.py:1071-1078): the first pass runs without
    gradients (forward-only graph), the re-computation runs inside the backward pass (never builds a graph there).
    Same losses, gradients and parameters as the eager layer."""
    import copy

    from torch.utils.checkpoint import checkpoint

    dev = torch.device("cuda:0")
    torch.manual_seed(6)
    eager = _FullLayer(256, 8, _Dir("ROWWISE_FROM_TOP_LEFT")).to(dev)
    eager.conv.seqlens = [10, 10]
    graphed = copy.deepcopy(eager)
    graphed.conv.seqlens = [10, 10]
    assert pkg.patch_layers(eager) == 1 and pkg.patch_layers(graphed, graphs=True) == 1
    opts = [torch.optim.SGD(m.parameters(), lr=0.05) for m in (eager, graphed)]
    for it in range(3):
        x = torch.randn(4, 100, 256, device=dev)
        res = []
        for m, opt in zip((eager, graphed), opts):
            xi = x.clone().requires_grad_(True)
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.float16):
                y = checkpoint(m, xi, use_reentrant=reentrant)
            loss = (y.float() ** 2).mean()
            loss.backward()
            opt.step()
            res.append((loss.detach(), xi.grad))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]), it
        for a, b in zip(eager.parameters(), graphed.parameters()):
            assert torch.equal(a, b), it
