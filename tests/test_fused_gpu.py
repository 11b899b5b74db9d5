"""GPU parity tests of the fused-normalisation paths (round 2): the halo-reuse 3x3 convolution, GroupNorm + SiLU
applied in its operand path from producer-side statistics, the statistics outputs of the staged epilogue, and
LayerNorm folded into the consumer GEMM.  References are plain PyTorch fp32 ops on the fp16-rounded inputs
(reference semantics: unet.py:470-489 ResnetBlock2D.forward, layer_norm.py:66-78, unet.py:499 Upsample2D)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, ref, atol, rtol, what):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    if bad.any():
        idx = torch.nonzero(bad)[0].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.4g} "
                             f"(ref absmax {ref.abs().max().item():.4g}) first at {idx}: got "
                             f"{got[tuple(idx)].item():.5g} ref {ref[tuple(idx)].item():.5g}")


def _rand(*shape, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale + shift).half()


def _pack(w):  # [Co, Ci, 3, 3] -> [Co, 9*Ci] with k = (ky*3+kx)*Ci + c
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _conv_ref(x, w, b=None):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None if b is None else b.float(), padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


def _chan_sums(x):  # NHWC fp16 -> [n, c, 2] (sum, sum of squares) in fp64 -> fp32
    xf = x.double().reshape(x.shape[0], -1, x.shape[-1])
    return torch.stack([xf.sum(1), (xf * xf).sum(1)], -1).float().contiguous()


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 128, 256),
                                         (2, 8, 8, 256, 128), (1, 5, 7, 64, 32), (3, 24, 24, 96, 64),
                                         (1, 64, 64, 64, 4), (1, 128, 128, 64, 64), (2, 96, 96, 128, 96),
                                         (1, 200, 136, 32, 32)])
def test_halo_conv_plain(cuda_lib, n, h, w, ci, co):
    x = _rand(n, h, w, ci, seed=1)
    wt = _rand(co, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    b = torch.randn(co, device="cuda")
    out = cuda_lib.conv3x3(x, _pack(wt), b, halo=True, out_dtype=torch.float32 if co == 4 else torch.float16)
    torch.cuda.synchronize()
    _close(out, _conv_ref(x, wt, b), 3e-3, 3e-3, f"halo conv {n}x{h}x{w} {ci}->{co}")


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 128, 256),
                                         (2, 8, 8, 256, 128), (1, 5, 7, 64, 32), (3, 24, 24, 96, 64),
                                         (1, 128, 128, 64, 64), (2, 96, 96, 128, 96), (1, 200, 136, 32, 32),
                                         (4, 64, 64, 64, 640)])
def test_halo_conv_tma_patches(cuda_lib, n, h, w, ci, co):
    """halo=2: the plain convolution whose patches arrive as one TMA box per 64-channel chunk (zero fill for the padding
    ring and the pad column) with the register epilogue; persistent CTAs with several tiles included (4x64x64 -> 640)."""
    x = _rand(n, h, w, ci, seed=1)
    wt = _rand(co, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    b = torch.randn(co, device="cuda")
    out = cuda_lib.conv3x3(x, _pack(wt), b, halo=2)
    torch.cuda.synchronize()
    _close(out, _conv_ref(x, wt, b), 3e-3, 3e-3, f"TMA halo conv {n}x{h}x{w} {ci}->{co}")
    assert torch.equal(out, cuda_lib.conv3x3(x, _pack(wt), b, halo=2))


def test_halo_conv_tma_two_sources_temb_residual(cuda_lib):
    n, h, w, c0, c1, co = 2, 32, 32, 640, 320, 640
    x0, x1 = _rand(n, h, w, c0, seed=1), _rand(n, h, w, c1, seed=2)
    wt = _rand(co, c0 + c1, 3, 3, scale=(9 * (c0 + c1)) ** -0.5, seed=3)
    temb = torch.randn(n, co + 64, device="cuda")  # strided per-image bias table
    res = _rand(n, h, w, co, seed=4)
    out = cuda_lib.conv3x3(x0, _pack(wt), temb[:, 32:], res, x1=x1, bias_rows=h * w, bias_stride=co + 64, halo=2)
    ref = _conv_ref(torch.cat([x0, x1], -1), wt) + temb[:, 32:32 + co].reshape(n, 1, 1, co) + res.float()
    _close(out, ref, 4e-3, 3e-3, "TMA halo conv two sources + temb + residual")


def test_halo_conv_two_sources_temb_residual(cuda_lib):
    n, h, w, c0, c1, co = 2, 32, 32, 640, 320, 640
    x0, x1 = _rand(n, h, w, c0, seed=1), _rand(n, h, w, c1, seed=2)
    wt = _rand(co, c0 + c1, 3, 3, scale=(9 * (c0 + c1)) ** -0.5, seed=3)
    temb = torch.randn(n, co + 64, device="cuda")  # strided per-image bias table
    res = _rand(n, h, w, co, seed=4)
    out = cuda_lib.conv3x3(x0, _pack(wt), temb[:, 32:], res, x1=x1, bias_rows=h * w, bias_stride=co + 64, halo=True)
    ref = _conv_ref(torch.cat([x0, x1], -1), wt) + temb[:, 32:32 + co].reshape(n, 1, 1, co) + res.float()
    _close(out, ref, 4e-3, 3e-3, "halo conv two sources + temb + residual")


@pytest.mark.parametrize("n,h,w,c0,c1,co,silu", [(2, 64, 64, 320, 0, 320, True), (2, 32, 32, 640, 320, 640, True),
                                                 (2, 16, 16, 1280, 640, 1280, True), (2, 8, 8, 1280, 0, 1280, True),
                                                 (1, 12, 20, 64, 32, 64, False), (2, 96, 96, 128, 64, 128, True),
                                                 (1, 256, 256, 128, 0, 128, True)])
def test_halo_conv_groupnorm_silu(cuda_lib, n, h, w, c0, c1, co, silu):
    """GroupNorm(32 groups over the concatenated channels) -> SiLU -> conv, statistics handed over as per-channel sums."""
    x0 = _rand(n, h, w, c0, seed=1, shift=0.3)
    x1 = _rand(n, h, w, c1, seed=2, scale=1.7) if c1 else None
    c = c0 + c1
    gamma = (1.0 + 0.2 * torch.randn(c, device="cuda")).contiguous()
    beta = (0.1 * torch.randn(c, device="cuda")).contiguous()
    wt = _rand(co, c, 3, 3, scale=(9 * c) ** -0.5, seed=3)
    b = torch.randn(co, device="cuda")
    xc = x0 if x1 is None else torch.cat([x0, x1], -1)
    gn = dict(chan0=_chan_sums(x0), chan1=None if x1 is None else _chan_sums(x1), gamma=gamma, beta=beta, groups=32,
              eps=1e-5, silu=silu)
    out = cuda_lib.conv3x3(x0, _pack(wt), b, x1=x1, halo=True, gn=gn)
    y = F.group_norm(xc.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    if silu:
        y = F.silu(y)
    ref = F.conv2d(y, wt.float(), b, padding=1).permute(0, 2, 3, 1)
    _close(out, ref, 6e-3, 4e-3, f"halo GN+SiLU conv {h}x{w} {c0}+{c1}->{co}")


def test_halo_upsample_conv(cuda_lib):
    n, h, w, c, co = 2, 16, 16, 256, 256
    x = _rand(n, h, w, c, seed=1)
    wt = _rand(co, c, 3, 3, scale=(9 * c) ** -0.5, seed=2)
    b = torch.randn(co, device="cuda")
    out = cuda_lib.conv3x3(x, _pack(wt), b, halo=True, upsample=True)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, wt.float(), b, padding=1).permute(0, 2, 3, 1)
    _close(out, ref, 3e-3, 3e-3, "halo upsample conv")


@pytest.mark.parametrize("halo", [False, True])
@pytest.mark.parametrize("n,h,w,ci,co", [(2, 64, 64, 64, 320), (2, 16, 16, 128, 640), (2, 8, 8, 128, 1280),
                                         (2, 96, 96, 64, 128)])
def test_conv_column_statistics(cuda_lib, halo, n, h, w, ci, co):
    x = _rand(n, h, w, ci, seed=1)
    wt = _rand(co, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    b = torch.randn(co, device="cuda")
    st = {}
    out = cuda_lib.conv3x3(x, _pack(wt), b, halo=halo, stats=st)
    out2 = cuda_lib.conv3x3(x, _pack(wt), b, halo=halo, stats={})  # tickets reset themselves: a second call works
    torch.cuda.synchronize()
    _close(out, _conv_ref(x, wt, b), 3e-3, 3e-3, "conv with statistics output")
    assert torch.equal(out, out2)
    ref = _chan_sums(out)
    _close(st["chan"], ref, 2e-2, 2e-4, f"column statistics halo={halo} {h}x{w}")


def test_linear_staged_residual_row_and_column_statistics(cuda_lib):
    for m, n, k, hw in [(8192, 320, 320, 4096), (2048, 640, 2560, 1024), (512, 1280, 1280, 256), (128, 1280, 1280, 64)]:
        x, w, r = _rand(m, k, seed=1), _rand(n, k, scale=k ** -0.5, seed=2), _rand(m, n, seed=3)
        b = torch.randn(n, device="cuda")
        st, rs = {}, {}
        out = cuda_lib.linear(x, w, b, r, stats=st, cs_hw=hw, rowstats=rs, static_w=True)
        torch.cuda.synchronize()
        _close(out, x.float() @ w.float().t() + b + r.float(), 4e-3, 2e-3, f"staged linear {m}x{n}x{k}")
        _close(st["chan"], _chan_sums(out.reshape(m // hw, hw, 1, n)), 2e-2, 2e-4, "column statistics (linear)")
        o = out.double()
        rows = rs["rows"].double().sum(0)
        _close(rows[:, 0], o.sum(1), 1e-2, 1e-4, "row sums")
        _close(rows[:, 1], (o * o).sum(1), 1e-2, 1e-4, "row sums of squares")


@pytest.mark.parametrize("geglu", [False, True])
def test_layernorm_folded_into_linear(cuda_lib, geglu):
    """producer GEMM leaves per-row sums; the consumer GEMM applies LayerNorm as a row scale + rank-1 correction."""
    m, c, n = 2048, 640, (5120 if geglu else 1920)
    x, wp, r = _rand(m, c, seed=1), _rand(c, c, scale=c ** -0.5, seed=2), _rand(m, c, seed=3, shift=0.5)
    rs = {}
    tok = cuda_lib.linear(x, wp, None, r, rowstats=rs, static_w=True)          # producer (residual epilogue)
    gamma = 1.0 + 0.2 * torch.randn(c, device="cuda")
    beta = 0.1 * torch.randn(c, device="cuda")
    w = _rand(n, c, scale=c ** -0.5, seed=4)
    b = torch.randn(n, device="cuda")
    wf = (w.float() * gamma[None, :]).half().contiguous()                    # gamma folded at pack time
    wg = wf.float().sum(1).contiguous()
    bf = (w.float() @ beta + b).contiguous()
    ln = dict(stat=rs["rows"], parts=rs["parts"], wg=wg, eps=1e-5)
    y = F.layer_norm(tok.float(), (c,), gamma, beta, 1e-5) @ w.float().t() + b
    if geglu:
        a, g = y.chunk(2, dim=1)
        ref = a * F.gelu(g)
        half = n // 2
        wi = torch.stack([wf[:half], wf[half:]], 1).reshape(n, c).contiguous()
        out = cuda_lib.linear(tok, wi, torch.stack([bf[:half], bf[half:]], 1).reshape(-1).contiguous(), geglu=True,
                              ln=dict(ln, wg=torch.stack([wg[:half], wg[half:]], 1).reshape(-1).contiguous()), static_w=True)
    else:
        ref = y
        out = cuda_lib.linear(tok, wf, bf, ln=ln, static_w=True)
    _close(out, ref, 1e-2, 6e-3, f"LayerNorm fold geglu={geglu}")


def test_halo_1x1_groupnorm_rowstats(cuda_lib):
    """SpatialTransformer entry: GroupNorm(eps 1e-6, no SiLU) -> proj_in 1x1 (unet.py:528-558) through the halo kernel's
    operand path, emitting the row statistics the first LayerNorm needs."""
    n, h, w, c = 2, 32, 32, 640
    x = _rand(n, h, w, c, seed=1, shift=-0.2)
    gamma = (1.0 + 0.2 * torch.randn(c, device="cuda")).contiguous()
    beta = (0.1 * torch.randn(c, device="cuda")).contiguous()
    wt = _rand(c, c, scale=c ** -0.5, seed=2)
    b = torch.randn(c, device="cuda")
    rs = {}
    gn = dict(chan0=_chan_sums(x), chan1=None, gamma=gamma, beta=beta, groups=32, eps=1e-6, silu=False)
    out = cuda_lib.conv3x3(x, wt, b, halo=True, gn=gn, taps=1, rowstats=rs)
    y = F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-6).permute(0, 2, 3, 1)
    ref = y @ wt.float().t() + b
    _close(out, ref, 5e-3, 3e-3, "halo 1x1 + GroupNorm")
    o = out.double().reshape(-1, c)
    _close(rs["rows"].double().sum(0)[:, 0], o.sum(1), 1e-2, 1e-4, "row sums (halo 1x1)")


@pytest.mark.parametrize("n,h,w,c0,c1,silu", [(2, 64, 64, 320, 0, True), (2, 16, 16, 1280, 640, True), (2, 8, 8, 1280, 0, False)])
def test_group_norm_apply_from_producer_statistics(cuda_lib, n, h, w, c0, c1, silu):
    x0 = _rand(n, h, w, c0, seed=1, shift=0.3)
    x1 = _rand(n, h, w, c1, seed=2, scale=1.7) if c1 else None
    c = c0 + c1
    gamma = (1.0 + 0.2 * torch.randn(c, device="cuda")).contiguous()
    beta = (0.1 * torch.randn(c, device="cuda")).contiguous()
    xc = x0 if x1 is None else torch.cat([x0, x1], -1)
    out = cuda_lib.group_norm_apply(x0, _chan_sums(x0), gamma, beta, 32, 1e-5, silu=silu, x1=x1,
                                    chan1=None if x1 is None else _chan_sums(x1))
    y = F.group_norm(xc.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    if silu:
        y = F.silu(y)
    _close(out, y.permute(0, 2, 3, 1), 3e-3, 3e-3, "group_norm_apply")


@pytest.mark.parametrize("n,h,w,c,cs0,cs1,co", [(2, 64, 64, 320, 640, 320, 320), (2, 32, 32, 640, 320, 0, 640),
                                                (2, 16, 16, 1280, 1280, 640, 1280), (2, 8, 8, 1280, 1280, 1280, 1280),
                                                (1, 24, 40, 64, 96, 32, 64), (4, 64, 64, 64, 72, 0, 640)])
def test_conv3x3_with_folded_shortcut(cuda_lib, n, h, w, c, cs0, cs1, co):
    """ResnetBlock2D's tail as ONE launch: conv2(h) + conv_shortcut(x ++ x1) (unet.py:483-489) -- the shortcut's 1x1
    matrix rides as extra centre-tap k-blocks of the 3x3 convolution (split-K plans and ragged channel chunks included)."""
    hh = _rand(n, h, w, c, seed=1)
    s0 = _rand(n, h, w, cs0, seed=2)
    s1 = _rand(n, h, w, cs1, seed=3) if cs1 else None
    wt = _rand(co, c, 3, 3, scale=(9 * c) ** -0.5, seed=4)
    ws = _rand(co, cs0 + cs1, scale=(cs0 + cs1) ** -0.5, seed=5)
    b = torch.randn(co, device="cuda")
    wcat = torch.cat([_pack(wt), ws], 1).contiguous()
    out = cuda_lib.conv3x3(hh, wcat, b, shortcut=(s0, s1))
    sc_in = s0 if s1 is None else torch.cat([s0, s1], -1)
    ref = _conv_ref(hh, wt, b) + sc_in.float() @ ws.float().t()
    _close(out, ref, 4e-3, 3e-3, f"conv + folded shortcut {n}x{h}x{w} {c}+{cs0}+{cs1}->{co}")
    assert torch.equal(out, cuda_lib.conv3x3(hh, wcat, b, shortcut=(s0, s1)))
