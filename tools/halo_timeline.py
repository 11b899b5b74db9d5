"""Timeline of one CTA of the halo convolution kernel (clock64 at fixed points; run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from b200sd import lib as L  # noqa: E402

L.load()
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*s, scale=1.0):
    return (torch.randn(*s, generator=g, device="cuda") * scale).half()


def chan(x):
    xf = x.float().reshape(x.shape[0], -1, x.shape[-1])
    return torch.stack([xf.sum(1), (xf * xf).sum(1)], -1).contiguous()


def run(name, n, h, w, ci, co, gn=True, taps=9, residual=False):
    x = rnd(n, h, w, ci)
    wt = rnd(co, taps * ci, scale=(taps * ci) ** -0.5)
    b = torch.randn(co, device="cuda")
    res = rnd(n, h, w, co) if residual else None
    gnd = dict(chan0=chan(x), chan1=None, gamma=torch.ones(ci, device="cuda"), beta=torch.zeros(ci, device="cuda"), groups=32,
               eps=1e-5, silu=True) if gn else None
    for it in range(3):
        if it == 2:
            os.environ["B200SD_DBG_PTR"] = hex(dbg.data_ptr())
        st = {}
        L.conv3x3(x, wt, b, res, halo=True, gn=gnd, stats=st, taps=taps)
        torch.cuda.synchronize()
    os.environ.pop("B200SD_DBG_PTR", None)
    d = dbg.cpu().tolist()
    t0 = d[0]
    rel = lambda v: (v - t0) if v else None
    kc = (ci + 63) // 64
    print(f"== {name}: {n}x{h}x{w} {ci}->{co} taps={taps} chunks={kc}")
    print("  setup->pdl_wait", rel(d[1]), "first loads issued", rel(d[2]), "table built", rel(d[3]))
    for j in range(min(kc, 24)):
        print(f"  chunk {j:2d}: loader got buffer {rel(d[8 + 2 * j])}, patch ready {rel(d[9 + 2 * j])} | mma saw patch {rel(d[64 + 2 * j])}, "
              f"issued last tap {rel(d[65 + 2 * j])}")
    print("  accumulator ready", rel(d[4]), "epilogue done", rel(d[5]))
    dbg.zero_()


run("conv 64x64 GN", 2, 64, 64, 320, 320)
run("conv 64x64 GN + residual", 2, 64, 64, 320, 320, residual=True)
run("1x1 64x64 GN", 2, 64, 64, 320, 320, taps=1)
run("conv 8x8 GN", 2, 8, 8, 1280, 1280)
run("conv 16x16 plain", 2, 16, 16, 1280, 1280, gn=False)


# ---- A/B timing: halo conv with fused GroupNorm vs round-1 path (cluster GroupNorm + 9-tap TMA conv) -----------------
def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / reps


print("\n== A/B per-call time (us), CUDA graph of 20 back-to-back calls")
for name, n, h, w, ci, co in [("64x64 320->320", 2, 64, 64, 320, 320), ("64x64 960->320", 2, 64, 64, 960, 320),
                              ("32x32 640->640", 2, 32, 32, 640, 640), ("32x32 1920->640", 2, 32, 32, 1920, 640),
                              ("16x16 1280->1280", 2, 16, 16, 1280, 1280), ("16x16 2560->1280", 2, 16, 16, 2560, 1280),
                              ("8x8 1280->1280", 2, 8, 8, 1280, 1280), ("8x8 2560->1280", 2, 8, 8, 2560, 1280)]:
    x = rnd(n, h, w, ci)
    wt = rnd(co, 9 * ci, scale=(9 * ci) ** -0.5)
    b = torch.randn(co, device="cuda")
    gam, bet = torch.ones(ci, device="cuda"), torch.zeros(ci, device="cuda")
    ch = chan(x)
    gnd = dict(chan0=ch, chan1=None, gamma=gam, beta=bet, groups=32, eps=1e-5, silu=True)
    t_halo_gn = timeit(lambda: L.conv3x3(x, wt, b, halo=True, gn=gnd, stats={}))
    t_halo = timeit(lambda: L.conv3x3(x, wt, b, halo=True))
    t_old = timeit(lambda: L.conv3x3(x, wt, b))
    t_old_stats = timeit(lambda: L.conv3x3(x, wt, b, stats={}))
    t_gn = timeit(lambda: L.group_norm(x, gam, bet, 32, 1e-5, silu=True))
    t_apply = timeit(lambda: L.group_norm_apply(x, ch, gam, bet, 32, 1e-5, silu=True))
    print(f"{name:18s} halo+GN+stats {t_halo_gn:7.1f} | halo plain {t_halo:7.1f} | 9-tap conv {t_old:7.1f} (+stats {t_old_stats:7.1f}) | "
          f"cluster GN {t_gn:6.1f} | GN apply {t_apply:6.1f}")
for name, m, nn, k in [("lin 8192x320x320", 8192, 320, 320), ("lin 2048x640x640", 2048, 640, 640), ("lin 512x1280x1280", 512, 1280, 1280)]:
    x, w2, r = rnd(m, k), rnd(nn, k, scale=k ** -0.5), rnd(m, nn)
    t_plain = timeit(lambda: L.linear(x, w2, static_w=True))
    os.environ["B200SD_STAGED"] = "0"
    t_res_old = timeit(lambda: L.linear(x, w2, None, r, static_w=True))
    os.environ["B200SD_STAGED"] = "1"
    t_res_staged = timeit(lambda: L.linear(x, w2, None, r, static_w=True))
    t_rows = timeit(lambda: L.linear(x, w2, None, r, static_w=True, rowstats={}))
    t_cols = timeit(lambda: L.linear(x, w2, None, r, static_w=True, stats={}, cs_hw=m // 2))
    print(f"{name:18s} plain {t_plain:6.1f} | +res (smem prefetch) {t_res_old:6.1f} | +res staged {t_res_staged:6.1f} | +rowstats {t_rows:6.1f} | "
          f"+colstats {t_cols:6.1f}")
