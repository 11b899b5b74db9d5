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
