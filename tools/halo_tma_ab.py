"""A/B of the plain 3x3 convolution forms on the UNet's shapes: 9-tap implicit GEMM vs halo reuse with TMA patches
(halo=2), per call in a CUDA graph of 20 back-to-back calls (weights L2-resident) -- plus residual + per-image bias."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sd  # noqa: E402,F401
from b200sd import lib as L  # noqa: E402


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).half()


def timeit(fn, reps=20, rounds=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    best = 1e9
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / reps


for name, n, h, w, c0, c1, co in [("64x64 320->320", 2, 64, 64, 320, 0, 320), ("64x64 640->320", 2, 64, 64, 320, 320, 320),
                                  ("64x64 960->320", 2, 64, 64, 640, 320, 320), ("32x32 640->640", 2, 32, 32, 640, 0, 640),
                                  ("32x32 1280->640", 2, 32, 32, 640, 640, 640), ("32x32 1920->640", 2, 32, 32, 1280, 640, 640),
                                  ("32x32 320->640", 2, 32, 32, 320, 0, 640), ("16x16 1280->1280", 2, 16, 16, 1280, 0, 1280),
                                  ("96x96 320->320 (XL)", 2, 96, 96, 320, 0, 320), ("48x48 640->640 (XL)", 2, 48, 48, 640, 0, 640),
                                  ("64x64 320->320 b16", 16, 64, 64, 320, 0, 320)]:
    x0 = rnd(n, h, w, c0)
    x1 = rnd(n, h, w, c1) if c1 else None
    wt = rnd(co, 9 * (c0 + c1), scale=(9 * (c0 + c1)) ** -0.5)
    temb = torch.randn(n, co, device="cuda")
    res = rnd(n, h, w, co)
    out = torch.empty(n, h, w, co, device="cuda", dtype=torch.float16)
    kw = dict(x1=x1, bias_rows=h * w, bias_stride=co, out=out)
    t_old = timeit(lambda: L.conv3x3(x0, wt, temb, res, **kw))
    t_new = timeit(lambda: L.conv3x3(x0, wt, temb, res, halo=2, **kw))
    gf = 2.0 * n * h * w * co * 9 * (c0 + c1) * 1e-9
    print(f"{name:22s} 9-tap {t_old:7.1f} us ({gf / t_old * 1e3:6.0f} TF/s) | TMA halo {t_new:7.1f} us ({gf / t_new * 1e3:6.0f} TF/s)",
          flush=True)
