"""Tuning aid: CTA pairs (tcgen05.mma cta_group::2) vs single-CTA tiles on the large-M GEMM / conv shapes of the
SD-2.1 UNet, per block_n.  Weights rotated through > L2 bytes; CUDA-graph replay; CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from b200sd import lib as L  # noqa: E402


def time_graph(fn, ncopies, reps=6):
    for i in range(ncopies):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(ncopies):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * ncopies)


SHAPES = [  # (name, mode, n_img, h, c0, c1, cout, geglu)
    ("conv64_320", 1, 2, 64, 320, 0, 320, False), ("conv64_640_320", 1, 2, 64, 320, 320, 320, False),
    ("conv64_960_320", 1, 2, 64, 640, 320, 320, False), ("conv32_640", 1, 2, 32, 640, 0, 640, False),
    ("conv32_1280_640", 1, 2, 32, 640, 640, 640, False),
    ("qkv64", 0, 2, 64, 320, 0, 960, False), ("o64", 0, 2, 64, 320, 0, 320, False), ("ff2_64", 0, 2, 64, 1280, 0, 320, False),
    ("geglu64", 0, 2, 64, 320, 0, 5120, True), ("geglu32", 0, 2, 32, 640, 0, 10240, True),
    ("qkv32", 0, 2, 32, 640, 0, 1920, False), ("ff2_32", 0, 2, 32, 2560, 0, 640, False),
    ("geglu16", 0, 2, 16, 1280, 0, 10240, True),
]
out = {}
for name, mode, nimg, h, c0, c1, co, geglu in SHAPES:
    taps = 9 if mode else 1
    wbytes = co * taps * (c0 + c1) * 2
    ncopies = max(2, min(24, (160 << 20) // wbytes + 1))
    ws = [(torch.randn(co, taps * (c0 + c1), device="cuda") * 0.02).half() for _ in range(ncopies)]
    bias = torch.randn(co, device="cuda")
    m = nimg * h * h
    if mode:
        x0 = torch.randn(nimg, h, h, c0, device="cuda").half()
        x1 = torch.randn(nimg, h, h, c1, device="cuda").half() if c1 else None
    else:
        x0 = torch.randn(m, c0, device="cuda").half()
    flops = 2.0 * m * co * taps * (c0 + c1)
    rows = {}
    for pair in (1, 0):
        os.environ["B200SD_2CTA"] = str(pair)
        for bn in (0, 96, 128, 160, 192, 256):
            def fn(i, bn=bn):
                if mode:
                    return L.conv3x3(x0, ws[i], bias, x1=x1, block_n=bn)
                return L.linear(x0, ws[i], bias, geglu=geglu, block_n=bn, static_w=True)
            try:
                us = time_graph(fn, ncopies)
            except Exception:
                continue
            rows[f"{'P' if pair else 'S'}:{bn}"] = round(us, 2)
    best = sorted((v, k) for k, v in rows.items())[:5]
    out[name] = {"gflop": round(flops / 1e9, 1), "auto_pair": rows.get("P:0"), "auto_single": rows.get("S:0"), "best": best,
                 "best_tflops": round(flops / best[0][0] / 1e6, 0)}
    print(name, json.dumps(out[name]), flush=True)
    L._tiled_cache.clear()
    del ws
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/gemm_pairs.json", "w"), indent=1)
