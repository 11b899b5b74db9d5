"""Per-launch cost of dependent GEMM chains (run on the GPU box): a CUDA graph of back-to-back, data-dependent
launches of one UNet GEMM shape, replayed with programmatic dependent launch on / off and with the shared-memory
footprint capped (two CTAs per SM) or not.  Prints one JSON line per (shape, mode).

    python tools/chain_bench.py [--pairs 25]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from b200sd import lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=25)
args = ap.parse_args()
L.load()
dev = torch.device("cuda")
g = torch.Generator(device="cpu").manual_seed(0)

# (name, kind, M or (n, h, w), C, N): a pair is C -> N followed by N -> C so the chain is data dependent
SHAPES = [
    ("lin 8192x320x320", "lin", 8192, 320, 320),
    ("lin 2048x640x640", "lin", 2048, 640, 640),
    ("lin 512x1280x1280", "lin", 512, 1280, 1280),
    ("lin 8192x320->1280", "lin", 8192, 320, 1280),
    ("conv 64x64 320->320", "conv", (2, 64, 64), 320, 320),
    ("conv 32x32 640->640", "conv", (2, 32, 32), 640, 640),
    ("conv 16x16 1280->1280", "conv", (2, 16, 16), 1280, 1280),
    ("conv 8x8 1280->1280", "conv", (2, 8, 8), 1280, 1280),
]


def build(kind, mshape, c, n):
    if kind == "lin":
        x = torch.randn(mshape, c, generator=g).half().to(dev)
        w1 = (torch.randn(n, c, generator=g) * (c ** -0.5)).half().to(dev)
        w2 = (torch.randn(c, n, generator=g) * (n ** -0.5)).half().to(dev)
        res = torch.randn(mshape, c, generator=g).half().to(dev)

        def pair(t):
            return L.linear(L.linear(t, w1, static_w=True), w2, None, res, static_w=True)
        flops = 2.0 * mshape * c * n * 2
    else:
        nimg, h, w = mshape
        x = torch.randn(nimg, h, w, c, generator=g).half().to(dev)
        w1 = (torch.randn(n, 9 * c, generator=g) * ((9 * c) ** -0.5)).half().to(dev)
        w2 = (torch.randn(c, 9 * n, generator=g) * ((9 * n) ** -0.5)).half().to(dev)
        res = torch.randn(nimg, h, w, c, generator=g).half().to(dev)

        def pair(t):
            return L.conv3x3(L.conv3x3(t, w1), w2, None, res)
        flops = 2.0 * nimg * h * w * 9 * c * n * 2
    return x, pair, flops


def measure(x, pair, pairs):
    t = x
    for _ in range(2):  # eager warm-up: weight tiling, attributes
        t = pair(t)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        t = x
        for _ in range(pairs):
            t = pair(t)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for name, kind, mshape, c, n in SHAPES:
    for smem_kb in (220, 108):
        for pdl in (0, 1):
            os.environ["B200SD_SMEM_KB"] = str(smem_kb)
            L.load().b200sd_set_pdl(pdl)
            x, pair, flops = build(kind, mshape, c, n)
            ms = measure(x, pair, args.pairs)
            us = ms * 1e3 / (2 * args.pairs)
            print(json.dumps({"shape": name, "smem_kb": smem_kb, "pdl": pdl, "us_per_launch": round(us, 2),
                              "tflops": round(flops * args.pairs / (ms * 1e-3) / 1e12, 1)}), flush=True)
os.environ.pop("B200SD_SMEM_KB", None)
L.load().b200sd_set_pdl(0)
