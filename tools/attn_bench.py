"""Times the attention kernel on the UNet's self-attention shapes with and without the stream-K schedule
(B200SD_ATTN_STREAMK=0): 50 back-to-back launches between CUDA events, best of 5."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sd  # noqa: E402,F401
from b200sd import lib  # noqa: E402


def timed(fn, n=50, reps=5):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n * 1e3)
    return best


for batch, heads, s in ((2, 5, 4096), (1, 5, 4096), (2, 10, 2304), (2, 10, 1024), (2, 20, 256)):
    c = heads * 64
    q, k, v = (torch.randn(batch * s, c, device="cuda").half() for _ in range(3))
    out = torch.empty_like(q)
    row = {"batch": batch, "heads": heads, "s": s}
    for mode in ("1", "0"):
        os.environ["B200SD_ATTN_STREAMK"] = mode
        lib.attention(q, k, v, batch, heads, s, s, out=out)
        us = timed(lambda: lib.attention(q, k, v, batch, heads, s, s, out=out))
        row["streamk_us" if mode == "1" else "tiles_us"] = round(us, 1)
        row["streamk_tflops" if mode == "1" else "tiles_tflops"] = round(4.0 * batch * heads * s * s * 64 / us * 1e-6, 1)
    print(row, flush=True)
