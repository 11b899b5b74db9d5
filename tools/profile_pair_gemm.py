"""One conv shape (64x64, 960 -> 320, 45.3 GF) launched single-CTA and as CTA pairs, for an ncu --set full capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from b200sd import lib as L  # noqa: E402

x0 = torch.randn(2, 64, 64, 640, device="cuda").half()
x1 = torch.randn(2, 64, 64, 320, device="cuda").half()
w = (torch.randn(320, 9 * 960, device="cuda") * 0.02).half()
b = torch.randn(320, device="cuda")
xl = torch.randn(8192, 320, device="cuda").half()
wl = (torch.randn(5120, 320, device="cuda") * 0.05).half()
bl = torch.randn(5120, device="cuda")
for rep in range(3):
    for pair in ("0", "1"):
        os.environ["B200SD_2CTA"] = pair
        L.conv3x3(x0, w, b, x1=x1)
        L.linear(xl, wl, bl, geglu=True, static_w=True)
torch.cuda.synchronize()
