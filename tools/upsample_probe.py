"""Time of the bilinear x2 upsample + skip sum (decoder input) at the bench's three decoder sizes, 64 frames."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bde2vid_b200 import ops  # noqa: E402

DEV = "cuda"
g = torch.Generator().manual_seed(0)
n = 64
for (h, w, c) in ((33, 44, 256), (66, 88, 128), (132, 176, 64)):
    skip = torch.randn(n, h, w, c, generator=g).to(DEV)
    x = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).to(DEV)
    dst = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=DEV)
    for _ in range(3):
        ops.upsample2x_sum(skip, x, 1.0, n, h, w, c, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.upsample2x_sum(skip, x, 1.0, n, h, w, c, dst)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byt = n * h * w * c * (4 + 2 + 4 * 2)
    print("upsample2x + skip  %d x %dx%dx%d: %.1f us, %.0f GB/s algorithmic" % (n, h, w, c, ms * 1e3, byt / ms / 1e6))
