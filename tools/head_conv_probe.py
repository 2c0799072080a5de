"""Time of the head convolution (planar fp32 voxels -> bf16 NHWC, 5x5, 5 -> 32 channels) at the bench's size."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bde2vid_b200 import ops  # noqa: E402

DEV = "cuda"
g = torch.Generator().manual_seed(0)
for N in (100, 800):
    vox = torch.randn(N, 5, 264, 352, generator=g).to(DEV)
    w = (torch.randn(32, 5, 5, 5, generator=g) * 0.1).to(DEV)
    b = (torch.randn(32, generator=g) * 0.1).to(DEV)
    out = torch.empty(N, 264, 352, 32, dtype=torch.bfloat16, device=DEV)
    for _ in range(3):
        ops.head_conv(vox, w, b, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.head_conv(vox, w, b, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("head conv  %d frames 264x352: %.3f ms = %.2f us per frame, %.0f GB/s algorithmic" % (
        N, ms, ms * 1e3 / N, N * (5 * 4 + 32 * 2) * 264 * 352 / ms / 1e6))
