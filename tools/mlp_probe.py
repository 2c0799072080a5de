"""Fused MLP kernels: time per launch of the level-3 kernel (C = 256, hidden 1024) for cluster sizes 1 / 2 / 4 and of the
level-1 kernel (C = 64), at the bench shapes (four 33 x 44 / 132 x 176 maps) and for one sequence."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402

from bde2vid_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
g = torch.Generator().manual_seed(0)


def bench(rows, C, reps=30):
    Hd = 4 * C
    x = (torch.randn(rows, C, generator=g)).to(DEV)
    w1 = (torch.randn(Hd, C, generator=g) / C ** 0.5).to(torch.bfloat16).to(DEV)
    w2 = (torch.randn(C, Hd, generator=g) / Hd ** 0.5).to(torch.bfloat16).to(DEV)
    b1, b2 = torch.zeros(Hd, device=DEV), torch.zeros(C, device=DEV)
    for _ in range(3):
        ops.mlp_fused(x, rows, C, Hd, w1, b1, w2, b2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.mlp_fused(x, rows, C, Hd, w1, b1, w2, b2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / reps


for B in (4, 1):
    rows = B * 33 * 44
    for cl in ("1", "2", "4"):
        os.environ["BDE2VID_MLP256_CLUSTER"] = cl
        us = bench(rows, 256)
        print("mlp256  B=%d rows=%d cluster=%s : %.1f us  (%.0f TFLOP/s)" % (B, rows, cl, us, rows * 256 * 1024 * 4 / us / 1e6))
    del os.environ["BDE2VID_MLP256_CLUSTER"]
    rows = B * 132 * 176
    us = bench(rows, 64)
    print("mlp64   B=%d rows=%d : %.1f us  (%.0f TFLOP/s, %.0f GB/s of x traffic)" % (B, rows, us, rows * 64 * 256 * 4 / us / 1e6, rows * 64 * 8 / us / 1e3))


# ---- phase timestamps of the C = 256 kernel (bde_tc_debug_enable): cycles since kernel entry, thread 0 of every CTA ----------
lib = _lib.require_device()
lib.bde_tc_debug_enable.argtypes = [C.c_size_t]
lib.bde_tc_debug_read.argtypes = [C.c_void_p, C.c_size_t]
NAMES = ("setup", "LN done", "fc1(0) done", "GELU loop done", "acc2 full", "cluster sync 1", "last phase begins", "end")
for B in (4, 1):
    rows = B * 33 * 44
    tiles = (rows + 127) // 128
    for cl in (1, 2):
        os.environ["BDE2VID_MLP256_CLUSTER"] = str(cl)
        n = tiles * cl
        lib.bde_tc_debug_enable(n)
        bench(rows, 256, reps=2)
        buf = np.zeros((n, 8), dtype=np.int64)
        lib.bde_tc_debug_read(buf.ctypes.data, n)
        lib.bde_tc_debug_enable(0)
        m = buf.mean(axis=0)
        print("mlp256 B=%d cluster=%d (%d CTAs) cycles since entry: " % (B, cl, n) + "  ".join("%s %d" % (NAMES[i], m[i]) for i in range(8) if m[i] > 0))
    del os.environ["BDE2VID_MLP256_CLUSTER"]
    rows = B * 132 * 176
    n = (rows + 127) // 128
    lib.bde_tc_debug_enable(n)
    bench(rows, 64, reps=2)
    buf = np.zeros((n, 8), dtype=np.int64)
    lib.bde_tc_debug_read(buf.ctypes.data, n)
    lib.bde_tc_debug_enable(0)
    m = buf[buf[:, 5] > 0].mean(axis=0)      # persistent kernel: only the launched CTAs wrote (their first tile)
    N64 = ("setup", "LN done", "acc1 full", "GELU done", "acc2 full", "end")
    print("mlp64  B=%d (%d tiles, persistent CTAs, 2 per SM; first tile) cycles since entry: " % (B, n) + "  ".join("%s %d" % (N64[i], m[i]) for i in range(6)))
