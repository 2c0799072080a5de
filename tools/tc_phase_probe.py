"""Per-CTA phase timing of the tcgen05 GEMM kernel (bring-up / tuning tool; needs a B200).
Slots (clock64, per CTA): 0 start, 1 setup done (barriers+TMEM), 2 producer prologue done, 3 producer loop done,
4 first stage full (MMA can start), 5 accumulator ready, 6 epilogue done, 7 TMEM freed."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bde2vid_b200 import _lib, ops  # noqa: E402
from bde2vid_b200.engine import _pack_conv  # noqa: E402

lib = _lib.require_device()
lib.bde_tc_debug_enable.argtypes = [C.c_size_t]
lib.bde_tc_debug_read.argtypes = [C.c_void_p, C.c_size_t]

CASES = [
    # name, n_img, h, w, c0, c1, n, k, stride, epi
    ("kv L1 linear K=64", 1, 72618, 1, 64, 0, 128, 1, 1, "store"),
    ("q L3 linear K=256", 1, 1715, 1, 256, 0, 256, 1, 1, "store"),
    ("fc1 L1 K=64 N=256", 1, 23232, 1, 64, 0, 256, 1, 1, "store"),
    ("fc2 L3 K=1024", 1, 1452, 1, 1024, 0, 256, 1, 1, "store"),
    ("LSTM L1", 1, 132, 176, 64, 64, 256, 3, 1, "lstm"),
    ("LSTM L2", 1, 66, 88, 128, 128, 512, 3, 1, "lstm"),
    ("LSTM L3", 1, 33, 44, 256, 256, 1024, 3, 1, "lstm"),
    ("dec2 5x5 64->32", 1, 264, 352, 64, 0, 32, 5, 1, "store"),
    ("dec1 5x5 128->64", 1, 132, 176, 128, 0, 64, 5, 1, "store"),
    ("enc0 5x5s2 32->64", 4, 264, 352, 32, 0, 64, 5, 2, "store"),
]
CASES_B4 = [
    ("LSTM L1 B4", 4, 132, 176, 64, 64, 256, 3, 1, "lstm"),
    ("LSTM L2 B4", 4, 66, 88, 128, 128, 512, 3, 1, "lstm"),
    ("LSTM L3 B4", 4, 33, 44, 256, 256, 1024, 3, 1, "lstm"),
    ("dec2 5x5 64->32 x8", 8, 264, 352, 64, 0, 32, 5, 1, "store"),
    ("dec1 5x5 128->64 x8", 8, 132, 176, 128, 0, 64, 5, 1, "store"),
    ("dec0 5x5 256->128 x8", 8, 66, 88, 256, 0, 128, 5, 1, "store"),
    ("enc0 5x5s2 32->64 x24", 24, 264, 352, 32, 0, 64, 5, 2, "store"),
    ("enc1 5x5s2 64->128 x24", 24, 132, 176, 64, 0, 128, 5, 2, "store"),
    ("enc2 5x5s2 128->256 x24", 24, 66, 88, 128, 0, 256, 5, 2, "store"),
    ("head 5x5 8->32 x24", 24, 264, 352, 8, 0, 32, 5, 1, "store"),
    ("fc1 L3 B4", 1, 5808, 1, 256, 0, 1024, 1, 1, "store"),
    ("fc2 L3 B4", 1, 5808, 1, 1024, 0, 256, 1, 1, "store"),
    ("proj L3 B4", 1, 6860, 1, 256, 0, 256, 1, 1, "store"),
]
CONFIGS = [("tile2d=1 deep=0", "1", "0"), ("tile2d=0 deep=0", "0", "0"), ("tile2d=1 deep=1", "1", "1")]
if len(sys.argv) > 1 and sys.argv[1] == "bn":
    # N-tile sweep on the batch-4 shapes
    SWEEP = True
else:
    SWEEP = False



def run_case(name, n_img, h, w, c0, c1, n, k, stride, epi):
    dev = "cuda"
    a0 = torch.randn(n_img, h, w, c0, device=dev).to(torch.bfloat16)
    a1 = torch.randn(n_img, h, w, c1, device=dev).to(torch.bfloat16) if c1 else None
    wt = torch.randn(n, c0 + c1, k, k, device=dev) / ((c0 + c1) * k * k) ** 0.5
    cm = k > 1 and c0 % 64 == 0
    pw, ld = _pack_conv(wt, torch.bfloat16, chunk_major=cm)
    bias = torch.randn(n, device=dev)
    ho, wo = (h + 2 * (k // 2) - k) // stride + 1, (w + 2 * (k // 2) - k) // stride + 1
    M = n_img * ho * wo
    kw = dict(n_img=n_img, h_in=h, w_in=w, c0=c0, n=n, ksize=k, stride=stride, pad=k // 2, a1=a1, c1=c1, w_ld=ld,
              engine=ops.ENGINE_TCGEN05, dtype=torch.bfloat16, k_order=int(cm))
    if epi == "lstm":
        out = torch.zeros(M, n // 4, device=dev, dtype=torch.bfloat16)
        kw.update(epi=ops.EPI_LSTM, c_prev=torch.randn(M, n // 4, device=dev), c_out=torch.zeros(M, n // 4, device=dev))
    else:
        out = torch.zeros(M, n, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a0, pw, bias, out, **kw)
    torch.cuda.synchronize()
    # plain timing: 20 back-to-back launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm(a0, pw, bias, out, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    ncta_max = 1 << 16
    lib.bde_tc_debug_enable(ncta_max)
    ops.gemm(a0, pw, bias, out, **kw)
    torch.cuda.synchronize()
    buf = np.zeros((ncta_max, 8), dtype=np.int64)
    lib.bde_tc_debug_read(buf.ctypes.data_as(C.c_void_p), ncta_max)
    lib.bde_tc_debug_enable(0)
    used = buf[buf[:, 7] != 0]
    d = lambda a, b: np.median(used[:, b] - used[:, a])  # noqa: E731
    kb = (k * k * (c0 + c1) + 63) // 64
    flops = 2.0 * M * n * k * k * (c0 + c1)
    print("%-20s M=%-6d N=%-5d kb=%-3d ctas=%-5d %7.1fus %6.1fTF/s | setup %4d prol %5d first-full %5d prod-loop %6d "
          "acc-ready %6d epi %5d total %6d" % (name, M, n, kb, len(used), us, flops / us / 1e6, d(0, 1), d(1, 2), d(1, 4),
                                                d(2, 3), d(0, 5), d(5, 6), d(0, 7)))


if SWEEP:
    for bn in ("0", "128", "256"):
        os.environ["BDE2VID_TC_BN"] = bn
        print("=== BN override " + bn)
        for case in (CASES if len(sys.argv) > 2 and sys.argv[2] == "b1" else CASES_B4):
            if case[6] % max(int(bn), 1) == 0 and (len(sys.argv) < 4 or sys.argv[3] in case[0]):
                run_case(*case)
    sys.exit(0)

for cfg_name, t2d, deep in CONFIGS:
    os.environ["BDE2VID_TC_TILE2D"] = t2d
    os.environ["BDE2VID_TC_DEEP"] = deep
    print("=== " + cfg_name)
    for case in CASES:
        if case[7] == 1 and cfg_name != CONFIGS[0][0]:
            continue
        run_case(*case)
