"""Level-3 whole-window attention: mma.sync form (BDE2VID_ATTN_TC256=0) vs the tcgen05 form (attn_tc256.cu): agreement, time per
launch and per-phase cycle counters (bde_tc_debug_enable), for 140 windows (4 sequences of 33 x 44) and 35 (one sequence)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bde2vid_b200 import _lib, ops  # noqa: E402
from bde2vid_b200.engine import window_token_map  # noqa: E402

DEV = "cuda"
lib = _lib.require_device()
lib.bde_tc_debug_enable.argtypes = [C.c_size_t]
lib.bde_tc_debug_read.argtypes = [C.c_void_p, C.c_size_t]
heads, q_ind, Cc = 16, 1, 256
for (B, D, dil, zero) in ((4, 3, False, None), (4, 3, True, 2), (1, 3, False, None), (4, 2, False, None), (2, 1, True, None)):
    g = torch.Generator().manual_seed(1 + B + D)
    h, w = 33, 44
    qi = min(q_ind, D - 1)
    P = B * h * w
    tm, _ = window_token_map(B, h, w, (7, 7), dil, DEV)
    nwin = tm.shape[0]
    frames = [(torch.randn(P, Cc, generator=g) * 1.3 + 0.1).to(DEV) for _ in range(D)]
    if zero is not None and zero < D:
        frames[zero] = None
    wq = torch.randn(3 * Cc, Cc, generator=g) / 16
    wq[:Cc] *= 0.25
    wqkv = wq.to(torch.bfloat16).to(DEV)
    bqkv = (torch.randn(3 * Cc, generator=g) * 0.1).to(DEV)
    tbl = (torch.randn(heads, D, 169, generator=g) * 0.5).to(DEV)
    wproj = (torch.randn(Cc, Cc, generator=g) / 16).to(torch.bfloat16).to(DEV)
    bproj = (torch.randn(Cc, generator=g) * 0.1).to(DEV)

    def run(xs):
        fr = list(frames)
        fr[qi] = xs
        ops.window_attention_fused(fr, qi, tm.view(-1), nwin, Cc, heads, wqkv, bqkv, tbl, wproj, bproj, xs=xs)

    outs = {}
    modes = ("0", "1", "1c2") if 2 * nwin <= 148 else ("0", "1")     # "1c2": one window per 2-CTA cluster (head groups split)
    for mode in modes:
        os.environ["BDE2VID_ATTN_TC256"] = mode[0]
        os.environ["BDE2VID_ATTN_TC256_CLUSTER"] = "2" if mode == "1c2" else "1"
        ncta = nwin * (2 if mode == "1c2" else 1)
        xs = frames[qi].clone()
        run(xs)
        torch.cuda.synchronize()
        outs[mode] = xs.clone()
        scratch = frames[qi].clone()
        for _ in range(3):
            run(scratch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run(scratch)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 50
        lib.bde_tc_debug_enable(1024)
        run(scratch)
        torch.cuda.synchronize()
        buf = np.zeros((1024, 8), dtype=np.int64)
        lib.bde_tc_debug_read(buf.ctypes.data_as(C.c_void_p), 1024)
        lib.bde_tc_debug_enable(0)
        if mode != "0":     # MMA-warp timeline: rows [ncta, 2 ncta) = cycles since kernel start at stage 0 / 8 / 16 / 24 and at the end
            tl = buf[ncta:2 * ncta].mean(0)
            print("   MMA warp timeline (cycles): stage0 %d  stage8 %d  stage16 %d  stage24 %d  end %d" % tuple(tl[:5]))
            buf = buf[:ncta]
        used = buf[buf[:, 0] != 0]
        m = used.mean(0) if len(used) else np.zeros(8)
        names = ("total LN gather qkv tbl attn proj" if mode == "0" else "total LN wait conv - attn proj+epi").split()
        print("B=%d D=%d dil=%d zero=%s tc256=%s  %d windows  %.1f us | cycles: %s" % (
            B, D, dil, zero, mode, nwin, us, "  ".join("%s %d" % (n, v) for n, v in zip(names, m[:7]))))
    d = float((outs["0"] - outs["1"]).abs().max())
    upd = float((outs["0"] - frames[qi]).abs().max())
    print("   max |tc256 - mma.sync| = %.3e  (update magnitude %.3f, finite %s)" % (d, upd, bool(torch.isfinite(outs["1"]).all())))
    assert d <= 2e-2 * max(1.0, upd)
    if "1c2" in outs:
        d2 = float((outs["1c2"] - outs["1"]).abs().max())
        print("   max |cluster of 2 - single CTA| = %.3e" % d2)
        assert d2 <= 1e-4 * max(1.0, upd)     # same products; only the fp32 order of the projection's K = 256 sum differs
del os.environ["BDE2VID_ATTN_TC256_CLUSTER"]
print("attn_tc256 probe ok")
