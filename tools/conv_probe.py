"""Bring-up probe for the TMA convolution kernel (gemm_tc_conv.cu): parity against torch conv2d and timing of
the bench shapes under one configuration of the BDE2VID_CONV_* switches (set by the caller's environment)."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bde2vid_b200 import ops  # noqa: E402
from bde2vid_b200.engine import _pack_conv  # noqa: E402

DEV = "cuda"
bf = torch.bfloat16


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, bf)


def conv_case(n, ci, co, h, w, k, s, split=False):
    g = torch.Generator().manual_seed(n + ci + co + h + w + k + s)
    x = torch.randn(n, ci, h, w, generator=g).to(bf).float()
    wt = (torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5).to(bf).float()
    b = torch.randn(co, generator=g)
    ref = F.conv2d(x.to(DEV), wt.to(DEV), b.to(DEV), stride=s, padding=k // 2).cpu()
    pw, ld = _pack_conv(wt.to(DEV), bf, chunk_major=True)
    ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
    out = torch.zeros(n, ho, wo, co, dtype=bf, device=DEV)
    if split:
        a0, a1, c0, c1 = nhwc(x[:, :ci // 2]), nhwc(x[:, ci // 2:]), ci // 2, ci // 2
    else:
        a0, a1, c0, c1 = nhwc(x), None, ci, 0
    ops.gemm(a0, pw, b.to(DEV), out, n_img=n, h_in=h, w_in=w, c0=c0, n=co, ksize=k, stride=s, pad=k // 2, a1=a1, c1=c1,
             w_ld=ld, engine=ops.ENGINE_TCGEN05, dtype=bf, k_order=1)
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    return float((got - ref).abs().max()), float(ref.abs().max())


def lstm_case(B, hid, h, w):
    g = torch.Generator().manual_seed(hid + h + w)
    rnd = lambda *s: torch.randn(*s, generator=g).to(bf).float()  # noqa: E731
    x, hp = rnd(B, hid, h, w), rnd(B, hid, h, w) * 0.5
    cprev = torch.randn(B, hid, h, w, generator=g)
    wt = (rnd(4 * hid, 2 * hid, 3, 3) / (18 * hid) ** 0.5).to(bf).float()
    b = torch.randn(4 * hid, generator=g) * 0.1
    gates = F.conv2d(torch.cat([x, hp], 1).to(DEV), wt.to(DEV), b.to(DEV), padding=1).cpu()
    i, f_, o, gg = gates.chunk(4, 1)
    c_ref = torch.sigmoid(f_) * cprev + torch.sigmoid(i) * torch.tanh(gg)
    h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
    wi = wt.view(4, hid, 2 * hid, 3, 3).permute(1, 0, 2, 3, 4).reshape(4 * hid, 2 * hid, 3, 3)
    bi = b.view(4, hid).t().reshape(-1).contiguous()
    pw, ld = _pack_conv(wi.to(DEV), bf, chunk_major=True)
    hout = torch.zeros(B, h, w, hid, dtype=bf, device=DEV)
    cout = torch.zeros(B, h, w, hid, dtype=torch.float32, device=DEV)
    ops.gemm(nhwc(x), pw, bi.to(DEV), hout, n_img=B, h_in=h, w_in=w, c0=hid, n=4 * hid, ksize=3, stride=1, pad=1,
             a1=nhwc(hp), c1=hid, w_ld=ld, epi=ops.EPI_LSTM, c_prev=cprev.permute(0, 2, 3, 1).contiguous().to(DEV), c_out=cout,
             engine=ops.ENGINE_TCGEN05, dtype=bf, k_order=1)
    torch.cuda.synchronize()
    eh = float((hout.float().cpu().permute(0, 3, 1, 2) - h_ref).abs().max())
    ec = float((cout.cpu().permute(0, 3, 1, 2) - c_ref).abs().max())
    return eh, ec


def time_case(name, n, ci, co, h, w, k, s, lstm=False, iters=20):
    g = torch.Generator().manual_seed(1)
    wt = (torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5)
    pw, ld = _pack_conv(wt.to(DEV), bf, chunk_major=True)
    b = torch.randn(co, generator=g).to(DEV)
    ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
    if lstm:
        hid = ci // 2
        a0 = torch.randn(n, h, w, hid, device=DEV).to(bf)
        a1 = torch.randn(n, h, w, hid, device=DEV).to(bf)
        out = torch.zeros(n, ho, wo, hid, dtype=bf, device=DEV)
        cprev = torch.randn(n, ho, wo, hid, device=DEV)
        cout = torch.zeros(n, ho, wo, hid, device=DEV)
        kw = dict(a1=a1, c1=hid, epi=ops.EPI_LSTM, c_prev=cprev, c_out=cout)
        c0 = hid
    else:
        a0 = torch.randn(n, h, w, ci, device=DEV).to(bf)
        out = torch.zeros(n, ho, wo, co, dtype=bf, device=DEV)
        kw = dict(act=ops.ACT_RELU)
        c0 = ci
    run = lambda: ops.gemm(a0, pw, b, out, n_img=n, h_in=h, w_in=w, c0=c0, n=co, ksize=k, stride=s, pad=k // 2, w_ld=ld,  # noqa: E731
                           engine=ops.ENGINE_TCGEN05, dtype=bf, k_order=1, **kw)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 2.0 * n * ho * wo * co * ci * k * k
    extra = ""
    if os.environ.get("PROBE_DBG", "0") == "1":
        import ctypes as C
        import numpy as np
        from bde2vid_b200 import _lib
        lib = _lib.require_device()
        lib.bde_tc_debug_enable.argtypes = [C.c_size_t]
        lib.bde_tc_debug_read.argtypes = [C.c_void_p, C.c_size_t]
        lib.bde_tc_debug_enable(148)
        run()
        torch.cuda.synchronize()
        buf = np.zeros((148, 8), dtype=np.int64)
        lib.bde_tc_debug_read(buf.ctypes.data_as(C.c_void_p), 148)
        lib.bde_tc_debug_enable(0)
        used = buf[buf[:, 0] != 0]
        if len(used):
            m = used.mean(0)
            extra = " | ctas %d total %6d mma-wait-acc %6d mma-wait-ops %6d epi-wait %6d epi-busy %6d prod-wait-bempty %6d" % (
                len(used), m[0], m[1], m[2], m[3], m[4], m[5])
    print("  time %-22s %8.1f us  %7.1f TF/s%s" % (name, us, fl / us * 1e-6, extra), flush=True)


def main():
    cfg = {k: v for k, v in os.environ.items() if k.startswith("BDE2VID_")}
    print("config", cfg, flush=True)
    cases = [(8, 128, 64, 40, 70, 5, 1, False), (6, 64, 32, 37, 90, 5, 1, False), (8, 256, 128, 24, 44, 3, 1, False),
             (1, 128, 64, 19, 37, 3, 1, False), (2, 256, 128, 18, 22, 5, 1, False), (3, 64, 32, 40, 56, 5, 1, False),
             (2, 128, 256, 17, 22, 3, 1, True), (1, 64, 64, 8, 16, 3, 1, False), (2, 512, 1024, 33, 44, 3, 1, True)]
    if os.environ.get("BDE2VID_CONV_S2", "0") == "1":
        cases += [(2, 64, 128, 33, 44, 5, 2, False), (2, 128, 256, 16, 24, 5, 2, False)]
    ok = True
    for c in cases:
        err, mx = conv_case(*c)
        good = err <= 1e-2 * max(1.0, mx)
        ok &= good
        print("  parity", c, "err %.4g ref-max %.3g %s" % (err, mx, "OK" if good else "FAIL"), flush=True)
    for lc in [(2, 64, 17, 22), (3, 128, 9, 19), (1, 256, 33, 44)]:
        eh, ec = lstm_case(*lc)
        good = ec <= 2e-3 and eh <= 6e-3
        ok &= good
        print("  parity lstm", lc, "h err %.4g c err %.4g %s" % (eh, ec, "OK" if good else "FAIL"), flush=True)
    print("PARITY", "OK" if ok else "FAIL", flush=True)
    if not ok and os.environ.get("PROBE_TIME_ANYWAY", "0") != "1":
        return
    B = 4
    time_case("lstm L1 (B=4)", B, 128, 256, 132, 176, 3, 1, lstm=True)
    time_case("lstm L2 (B=4)", B, 256, 512, 66, 88, 3, 1, lstm=True)
    time_case("lstm L3 (B=4)", B, 512, 1024, 33, 44, 3, 1, lstm=True)
    time_case("lstm L1 (B=1)", 1, 128, 256, 132, 176, 3, 1, lstm=True)
    time_case("lstm L2 (B=1)", 1, 256, 512, 66, 88, 3, 1, lstm=True)
    time_case("lstm L3 (B=1)", 1, 512, 1024, 33, 44, 3, 1, lstm=True)
    time_case("lstm L2 (B=2)", 2, 256, 512, 66, 88, 3, 1, lstm=True)
    time_case("lstm L3 (B=2)", 2, 512, 1024, 33, 44, 3, 1, lstm=True)
    time_case("dec0 256->128 (24 fr)", 24, 256, 128, 66, 88, 5, 1)
    time_case("dec1 128->64 (24 fr)", 24, 128, 64, 132, 176, 5, 1)
    time_case("dec2 64->32 (24 fr)", 24, 64, 32, 264, 352, 5, 1, iters=5)
    time_case("enc1 64->128 s2 (24 fr)", 24, 64, 128, 132, 176, 5, 2)
    time_case("enc2 128->256 s2 (24 fr)", 24, 128, 256, 66, 88, 5, 2)


if __name__ == "__main__":
    t = time.time()
    main()
    print("probe done in %.1f s" % (time.time() - t))
