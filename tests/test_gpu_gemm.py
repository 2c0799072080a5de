"""GPU parity: the implicit-GEMM engines (CUDA-core fp32 and tcgen05 bf16) through bde_gemm,
against functional torch on the CPU (the oracle's arithmetic: conv2d / linear in fp32)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, dtype)


def run_conv(x_list, w, b, stride, engine, dtype, epi="store", act=0, chunk_major=False, **extra):
    """x_list: 1 or 2 NCHW cpu tensors (channel concat); w [Co, Ci_total, k, k]; returns NCHW cpu float."""
    from bde2vid_b200 import ops
    from bde2vid_b200.engine import _pack_conv
    n, _, h, wd = x_list[0].shape
    k = w.shape[-1]
    pw, ld = _pack_conv(w.to(DEV), dtype, chunk_major=chunk_major)
    a0 = nhwc(x_list[0], dtype)
    a1 = nhwc(x_list[1], dtype) if len(x_list) > 1 else None
    co = w.shape[0]
    ho, wo = (h + 2 * (k // 2) - k) // stride + 1, (wd + 2 * (k // 2) - k) // stride + 1
    out = torch.zeros(n, ho, wo, co, dtype=dtype, device=DEV)
    ops.gemm(a0, pw, b.to(DEV).float().contiguous(), out, n_img=n, h_in=h, w_in=wd, c0=x_list[0].shape[1], n=co,
             ksize=k, stride=stride, pad=k // 2, a1=a1, c1=0 if a1 is None else x_list[1].shape[1], w_ld=ld,
             act=act, engine=engine, dtype=dtype, k_order=int(chunk_major), **extra)
    torch.cuda.synchronize()
    return out.float().cpu().permute(0, 3, 1, 2)


def bf16r(t):
    return t.to(torch.bfloat16).float()


CONV_CASES = [
    # n, cin, cout, h, w, k, stride
    (2, 8, 32, 24, 40, 5, 1),      # head-like (cin padded to 8), K = 200 -> K tail
    (3, 32, 64, 26, 38, 5, 2),     # encoder 0, ctot < 64 path, odd M
    (1, 64, 128, 33, 44, 5, 2),    # encoder 1
    (2, 128, 256, 16, 24, 5, 2),   # encoder 2
    (1, 256, 128, 18, 22, 5, 1),   # decoder 0
    (2, 64, 64, 20, 12, 3, 1),
    (1, 64, 256, 30, 30, 1, 1),    # linear (fc1-like), M = 900
    (1, 256, 1024, 7, 50, 1, 1),   # wide N
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_simt_fp32_conv(case):
    from bde2vid_b200 import ops
    n, ci, co, h, w, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(n, ci, h, w, generator=g)
    wt = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    b = torch.randn(co, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, stride=s, padding=k // 2))
    got = run_conv([x], wt, b, s, ops.ENGINE_SIMT, torch.float32, act=ops.ACT_RELU)
    assert (got - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max())


@pytest.mark.parametrize("b_cpasync", ["0", "1"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_tcgen05_bf16_conv(case, b_cpasync):
    from bde2vid_b200 import ops
    n, ci, co, h, w, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    x = bf16r(torch.randn(n, ci, h, w, generator=g))
    wt = bf16r(torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5)
    b = torch.randn(co, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, stride=s, padding=k // 2))
    os.environ["BDE2VID_TC_B_CPASYNC"] = b_cpasync
    try:
        got = run_conv([x], wt, b, s, ops.ENGINE_TCGEN05, torch.bfloat16, act=ops.ACT_RELU)
    finally:
        os.environ["BDE2VID_TC_B_CPASYNC"] = "0"
    err = (got - ref).abs().max()
    print("tcgen05 conv", case, "b_cpasync", b_cpasync, "max err", float(err), "ref max", float(ref.abs().max()))
    # operands are exactly representable; the only rounding is the bf16 store of the output
    assert err <= 1e-2 * max(1.0, ref.abs().max())


def lstm_ref(x, h, c, wt, b):
    gates = F.conv2d(torch.cat([x, h], 1), wt, b, padding=1)
    i, f, o, g = gates.chunk(4, 1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c2), c2


@pytest.mark.parametrize("engine_name", ["simt", "tcgen05"])
@pytest.mark.parametrize("hid,h,w", [(64, 17, 22), (128, 9, 12), (256, 8, 11)])
def test_lstm_epilogue(engine_name, hid, h, w):
    from bde2vid_b200 import ops
    from bde2vid_b200.engine import _pack_conv
    tc = engine_name == "tcgen05"
    dtype = torch.bfloat16 if tc else torch.float32
    engine = ops.ENGINE_TCGEN05 if tc else ops.ENGINE_SIMT
    g = torch.Generator().manual_seed(hid + h)
    rnd = (lambda *s: bf16r(torch.randn(*s, generator=g))) if tc else (lambda *s: torch.randn(*s, generator=g))
    B = 2
    x, hp = rnd(B, hid, h, w), rnd(B, hid, h, w) * 0.5
    cp = torch.randn(B, hid, h, w, generator=g)
    wt = rnd(4 * hid, 2 * hid, 3, 3) / (18 * hid) ** 0.5
    wt = bf16r(wt) if tc else wt
    b = torch.randn(4 * hid, generator=g) * 0.1
    h_ref, c_ref = lstm_ref(x, hp, cp, wt, b)
    # gate-interleaved packing (n = 4*c + gate), as the engine does
    wi = wt.view(4, hid, 2 * hid, 3, 3).permute(1, 0, 2, 3, 4).reshape(4 * hid, 2 * hid, 3, 3)
    bi = b.view(4, hid).t().reshape(-1).contiguous()
    pw, ld = _pack_conv(wi.to(DEV), dtype)
    hout = torch.zeros(B, h, w, hid, dtype=dtype, device=DEV)
    cout = torch.zeros(B, h, w, hid, dtype=torch.float32, device=DEV)
    ops.gemm(nhwc(x, dtype), pw, bi.to(DEV), hout, n_img=B, h_in=h, w_in=w, c0=hid, n=4 * hid, ksize=3, stride=1, pad=1,
             a1=nhwc(hp, dtype), c1=hid, w_ld=ld, epi=ops.EPI_LSTM, c_prev=nhwc(cp, torch.float32), c_out=cout,
             engine=engine, dtype=dtype)
    torch.cuda.synchronize()
    eh = (hout.float().cpu().permute(0, 3, 1, 2) - h_ref).abs().max()
    ec = (cout.cpu().permute(0, 3, 1, 2) - c_ref).abs().max()
    print("lstm", engine_name, hid, "h err", float(eh), "c err", float(ec))
    assert ec <= (2e-3 if tc else 2e-5) and eh <= (6e-3 if tc else 2e-5)
    # first step: c_prev = None means zeros
    h0_ref, c0_ref = lstm_ref(x, hp, torch.zeros_like(cp), wt, b)
    ops.gemm(nhwc(x, dtype), pw, bi.to(DEV), hout, n_img=B, h_in=h, w_in=w, c0=hid, n=4 * hid, ksize=3, stride=1, pad=1,
             a1=nhwc(hp, dtype), c1=hid, w_ld=ld, epi=ops.EPI_LSTM, c_prev=None, c_out=cout, engine=engine, dtype=dtype)
    torch.cuda.synchronize()
    assert (cout.cpu().permute(0, 3, 1, 2) - c0_ref).abs().max() <= (2e-3 if tc else 2e-5)


@pytest.mark.parametrize("engine_name", ["simt", "tcgen05"])
def test_store_variants_and_scatter(engine_name):
    from bde2vid_b200 import ops
    from bde2vid_b200.engine import _pack_linear
    tc = engine_name == "tcgen05"
    dtype = torch.bfloat16 if tc else torch.float32
    engine = ops.ENGINE_TCGEN05 if tc else ops.ENGINE_SIMT
    tol = 2e-2 if tc else 3e-5
    g = torch.Generator().manual_seed(7)
    M, K, N = 777, 256, 64
    a = torch.randn(M, K, generator=g)
    wt = torch.randn(N, K, generator=g) / K ** 0.5
    if tc:
        a, wt = bf16r(a), bf16r(wt)
    b = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    pw, ld = _pack_linear(wt.to(DEV), dtype)
    ad = a.to(DEV, dtype).contiguous()
    bd = b.to(DEV)
    # fp32 output + residual + second copy (fc2 + residual, DTransformer.py:302-304)
    out = torch.zeros(M, N, dtype=torch.float32, device=DEV)
    out2 = torch.zeros(M, N, dtype=dtype, device=DEV)
    ops.gemm(ad, pw, bd, out, n_img=1, h_in=M, w_in=1, c0=K, n=N, w_ld=ld, out_f32=True, residual=res.to(DEV),
             out2=out2, engine=engine, dtype=dtype)
    ref = res + F.linear(a, wt, b)
    torch.cuda.synchronize()
    assert (out.cpu() - ref).abs().max() <= tol
    assert (out2.float().cpu() - ref).abs().max() <= (5e-2 if tc else tol)
    # GELU
    outg = torch.zeros(M, N, dtype=dtype, device=DEV)
    ops.gemm(ad, pw, bd, outg, n_img=1, h_in=M, w_in=1, c0=K, n=N, w_ld=ld, act=ops.ACT_GELU, engine=engine, dtype=dtype)
    torch.cuda.synchronize()
    assert (outg.float().cpu() - F.gelu(F.linear(a, wt, b))).abs().max() <= (3e-2 if tc else tol)
    # in-place residual (out aliases residual), as the executor uses it
    xs = res.to(DEV).clone()
    ops.gemm(ad, pw, bd, xs, n_img=1, h_in=M, w_in=1, c0=K, n=N, w_ld=ld, out_f32=True, residual=xs, engine=engine, dtype=dtype)
    torch.cuda.synchronize()
    assert (xs.cpu() - ref).abs().max() <= tol
    # scatter: rows -> destination rows (-1 dropped), accumulate into fp32
    P = 900
    perm = torch.randperm(P, generator=g)[:M].to(torch.int32)
    perm[::7] = -1
    dst0 = torch.randn(P, N, generator=g)
    dst = dst0.to(DEV).clone()
    ops.gemm(ad, pw, bd, dst, n_img=1, h_in=M, w_in=1, c0=K, n=N, w_ld=ld, epi=ops.EPI_SCATTER, row_map=perm.to(DEV),
             engine=engine, dtype=dtype)
    torch.cuda.synchronize()
    refd = dst0.clone()
    lin = F.linear(a, wt, b)
    keep = perm >= 0
    refd[perm[keep].long()] += lin[keep]
    assert (dst.cpu() - refd).abs().max() <= tol


def test_tcgen05_large_k_and_m_tiles():
    """Long K loop (many pipeline wraps) and many M tiles, checked against the CUDA-core engine."""
    from bde2vid_b200 import ops
    from bde2vid_b200.engine import _pack_conv
    g = torch.Generator().manual_seed(11)
    n, ci, co, h, w = 2, 512, 128, 40, 44
    x = bf16r(torch.randn(n, ci, h, w, generator=g))
    wt = bf16r(torch.randn(co, ci, 3, 3, generator=g) / (9 * ci) ** 0.5)
    b = torch.randn(co, generator=g)
    got_tc = run_conv([x], wt, b, 1, ops.ENGINE_TCGEN05, torch.bfloat16)
    got_simt = run_conv([x], wt, b, 1, ops.ENGINE_SIMT, torch.bfloat16)
    ref = F.conv2d(x, wt, b, padding=1)
    print("large: tc-vs-ref", float((got_tc - ref).abs().max()), "simt-vs-ref", float((got_simt - ref).abs().max()))
    assert (got_tc - ref).abs().max() <= 2e-2
    assert (got_simt - ref).abs().max() <= 2e-2


@pytest.mark.parametrize("engine_name", ["simt", "tcgen05"])
@pytest.mark.parametrize("case", [(2, 64, 128, 33, 44, 5, 2), (1, 128, 64, 19, 37, 3, 1), (2, 256, 128, 18, 22, 5, 1),
                                  (3, 64, 32, 40, 56, 5, 1)])
def test_chunk_major_k_order(engine_name, case):
    """k_order=1 (64-channel chunk outermost) must give the same convolution; sizes exercise the 2-D pixel
    tiles of the tcgen05 engine with ragged borders."""
    from bde2vid_b200 import ops
    tc = engine_name == "tcgen05"
    n, ci, co, h, w, k, s = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    x = bf16r(torch.randn(n, ci, h, w, generator=g))
    wt = bf16r(torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5)
    b = torch.randn(co, generator=g)
    ref = F.conv2d(x, wt, b, stride=s, padding=k // 2)
    got = run_conv([x], wt, b, s, ops.ENGINE_TCGEN05 if tc else ops.ENGINE_SIMT, torch.bfloat16, chunk_major=True)
    err = (got - ref).abs().max()
    print("chunk-major", engine_name, case, float(err))
    assert err <= 1e-2 * max(1.0, ref.abs().max())


@pytest.mark.parametrize("C,N,D", [(64, 192, 3), (256, 768, 3), (128, 384, 2), (64, 256, 1), (256, 1024, 1)])
def test_tcgen05_layernorm_gather_gemm(C, N, D):
    """Fused A operand: window-token gather + LayerNorm (affine folded into the weights) feeding the GEMM.
    Reference: layer_norm + linear in fp32 on the CPU (DTransformer.py:183-189, :281)."""
    from bde2vid_b200 import ops
    from bde2vid_b200.engine import _pack_linear, window_token_map
    g = torch.Generator().manual_seed(C + N)
    B, H, W = 2, 9, 13
    frames = [torch.randn(B * H * W, C, generator=g) * 2 + 0.5 for _ in range(D)]
    if D >= 3:
        frames[2] = None
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    Wt = torch.randn(N, C, generator=g) / C ** 0.5
    b = torch.randn(N, generator=g)
    Wf, bf = Wt * gamma[None, :], Wt @ beta + b            # folded affine
    pw, ld = _pack_linear(Wf.to(DEV), torch.bfloat16)
    if D == 1:
        tm, rows = None, B * H * W
        src = frames[0]
        ref = F.linear(F.layer_norm(src, (C,), gamma, beta, 1e-5), Wt, b)
        ntok = 1
    else:
        tmd, _ = window_token_map(B, H, W, (7, 7), True, DEV)
        nwin, ntok = tmd.shape
        tm = tmd.view(-1)
        rows = nwin * D * ntok
        tmc = tmd.cpu().long()
        toks = []
        for d in range(D):
            s = torch.zeros(B * H * W, C) if frames[d] is None else frames[d]
            toks.append(s[tmc.clamp(min=0).reshape(-1)].reshape(nwin, ntok, C) * (tmc >= 0).unsqueeze(-1))
        x = torch.stack(toks, 1).reshape(rows, C)
        ref = F.linear(F.layer_norm(x, (C,), gamma, beta, 1e-5), Wt, b)
    out = torch.zeros(rows, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(None, pw, bf.to(DEV).contiguous(), out, n_img=1, h_in=rows, w_in=1, c0=C, n=N, w_ld=ld,
             engine=ops.ENGINE_TCGEN05, dtype=torch.bfloat16,
             ln_frames=[None if f is None else f.to(DEV) for f in frames], ln_tok_map=tm, ln_n_tok=ntok)
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max()
    print("ln-gemm", C, N, D, float(err), float(ref.abs().max()))
    assert err <= 3e-2 * max(1.0, float(ref.abs().max()))


# ---------------------------------------------------------------------------------------------------------------------
# TMA-fed persistent convolution kernel (gemm_tc_conv.cu): every form it can take, against torch conv2d
# ---------------------------------------------------------------------------------------------------------------------
TMA_CONV_CASES = [
    # n, cin, cout, h, w, k, stride
    (8, 128, 64, 40, 70, 5, 1),     # decoder-like, enough tiles for the dual-tile form, ragged right / bottom border
    (6, 64, 32, 37, 90, 5, 1),      # N = 32 (dec2-like)
    (8, 256, 128, 24, 44, 3, 1),    # 3x3, N = 128
    (1, 64, 64, 8, 16, 3, 1),       # exactly one tile
    (2, 64, 128, 33, 44, 5, 2),     # stride 2: strided TMA taps
    (2, 128, 256, 16, 24, 5, 2),
    (4, 256, 512, 33, 44, 1, 1),    # 1x1: a per-pixel linear layer on an NHWC map (k | v precompute)
]
TMA_CONV_MODES = {
    "default": {},
    "single_tile": {"BDE2VID_CONV_DUAL": "0"},
    "tap_boxes": {"BDE2VID_CONV_HALO": "0"},
    "cta_pair": {"BDE2VID_CONV_PAIR": "1", "BDE2VID_CONV_DUAL": "0"},
    "x_major_atoms": {"BDE2VID_CONV_SWAP": "0"},
    "generic_epilogue": {"BDE2VID_CONV_EPI_SPEC": "0"},
    "old_engine": {"BDE2VID_CONV_TMA": "0"},
}


@pytest.mark.parametrize("mode", sorted(TMA_CONV_MODES))
@pytest.mark.parametrize("case", TMA_CONV_CASES)
def test_conv_tma_forms(case, mode):
    from bde2vid_b200 import ops
    n, ci, co, h, w, k, s = case
    g = torch.Generator().manual_seed(sum(case) + 5)
    x = bf16r(torch.randn(n, ci, h, w, generator=g))
    wt = bf16r(torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5)
    b = torch.randn(co, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, stride=s, padding=k // 2))
    old = {kk: os.environ.get(kk) for kk in TMA_CONV_MODES[mode]}
    os.environ.update(TMA_CONV_MODES[mode])
    try:
        got = run_conv([x], wt, b, s, ops.ENGINE_TCGEN05, torch.bfloat16, act=ops.ACT_RELU, chunk_major=True)
    finally:
        for kk, v in old.items():
            if v is None:
                os.environ.pop(kk, None)
            else:
                os.environ[kk] = v
    err = (got - ref).abs().max()
    print("conv_tma", mode, case, float(err))
    assert err <= 1e-2 * max(1.0, ref.abs().max())


@pytest.mark.parametrize("pair", ["0", "1"])
@pytest.mark.parametrize("B,hid,h,w", [(2, 64, 17, 22), (3, 128, 9, 19), (1, 256, 33, 44), (4, 64, 40, 48)])
def test_conv_tma_lstm_epilogue(B, hid, h, w, pair):
    """ConvLSTM step through the TMA conv kernel (chunk-major weights, two A sources, specialised gate epilogue;
    submodules.py:316-332), with and without the cta_group::2 pair form."""
    from bde2vid_b200 import ops
    from bde2vid_b200.engine import _pack_conv
    g = torch.Generator().manual_seed(hid + h + w)
    rnd = lambda *s: bf16r(torch.randn(*s, generator=g))  # noqa: E731
    x, hp = rnd(B, hid, h, w), rnd(B, hid, h, w) * 0.5
    cp = torch.randn(B, hid, h, w, generator=g)
    wt = bf16r(rnd(4 * hid, 2 * hid, 3, 3) / (18 * hid) ** 0.5)
    b = torch.randn(4 * hid, generator=g) * 0.1
    h_ref, c_ref = lstm_ref(x, hp, cp, wt, b)
    wi = wt.view(4, hid, 2 * hid, 3, 3).permute(1, 0, 2, 3, 4).reshape(4 * hid, 2 * hid, 3, 3)
    bi = b.view(4, hid).t().reshape(-1).contiguous()
    pw, ld = _pack_conv(wi.to(DEV), torch.bfloat16, chunk_major=True)
    hout = torch.zeros(B, h, w, hid, dtype=torch.bfloat16, device=DEV)
    cout = torch.zeros(B, h, w, hid, dtype=torch.float32, device=DEV)
    old = os.environ.get("BDE2VID_CONV_PAIR")
    os.environ["BDE2VID_CONV_PAIR"] = pair
    try:
        ops.gemm(nhwc(x, torch.bfloat16), pw, bi.to(DEV), hout, n_img=B, h_in=h, w_in=w, c0=hid, n=4 * hid, ksize=3, stride=1,
                 pad=1, a1=nhwc(hp, torch.bfloat16), c1=hid, w_ld=ld, epi=ops.EPI_LSTM, c_prev=nhwc(cp, torch.float32),
                 c_out=cout, engine=ops.ENGINE_TCGEN05, dtype=torch.bfloat16, k_order=1)
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("BDE2VID_CONV_PAIR", None)
        else:
            os.environ["BDE2VID_CONV_PAIR"] = old
    eh = (hout.float().cpu().permute(0, 3, 1, 2) - h_ref).abs().max()
    ec = (cout.cpu().permute(0, 3, 1, 2) - c_ref).abs().max()
    print("conv_tma lstm pair", pair, (B, hid, h, w), float(eh), float(ec))
    assert ec <= 2e-3 and eh <= 6e-3


@pytest.mark.parametrize("k,stride", [(3, 1), (5, 2)])
def test_conv_tma_pitched_source(k, stride):
    """A operand = channel slice [C, 2C) of a wider NHWC map through a0_ld (how each ConvLSTM chain reads its half of the
    merged forward / backward encoder conv output)."""
    from bde2vid_b200 import ops
    from bde2vid_b200.engine import _pack_conv
    g = torch.Generator().manual_seed(k + stride)
    n, C, co, h, w = 3, 64, 128, 24, 40
    wide = bf16r(torch.randn(n, 2 * C, h, w, generator=g))
    wt = bf16r(torch.randn(co, C, k, k, generator=g) / (C * k * k) ** 0.5)
    b = torch.randn(co, generator=g)
    ref = F.conv2d(wide[:, C:], wt, b, stride=stride, padding=k // 2)
    pw, ld = _pack_conv(wt.to(DEV), torch.bfloat16, chunk_major=True)
    a = nhwc(wide, torch.bfloat16)                                   # [n, h, w, 2C]
    ho, wo = (h + 2 * (k // 2) - k) // stride + 1, (w + 2 * (k // 2) - k) // stride + 1
    out = torch.zeros(n, ho, wo, co, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a[..., C:], pw, b.to(DEV), out, n_img=n, h_in=h, w_in=w, c0=C, n=co, ksize=k, stride=stride, pad=k // 2, w_ld=ld,
             engine=ops.ENGINE_TCGEN05, dtype=torch.bfloat16, k_order=1, a0_ld=2 * C)
    torch.cuda.synchronize()
    err = (out.float().cpu().permute(0, 3, 1, 2) - ref).abs().max()
    print("pitched", k, stride, float(err))
    assert err <= 1e-2 * max(1.0, ref.abs().max())
    # the cp.async engine refuses a pitched source loudly instead of reading the wrong pixels
    os.environ["BDE2VID_CONV_TMA"] = "0"
    try:
        with pytest.raises(RuntimeError):
            ops.gemm(a[..., C:], pw, b.to(DEV), out, n_img=n, h_in=h, w_in=w, c0=C, n=co, ksize=k, stride=stride, pad=k // 2,
                     w_ld=ld, engine=ops.ENGINE_TCGEN05, dtype=torch.bfloat16, k_order=1, a0_ld=2 * C)
    finally:
        os.environ.pop("BDE2VID_CONV_TMA", None)
