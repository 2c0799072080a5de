"""GPU parity: the drop-in BDE2VID model (whole forward through the C ABI) against the reference's
committed outputs (tests/golden) and against the oracle port on fresh inputs."""
import numpy as np
import pytest
import torch

from bde2vid_b200 import synth
from conftest import load_golden
from oracle import oracle_torch as O
from oracle.make_golden import MODEL_CASES, VARIANT_CASES, gen_cfg, voxel_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"

# frame gate of the north star: max-abs <= 2e-3 vs the reference fp32 path (bf16 operands, fp32 accumulate);
# the CUDA-core fp32 engine must be far tighter.
TOL = {"fp32": 2e-4, "bf16": 2e-3, "bf16-simt": 2e-3}


def build_model(over, wseed, precision):
    from bde2vid_b200.model import BDE2VID
    cfg = O.full_cfg(over)
    gen_cfg = {k: cfg[k] for k in ("type", "num_bins", "basechannels", "num_encoders", "ks", "num_res_blocks",
                                    "buffer_index", "q_idx", "depths", "num_heads", "losses")}
    model = BDE2VID(generator=gen_cfg)
    sd = synth.init_state_dict(cfg, wseed, stress=True)
    missing = model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    model.generator.precision = precision
    return model, cfg, sd


@pytest.mark.parametrize("precision", ["fp32", "bf16-simt", "bf16"])
@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_golden_frames(name, precision):
    H, W, T, N, over, wseed, sid = MODEL_CASES[name]
    g = load_golden(name)
    model, cfg, sd = build_model(over, wseed, precision)
    vox, _ = voxel_inputs(sid, T, H, W, N)
    with torch.no_grad():
        out = model([{"events": v.to(DEV)} for v in vox])
    assert len(out) == T and out[0].shape == (1, 1, H, W) and out[0].dtype == torch.float32
    frames = torch.cat(out, 0).cpu().numpy()
    err = np.abs(frames - g["ref_frames"]).max()
    mse_delta = abs(float(((frames - g["ref_frames"]) ** 2).mean()))
    print(name, precision, "max-abs", float(err), "mse", mse_delta)
    assert err <= TOL[precision]
    assert mse_delta <= 1e-3
    ssim = [O.ssim_uniform7(torch.from_numpy(frames[t]), torch.from_numpy(g["ref_frames"][t])) for t in range(T)]
    assert min(ssim) >= 1 - 1e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_and_repeatability(precision):
    """Second and third calls go through the captured CUDA graph; results must not change and states
    must not leak between calls (bde2vid.py:31 resets them)."""
    H, W, T, N, over, wseed, sid = MODEL_CASES["bde2vid_56x80_T3_q0"]
    model, cfg, sd = build_model(over, wseed, precision)
    vox, _ = voxel_inputs(sid, T, H, W, N)
    inp = [{"events": v.to(DEV)} for v in vox]
    with torch.no_grad():
        a = torch.cat(model(inp), 0)
        b = torch.cat(model(inp), 0)
        c = torch.cat(model(inp), 0)
        other = torch.cat(model([{"events": torch.zeros_like(v["events"])} for v in inp]), 0)
        d = torch.cat(model(inp), 0)
    assert torch.equal(a, b) and torch.equal(b, c) and torch.equal(c, d)
    assert not torch.equal(other, a)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fresh_inputs_vs_oracle_batch2(precision):
    """Batch of 2 independent sequences, odd attention padding, against the oracle run here."""
    over = dict(depths=[2, 0, 2])
    model, cfg, sd = build_model(over, 4, precision)
    g = torch.Generator().manual_seed(9)
    T, H, W = 4, 72, 88
    vox = [torch.randn(2, 5, H, W, generator=g) * (torch.rand(2, 5, H, W, generator=g) < 0.3) for _ in range(T)]
    with torch.no_grad():
        ref = O.bde2vid_forward(sd, cfg, vox)
        out = model([{"events": v.to(DEV)} for v in vox])
    err = max(float((a.cpu() - b).abs().max()) for a, b in zip(out, ref))
    print("batch2", precision, err)
    assert err <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_events_to_frames(precision):
    """Raw events -> voxeliser -> UNet in one call, at a size that needs Croper padding."""
    H, W, T, N = 60, 90, 3, 2500          # pads to 64 x 96
    model, cfg, sd = build_model(dict(depths=[1, 0, 1]), 6, precision)
    ev = synth.gen_events(31, T, H, W, N)
    xs, ys, ts, ps, off = synth.to_loader_format_seq(ev)
    prm = O.croper_params(W, H, 3)
    vox = []
    for w in range(T):
        a, b = int(off[w]), int(off[w + 1])
        vox.append(O.pad_voxel(torch.from_numpy(O.voxel_grid(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W)))[None], prm))
    with torch.no_grad():
        ref = [O.crop_image(f, prm) for f in O.bde2vid_forward(sd, cfg, vox)]
        for _ in range(3):   # eager, capture, replay
            out = model.reconstruct_events(*(torch.from_numpy(a).to(DEV) for a in (xs, ys, ts, ps, off)), (H, W))
    assert out[0].shape == (1, 1, H, W)
    err = max(float((a.cpu() - b).abs().max()) for a, b in zip(out, ref))
    print("fused", precision, err)
    assert err <= TOL[precision]


def test_no_cpu_fallback():
    from bde2vid_b200.model import BDE2VID
    cfg = O.full_cfg(dict(depths=[1, 0, 1]))
    gen_cfg = {k: cfg[k] for k in ("type", "num_bins", "basechannels", "num_encoders", "ks", "num_res_blocks",
                                    "buffer_index", "q_idx", "depths", "num_heads", "losses")}
    model = BDE2VID(generator=gen_cfg).eval()
    with pytest.raises(RuntimeError):
        model([{"events": torch.zeros(1, 5, 64, 64)}])
    with pytest.raises(NotImplementedError):
        BDE2VID(generator=dict(gen_cfg, norm="GN"))


def test_batched_event_sequences_match_single():
    """reconstruct_events_batch(B sequences) must equal B separate reconstruct_events calls (same kernels,
    different M), and the oracle."""
    H, W, T, N = 60, 90, 3, 2500
    model, cfg, sd = build_model(dict(depths=[1, 0, 1]), 6, "bf16")
    seqs = []
    for sid in (31, 32, 33):
        ev = synth.gen_events(sid, T, H, W, N)
        seqs.append(tuple(torch.from_numpy(a).to(DEV) for a in synth.to_loader_format_seq(ev)))
    with torch.no_grad():
        single = [model.reconstruct_events(*s, (H, W)) for s in seqs]
        for _ in range(3):
            batched = model.reconstruct_events_batch(seqs, (H, W))
    for b in range(3):
        err = max(float((x - y).abs().max()) for x, y in zip(single[b], batched[b]))
        print("batched vs single", b, err)
        assert err <= 1e-3


def test_fused_attention_matches_unfused():
    """The one-kernel attention half (gather + LN + q/k/v + attention [+ proj + scatter]) against the
    GEMM + attention-kernel path it replaces, on the default depths (plain and dilated blocks, D = 3)
    and on a D = 2 configuration."""
    for name in ("bde2vid_64x96_T6", "bde2vid_56x80_T3_q0"):
        H, W, T, N, over, wseed, sid = MODEL_CASES[name]
        vox, _ = voxel_inputs(sid, T, H, W, N)
        outs = {}
        for fuse in (True, False):
            model, cfg, sd = build_model(over, wseed, "bf16")
            eng = model.generator.engine()
            if not fuse:
                eng.fuse_attn = False
                for blocks in eng.attn:
                    for blk in blocks:
                        blk["tbl"] = None
            else:
                assert all(blk["tbl"] is not None for blocks in eng.attn for blk in blocks), "fused path not selected"
            with torch.no_grad():
                outs[fuse] = torch.cat(model([{"events": v.to(DEV)} for v in vox]), 0)
        err = float((outs[True] - outs[False]).abs().max())
        print(name, "fused vs unfused", err)
        assert err <= 1e-3


def test_programmatic_dependent_launch_is_bitwise_neutral(monkeypatch):
    """The chain kernels (attention, fused MLP, TMA conv) are launched with the programmatic-stream-serialization attribute
    so that their prologues overlap the previous kernel's tail (common.cuh: launch_pdl / pdl_wait).  The overlap must not
    change a single bit: BDE2VID_PDL=0 (plain stream order) against the default, eager and through the CUDA graph, several
    replays each (a race would show up as run-to-run differences)."""
    H, W, T, N = 64, 96, 6, 2500
    vox, _ = voxel_inputs(5, T, H, W, N)
    outs = {}
    for pdl in ("0", "1"):
        monkeypatch.setenv("BDE2VID_PDL", pdl)
        model, cfg, sd = build_model(dict(depths=[2, 0, 2]), 3, "bf16")
        with torch.no_grad():
            runs = [torch.cat(model([{"events": v.to(DEV)} for v in vox]), 0).clone() for _ in range(4)]
        for r in runs[1:]:
            assert torch.equal(r, runs[0]), "PDL=%s: replays differ" % pdl
        outs[pdl] = runs[0]
    assert torch.equal(outs["0"], outs["1"])


def test_long_recurrence_bf16_vs_oracle():
    """The north-star frame gate after a long recurrent chain: T = 40 windows, bf16 tcgen05 engine (approximate
    gate activations, bf16 h state) against the fp32 oracle; max-abs <= 2e-3, MSE / SSIM deltas <= 1e-3."""
    H, W, T, N = 64, 96, 40, 2500
    over = dict(depths=[2, 0, 2])
    model, cfg, sd = build_model(over, 11, "bf16")
    vox, _ = voxel_inputs(21, T, H, W, N)
    with torch.no_grad():
        ref = torch.cat(O.bde2vid_forward(sd, cfg, vox), 0)
        out = torch.cat(model([{"events": v.to(DEV)} for v in vox]), 0).cpu()
    err = float((out - ref).abs().max())
    late = float((out[T // 2:] - ref[T // 2:]).abs().max())
    print("long recurrence T=%d: max-abs %.3e (second half %.3e) mse %.3e" % (T, err, late, float(((out - ref) ** 2).mean())))
    assert err <= 2e-3
    assert float(((out - ref) ** 2).mean()) <= 1e-3
    assert min(O.ssim_uniform7(out[t], ref[t]) for t in (0, T // 2, T - 1)) >= 1 - 1e-3


# ------------------------------------------------------------------------------------------------------------------
# architecture variants of the generator constructor (ConvGRU, concat skips, BN / IN, nwindow_size, residual tail ...)
# ------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(VARIANT_CASES))
def test_variant_golden_frames(name, precision):
    """Every architecture option of the reference constructor against frames of the UNMODIFIED reference
    (tests/golden/var_*.npz, written by oracle/make_golden.py), through the C ABI."""
    from bde2vid_b200.model import BDE2VID
    H, W, T, N, over, wseed, sid = VARIANT_CASES[name]
    g = load_golden(name)
    cfg = gen_cfg(over)
    model = BDE2VID(generator=dict(cfg))
    sd = synth.random_state_dict_like(model.state_dict(), wseed)
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    model.generator.precision = precision
    vox, _ = voxel_inputs(sid, T, H, W, N)
    with torch.no_grad():
        for _ in range(3):                       # eager, graph capture, replay
            out = model([{"events": v.to(DEV)} for v in vox])
    frames = torch.cat(out, 0).cpu().numpy()
    ref = g["ref_frames"]
    err = float(np.abs(frames - ref).max())
    # 'var_all' has the Identity output activation: the frame is the raw predI output, gate relative to its range
    scale = max(1.0, float(np.abs(ref).max()))
    print(name, precision, "max-abs", err, "scale", scale)
    assert err <= TOL[precision] * scale
    assert float(((frames - ref) ** 2).mean()) <= 1e-3


# ------------------------------------------------------------------------------------------------------------------
# the BENCHMARKED path at the benchmarked sizes (the headline had no oracle check)
# ------------------------------------------------------------------------------------------------------------------

def _bench_model(precision="bf16", stress=False):
    from bde2vid_b200.model import MODELS
    ns = {}
    exec(synth.ASSUMED_CFG_STR, ns)
    mcfg = ns["model"]
    cfg = O.full_cfg(mcfg["generator"])
    sd = synth.init_state_dict(cfg, 0, stress=stress)
    model = MODELS.build(mcfg)
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    model.generator.precision = precision
    return model, cfg, sd


def _oracle_frames(sd, cfg, ev, H, W, T):
    xs, ys, ts, ps, off = synth.to_loader_format_seq(ev)
    prm = O.croper_params(W, H, 3)
    vox = []
    for w in range(T):
        a, b = int(off[w]), int(off[w + 1])
        vox.append(O.pad_voxel(torch.from_numpy(O.voxel_grid(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W)))[None], prm))
    with torch.no_grad():
        return torch.cat([O.crop_image(f, prm) for f in O.bde2vid_forward(sd, cfg, vox)], 0)


def _gate(out, ref, what):
    err = float((out - ref).abs().max())
    mse_d = float(((out - ref) ** 2).mean())
    T = out.shape[0]
    ssim = min(O.ssim_uniform7(out[t], ref[t], data_range=2.0) for t in sorted({0, T // 2, T - 1}))
    print("%s: max-abs %.3e  mse %.3e  min ssim %.6f" % (what, err, mse_d, ssim))
    assert err <= 2e-3, what            # north-star frame gate (bf16 operands, fp32 accumulate)
    assert mse_d <= 1e-3 and ssim >= 1 - 1e-3
    return err


@pytest.mark.parametrize("B", [8, 4])
def test_bench_path_346x260_batched_vs_oracle(B):
    """bench.py's configuration: assumed cfg, seed-0 weights, 346x260 -> 264x352, EIGHT (the default; four in round 1)
    sequences batched through reconstruct_events_batch and the CUDA graph.  Level 3 has 280 / 140 windows (>= 64): the
    whole-window attention kernel, direct0, merged encoders with pitched TMA sources, 8-frame decoder chunks on the side
    stream -- all compared with the fp32 oracle."""
    H, W, T, N = 260, 346, 8, 31500
    model, cfg, sd = _bench_model()
    evs = [synth.gen_events(100 + b, T, H, W, N) for b in range(B)]
    seqs = [tuple(torch.from_numpy(a).to(DEV) for a in synth.to_loader_format_seq(ev)) for ev in evs]
    with torch.no_grad():
        for _ in range(3):                                   # eager, capture, replay
            out = model.reconstruct_events_batch(seqs, (H, W))
    eng = model.generator.engine()
    plan = eng.plan(T, B, 264, 352)
    assert True in plan.graphs, "the CUDA-graph path was not taken"
    assert plan.lv[2]["nwin"] >= 64 and eng.fuse_win256 and all(b["tbl"] is not None for b in eng.attn[2])
    for b in (0, B - 1):                                     # two of the sequences (CPU oracle: ~4 s per sequence)
        ref = _oracle_frames(sd, cfg, evs[b], H, W, T)
        got = torch.cat(out[b], 0).cpu().reshape(T, H, W)
        _gate(got, ref.reshape(T, H, W), "bench path seq %d" % b)


def test_single_sequence_T100_346x260_vs_oracle():
    """north_star: the frame gate 'after >= 100 recurrent steps' at the benchmark resolution (one sequence, T = 100)."""
    H, W, T, N = 260, 346, 100, 31500
    model, cfg, sd = _bench_model()
    ev = synth.gen_events(0, T, H, W, N)
    seq = tuple(torch.from_numpy(a).to(DEV) for a in synth.to_loader_format_seq(ev))
    with torch.no_grad():
        out = model.reconstruct_events(*seq, (H, W))
    got = torch.cat(out, 0).cpu().reshape(T, H, W)
    ref = _oracle_frames(sd, cfg, ev, H, W, T).reshape(T, H, W)
    _gate(got, ref, "T=100 single sequence")
    _gate(got[T // 2:], ref[T // 2:], "T=100, second half")


def test_config1_240x180_raw_events_vs_oracle():
    """BASELINE.json configs[0] shape (240x180 -> 184x240) fed with the ON-DISK event dtypes (int16 / float64 / bool)."""
    H, W, T, N = 180, 240, 12, 15000
    model, cfg, sd = _bench_model(stress=True)
    ev = synth.gen_events(7, T, H, W, N)
    raw = (torch.from_numpy(ev["xs"]), torch.from_numpy(ev["ys"]), torch.from_numpy(ev["ts"]), torch.from_numpy(ev["ps"]),
           torch.from_numpy(ev["offsets"]))
    with torch.no_grad():
        for _ in range(2):
            out = model.reconstruct_events(*(a.to(DEV) for a in raw), (H, W))
    got = torch.cat(out, 0).cpu().reshape(T, H, W)
    ref = _oracle_frames(sd, cfg, ev, H, W, T).reshape(T, H, W)
    _gate(got, ref, "240x180 raw ingest")


def test_config5_1280x720_vs_oracle():
    """BASELINE.json configs[4] shape: Prophesee Gen4 1280x720 (no padding), 333 333 events per window, T = 2."""
    H, W, T, N = 720, 1280, 2, 333333
    model, cfg, sd = _bench_model()
    ev = synth.gen_events(2, T, H, W, N)
    seq = tuple(torch.from_numpy(a).to(DEV) for a in synth.to_loader_format_seq(ev))
    with torch.no_grad():
        out = model.reconstruct_events(*seq, (H, W))
    got = torch.cat(out, 0).cpu().reshape(T, H, W)
    ref = _oracle_frames(sd, cfg, ev, H, W, T).reshape(T, H, W)
    _gate(got, ref, "1280x720")


def test_driver_chunk_T1000_smoke():
    """One driver chunk (subseq_L = 1000, eval_models_seq.py:216-219) at 346x260: buffers for 1000 frames stay resident
    (~22 GB), the graph captures, frames are finite and the first frames equal the T = 8 run's only where the
    bidirectional recurrence allows it (they do not: the backward chain sees 1000 frames) -- so check determinism."""
    H, W, T, N = 260, 346, 1000, 2000
    model, cfg, sd = _bench_model()
    ev = synth.gen_events(3, T, H, W, N)
    seq = tuple(torch.from_numpy(a).to(DEV) for a in synth.to_loader_format_seq(ev))
    with torch.no_grad():
        a = torch.cat(model.reconstruct_events(*seq, (H, W)), 0)
        b = torch.cat(model.reconstruct_events(*seq, (H, W)), 0)     # second call: graph capture + replay
    assert a.shape == (T, 1, H, W) and bool(torch.isfinite(a).all())
    assert float(a.min()) > 0.0 and float(a.max()) < 1.0
    assert torch.equal(a, b)
    eng = model.generator.engine()
    assert eng.plan(T, 1, 264, 352).nbytes > 15e9
    eng.plans.clear()
    torch.cuda.empty_cache()


def test_loader_contract_on_fused_path():
    """1- and 2-event windows must give zero grids on the fused events -> frames path (h5_dataset.py:219-221),
    not NaN frames; events outside the sensor are dropped and reported."""
    import warnings
    H, W, T, N = 60, 90, 4, 2500
    from bde2vid_b200.model import BDE2VID
    cfg = O.full_cfg(dict(depths=[1, 0, 1]))
    gen_cfg_ = {k: cfg[k] for k in ("type", "num_bins", "basechannels", "num_encoders", "ks", "num_res_blocks",
                                     "buffer_index", "q_idx", "depths", "num_heads", "losses")}
    model = BDE2VID(generator=gen_cfg_)
    sd = synth.init_state_dict(cfg, 6, stress=True)
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    ev = synth.gen_events(31, T, H, W, N)
    off = ev["offsets"].copy()
    off[2] = off[1] + 1                                     # window 1 has ONE event, window 2 the rest of two windows
    ev = dict(ev, offsets=off)
    xs, ys, ts, ps, _ = synth.to_loader_format_seq(ev)
    prm = O.croper_params(W, H, 3)
    vox = []
    for w in range(T):
        a, b = int(off[w]), int(off[w + 1])
        g = O.loader_voxel(ev["xs"][a:b], ev["ys"][a:b], ev["ts"][a:b], ev["ps"][a:b], 5, (H, W))
        vox.append(O.pad_voxel(torch.from_numpy(g)[None], prm))
    assert not vox[1].any()
    with torch.no_grad():
        ref = torch.cat([O.crop_image(f, prm) for f in O.bde2vid_forward(sd, cfg, vox)], 0)
        for _ in range(3):
            out = model.reconstruct_events(*(torch.from_numpy(a).to(DEV) for a in (xs, ys, ts, ps, off)), (H, W))
    got = torch.cat(out, 0).cpu()
    assert bool(torch.isfinite(got).all())
    assert float((got - ref).abs().max()) <= 2e-3
    # out-of-sensor events: dropped, counted, reported at the next call
    xs_bad = xs.copy()
    xs_bad[5] = W + 3
    with torch.no_grad():
        model.reconstruct_events(*(torch.from_numpy(a).to(DEV) for a in (xs_bad, ys, ts, ps, off)), (H, W))
        torch.cuda.synchronize()
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            model.reconstruct_events(*(torch.from_numpy(a).to(DEV) for a in (xs, ys, ts, ps, off)), (H, W))
    assert any("outside the sensor" in str(w.message) for w in wlist)


def test_fused_path_with_voxel_normalisation():
    """normalize='robust' / 'legacy' on the fused path == the oracle's loader transform (h5_dataset.py:226) + forward."""
    H, W, T, N = 60, 90, 3, 2500
    model, cfg, sd = build_model(dict(depths=[1, 0, 1]), 6, "bf16")
    ev = synth.gen_events(33, T, H, W, N)
    xs, ys, ts, ps, off = synth.to_loader_format_seq(ev)
    prm = O.croper_params(W, H, 3)
    for name, fn in (("legacy", O.legacy_norm), ("robust", lambda v: O.robust_norm(v, 0, 95))):
        vox = []
        for w in range(T):
            a, b = int(off[w]), int(off[w + 1])
            g = fn(torch.from_numpy(O.voxel_grid(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))))
            vox.append(O.pad_voxel(g[None], prm))
        with torch.no_grad():
            ref = torch.cat([O.crop_image(f, prm) for f in O.bde2vid_forward(sd, cfg, vox)], 0)
            for _ in range(3):
                out = model.reconstruct_events(*(torch.from_numpy(a).to(DEV) for a in (xs, ys, ts, ps, off)), (H, W),
                                               normalize=name)
        err = float((torch.cat(out, 0).cpu() - ref).abs().max())
        print("fused + %s norm: max-abs %.3e" % (name, err))
        assert err <= 2e-3
