"""GPU parity: voxeliser kernels (through the C ABI) vs the oracle restatement."""
import numpy as np
import pytest
import torch

from bde2vid_b200 import synth
from conftest import load_golden
from oracle import oracle_torch as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _run(ev_arrays, offsets, H, W, algo, pad=(0, 0, None, None), bins=5):
    from bde2vid_b200 import ops
    xs, ys, ts, ps = (_dev(a) for a in ev_arrays)
    oob = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = ops.voxelize_seq(xs, ys, ts, ps, _dev(offsets), bins, H, W, pad[0], pad[1], pad[2], pad[3],
                           oob_count=oob, algo=algo)
    torch.cuda.synchronize()
    return out.cpu().numpy(), int(oob.item())


def _value_gate(got, ref, mass):
    """|a-b| <= 1e-5 * max(1, sum|contrib|): the well-posed form of the 1e-5 relative gate (parallel
    accumulation order differs from the reference's event order)."""
    err = np.abs(got - ref)
    tol = 1e-5 * np.maximum(1.0, mass)
    assert (err <= tol).all(), "max err %.3e" % err.max()


@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5])
def test_unique_pixels_bit_exact(algo):
    """One event per pixel -> no accumulation-order freedom: bins AND weights must be bit-exact."""
    H, W, N = 64, 80, 64 * 80
    rng = np.random.default_rng(0)
    perm = rng.permutation(N)
    xs = (perm % W).astype(np.float32)
    ys = (perm // W).astype(np.float32)
    ts = np.sort(rng.random(N)).astype(np.float32)
    ts = (ts - ts[0]).astype(np.float32)
    ps = (rng.integers(0, 2, N) * 2 - 1).astype(np.float32)
    got, oob = _run((xs, ys, ts, ps), np.array([0, N], np.int64), H, W, algo)
    ref = O.voxel_grid(xs, ys, ts, ps, 5, (H, W))
    assert oob == 0
    assert np.array_equal(got[0], ref)
    b0, _ = O.voxel_bin_indices(ts, 5)
    # the bin with the larger |weight| among the two touched is floor(t_norm) or its right neighbour
    nz = (got[0] != 0)
    for b in range(5):
        touched = set(zip(*np.nonzero(nz[b])))
        allowed = set((int(y), int(x)) for x, y, k in zip(xs, ys, b0) if k == b or k + 1 == b)
        assert touched <= allowed


@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5])
def test_golden_small(algo):
    g = load_golden("voxel_small")
    for name in ("a", "b"):
        H, W, N, T, sid = [int(v) for v in g["meta_" + name]]
        ev = synth.gen_events(sid, T, H, W, N)
        xs, ys, ts, ps, off = synth.to_loader_format_seq(ev)
        got, oob = _run((xs, ys, ts, ps), off, H, W, algo)
        assert oob == 0
        for w in range(T):
            a, b = int(off[w]), int(off[w + 1])
            mass = O.voxel_abs_mass(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
            _value_gate(got[w], g["ref_%s_%d" % (name, w)], mass)


@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("H,W,N", [(180, 240, 15000), (260, 346, 31500)])
def test_sensor_shapes_padded(algo, H, W, N):
    T = 3
    ev = synth.gen_events(21, T, H, W, N)
    xs, ys, ts, ps, off = synth.to_loader_format_seq(ev)
    prm = O.croper_params(W, H, 3)
    pl, pr, pt, pb = prm["pad"]
    got, oob = _run((xs, ys, ts, ps), off, H, W, algo, pad=(pt, pl, prm["Hp"], prm["Wp"]))
    assert oob == 0 and got.shape == (T, 5, prm["Hp"], prm["Wp"])
    for w in range(T):
        a, b = int(off[w]), int(off[w + 1])
        ref = O.voxel_grid(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
        mass = O.voxel_abs_mass(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
        inner = got[w][:, pt:pt + H, pl:pl + W]
        _value_gate(inner, ref, mass)
        border = got[w].copy()
        border[:, pt:pt + H, pl:pl + W] = 0
        assert not border.any()          # Croper.pad zero fill


def test_gen4_shape_atomic_vs_oracle():
    H, W, N = 720, 1280, 333333
    ev = synth.gen_events(2, 2, H, W, N)
    xs, ys, ts, ps, off = synth.to_loader_format_seq(ev)
    got, oob = _run((xs, ys, ts, ps), off, H, W, 0)
    assert oob == 0
    for w in range(2):
        a, b = int(off[w]), int(off[w + 1])
        ref = O.voxel_grid(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
        mass = O.voxel_abs_mass(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
        _value_gate(got[w], ref, mass)


@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5])
def test_edge_cases(algo):
    H, W = 16, 24
    # ragged windows: empty, one event, two events, unaligned starts; duplicates on one pixel
    xs = np.array([3, 5, 5, 5, 5, 7, 0, 23, 1, 2, 3], np.float32)
    ys = np.array([2, 4, 4, 4, 4, 9, 0, 15, 1, 1, 1], np.float32)
    ts = np.array([0, 0, .1, .2, .9, 1.0, 0, .5, 0, .25, .5], np.float32)
    ps = np.array([1, 1, 1, -1, 1, -1, 1, 1, -1, 1, 1], np.float32)
    off = np.array([0, 0, 1, 6, 8, 11], np.int64)     # windows: [], [0], [1..5], [6,7], [8..10]
    got, oob = _run((xs, ys, ts, ps), off, H, W, algo)
    assert oob == 0
    assert not got[0].any()                                   # empty window -> zeros
    assert np.isnan(got[1][:, 2, 3]).all()                    # single event: dt == 0 -> NaN like the reference
    assert np.isfinite(got[1][:, :2]).all()
    for w in (2, 3, 4):
        a, b = int(off[w]), int(off[w + 1])
        ref = O.voxel_grid(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
        assert np.abs(got[w] - ref).max() <= 1e-6


def test_out_of_range_events_are_counted():
    H, W = 8, 8
    xs = np.array([1, 8, -1, 2], np.float32); ys = np.array([1, 2, 3, 9], np.float32)
    ts = np.array([0, .1, .2, 1], np.float32); ps = np.ones(4, np.float32)
    for algo in (1, 2, 3, 4, 5):
        got, oob = _run((xs, ys, ts, ps), np.array([0, 4], np.int64), H, W, algo)
        assert oob == 3
        assert got[0][0, 1, 1] == 1.0 and np.count_nonzero(got[0]) == 1


def test_reference_api_signature():
    from bde2vid_b200.voxel import events_to_voxel_torch
    ev = synth.gen_events(5, 1, 180, 240, 6000)
    xs, ys, ts, ps = synth.to_loader_format(ev, 0)
    v = events_to_voxel_torch(*(torch.from_numpy(a).to(DEV) for a in (xs, ys, ts, ps)), 5, sensor_size=(180, 240))
    assert v.shape == (5, 180, 240) and v.dtype == torch.float32 and v.is_cuda
    ref = O.voxel_grid(xs, ys, ts, ps, 5, (180, 240))
    _value_gate(v.cpu().numpy(), ref, O.voxel_abs_mass(xs, ys, ts, ps, 5, (180, 240)))
    with pytest.raises(IndexError):
        bad = torch.tensor([0., 240.], device=DEV)
        events_to_voxel_torch(bad, torch.zeros(2, device=DEV), torch.tensor([0., 1.], device=DEV),
                              torch.ones(2, device=DEV), 5, sensor_size=(180, 240))
    with pytest.raises(RuntimeError):
        events_to_voxel_torch(torch.zeros(3), torch.zeros(3), torch.tensor([0., .5, 1.]), torch.ones(3), 5)


# ------------------------------------------------------------------------------------------------------------------
# loader contract, raw (on-disk dtype) ingest, normalisation variants, hot-pixel mask  
# ------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5])
def test_min_events_loader_contract(algo):
    """h5_dataset.py:219-221: windows with fewer than 3 events give an all-zero grid (no NaN from dt == 0)."""
    from bde2vid_b200 import ops
    H, W = 16, 24
    xs = np.array([3, 5, 6, 7, 8, 9, 1, 2, 3, 4], np.float32)
    ys = np.array([2, 4, 5, 9, 1, 3, 1, 1, 1, 1], np.float32)
    ts = np.array([0, 0, .5, 0, .2, 1.0, 0, 0, 0, 0], np.float32)
    ps = np.array([1, 1, -1, 1, -1, 1, 1, 1, 1, 1], np.float32)
    off = np.array([0, 0, 1, 3, 6, 10], np.int64)      # windows of 0, 1, 2, 3 and 4 (equal timestamps) events
    out = ops.voxelize_seq(*(_dev(a) for a in (xs, ys, ts, ps)), _dev(off), 5, H, W, algo=algo, min_events=3).cpu().numpy()
    assert not out[0].any() and not out[1].any() and not out[2].any()
    ref = O.voxel_grid(xs[3:6], ys[3:6], ts[3:6], ps[3:6], 5, (H, W))
    assert np.abs(out[3] - ref).max() <= 1e-6
    assert np.isnan(out[4][:, 1, 1:5]).all()             # >= 3 events with dt == 0: the reference's NaNs, kept


@pytest.mark.parametrize("algo", [2, 5])
@pytest.mark.parametrize("H,W,N,T", [(260, 346, 31500, 3), (37, 53, 1001, 4)])
def test_raw_ingest_bit_exact_vs_loader_format(algo, H, W, N, T):
    """On-disk dtypes (int16, int16, float64, bool) converted in the kernel == the float32 loader format produced on
    the host exactly as h5_dataset.py:222-225,:414 does; unique pixels make the comparison bit-exact, random ones pass
    the value gate; windows start at unaligned offsets (ragged CSR)."""
    from bde2vid_b200 import ops
    ev = synth.gen_events(71, T, H, W, N)
    off = ev["offsets"].copy()
    off[1:-1] += np.arange(1, T) * 3 + 1                  # ragged, unaligned window starts
    ev = dict(ev, offsets=off)
    xs, ys, ts, ps, _ = synth.to_loader_format_seq(ev)
    f32 = ops.voxelize_seq(*(_dev(a) for a in (xs, ys, ts, ps)), _dev(off), 5, H, W, algo=algo, min_events=3).cpu().numpy()
    raw = ops.voxelize_raw(_dev(ev["xs"]), _dev(ev["ys"]), _dev(ev["ts"]), _dev(ev["ps"]), _dev(off), 5, H, W,
                           algo=algo, min_events=3).cpu().numpy()
    for w in range(T):
        a, b = int(off[w]), int(off[w + 1])
        ref = O.loader_voxel(ev["xs"][a:b], ev["ys"][a:b], ev["ts"][a:b], ev["ps"][a:b], 5, (H, W))
        mass = O.voxel_abs_mass(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
        _value_gate(raw[w], ref, mass)
        _value_gate(f32[w], ref, mass)
    # bit-exact form: one event per pixel
    n = H * W
    rng = np.random.default_rng(1)
    perm = rng.permutation(n)
    rx, ry = (perm % W).astype(np.int16), (perm // W).astype(np.int16)
    rt = 17.25 + np.sort(rng.random(n))                   # float64 seconds with a large offset
    rp = rng.integers(0, 2, n).astype(bool)
    one = np.array([0, n], np.int64)
    got = ops.voxelize_raw(_dev(rx), _dev(ry), _dev(rt), _dev(rp), _dev(one), 5, H, W, algo=algo).cpu().numpy()[0]
    assert np.array_equal(got, O.loader_voxel(rx, ry, rt, rp, 5, (H, W)))


def test_clustered_stream_all_algorithms():
    """Spatially clustered events (many collisions per cell) through every algorithm, incl. the warp-aggregated one."""
    H, W, N, T = 120, 160, 20000, 2
    ev = synth.gen_events_clustered(3, T, H, W, N, blobs=6, sigma=3.0)
    xs, ys, ts, ps, off = synth.to_loader_format_seq(ev)
    for algo in (1, 2, 3, 4, 5):
        got, oob = _run((xs, ys, ts, ps), off, H, W, algo)
        assert oob == 0
        for w in range(T):
            a, b = int(off[w]), int(off[w + 1])
            ref = O.voxel_grid(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
            _value_gate(got[w], ref, O.voxel_abs_mass(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W)))


@pytest.mark.parametrize("H,W,pad", [(48, 64, (0, 0)), (60, 90, (2, 3))])
def test_voxel_normalize_vs_oracle(H, W, pad):
    """LegacyNorm / RobustNorm on the device (per window, sensor area of the padded grid) vs the oracle functions, which
    are bit-equal to the reference classes (utils_func/data_augmentation.py:311-330, utils_func/utils.py:7-51)."""
    from bde2vid_b200 import ops
    T, N = 3, 3000
    ev = synth.gen_events(81, T, H, W, N)
    off = ev["offsets"].copy()
    off[1] = off[0]                                        # window 0 empty: both norms must leave zeros alone
    xs, ys, ts, ps, _ = synth.to_loader_format_seq(dict(ev, offsets=off))
    pt, pl = pad
    Hp, Wp = H + 2 * pt, W + 2 * pl
    args = [_dev(a) for a in (xs, ys, ts, ps)] + [_dev(off), 5, H, W, pt, pl, Hp, Wp]
    base = ops.voxelize_seq(*args, min_events=3)
    for mode, fn in ((ops.NORM_LEGACY, O.legacy_norm), (ops.NORM_ROBUST, lambda v: O.robust_norm(v, 0, 95)),
                     ("r2", lambda v: O.robust_norm(v, 2, 98))):
        g = base.clone()
        stats = torch.zeros(T, 4, device=DEV)
        if mode == "r2":
            ops.voxel_normalize(g, H, W, pt, pl, ops.NORM_ROBUST, 2, 98, stats=stats)
        else:
            ops.voxel_normalize(g, H, W, pt, pl, mode, stats=stats)
        g = g.cpu()
        for w in range(T):
            inner = base[w, :, pt:pt + H, pl:pl + W].cpu()
            ref = fn(inner.clone())
            got = g[w, :, pt:pt + H, pl:pl + W]
            scale = max(1.0, float(ref.abs().max()))
            assert float((got - ref).abs().max()) <= 2e-6 * scale, (mode, w, float((got - ref).abs().max()))
            ring = g[w].clone()
            ring[:, pt:pt + H, pl:pl + W] = 0
            assert not ring.any()                          # the padding ring stays zero (norm happens before Croper.pad)


def test_hot_pixel_mask_vs_oracle():
    from bde2vid_b200 import ops
    H, W, N = 40, 56, 6000
    ev = synth.gen_events_clustered(9, 1, H, W, N, blobs=3, sigma=2.0)
    pm = ev["ps"] * 2.0 - 1.0
    for num_hot in (0, 1, 25, 400):
        ref = O.hot_event_mask(ev["xs"], ev["ys"], pm, (H, W), num_hot)
        got = ops.hot_pixel_mask(_dev(ev["xs"]), _dev(ev["ys"]), _dev(ev["ps"]), H, W, num_hot).cpu().numpy()
        assert np.array_equal(got, ref.astype(np.float32)), num_hot
    # the mask multiplies the voxel grid (h5_dataset.py:364)
    mask = _dev(O.hot_event_mask(ev["xs"], ev["ys"], pm, (H, W), 25).astype(np.float32))
    off = _dev(ev["offsets"])
    got = ops.voxelize_raw(_dev(ev["xs"]), _dev(ev["ys"]), _dev(ev["ts"]), _dev(ev["ps"]), off, 5, H, W, hot_mask=mask).cpu().numpy()[0]
    ref = O.loader_voxel(ev["xs"], ev["ys"], ev["ts"], ev["ps"], 5, (H, W), hot_mask=mask.cpu().numpy())
    xs, ys, ts, ps = synth.to_loader_format(ev, 0)
    _value_gate(got, ref, O.voxel_abs_mass(xs, ys, ts, ps, 5, (H, W)))
    assert not got[:, mask.cpu().numpy() == 0].any()


@pytest.mark.parametrize("fmt", ["f32", "raw"])
@pytest.mark.parametrize("H,W,pad", [(180, 240, (2, 0, 184, 240)), (37, 53, (1, 2, 40, 57))])
def test_chunk_pipeline_many_chunks(monkeypatch, fmt, H, W, pad):
    """The default algorithm as a chain of MANY chunks (1 MB chunks: one or a few windows each): every launch zeroes the next
    chunk's grids and the launches are linked by programmatic dependent launch.  The destination starts as NaN garbage, some
    windows are empty / below min_events (their grids must still be zeroed by the previous launch), the batch stride leaves
    foreign grids between the windows untouched.  Equal to the one-chunk result (bit-exact on the zero pattern, value gate
    otherwise), with PDL on and off, with in-kernel zeroing on and off; the odd 37 x 53 sensor padded to 40 x 57 covers rows
    that are not multiples of 16 bytes."""
    from bde2vid_b200 import ops
    T, N, B = 23, 4000, 2
    ev = synth.gen_events(5, T, H, W, N)
    off = ev["offsets"].copy()
    off[7] = off[6]                                        # window 6 empty
    off[12] = off[11] + 2                                  # window 11: 2 events (< min_events 3 -> zeros)
    ev = dict(ev, offsets=off)
    xs, ys, ts, ps, _ = synth.to_loader_format_seq(ev)
    pt, pl, Hp, Wp = pad
    ge = 5 * Hp * Wp
    if fmt == "f32":
        args = [_dev(a) for a in (xs, ys, ts, ps)]
    else:
        args = [_dev(ev["xs"]), _dev(ev["ys"]), _dev(ev["ts"]), _dev(ev["ps"].view(np.uint8))]
    results = {}
    for name, env in (("one_chunk", {"BDE2VID_VOXEL_CHUNK_MB": "4096"}), ("chunks", {"BDE2VID_VOXEL_CHUNK_MB": "1"}),
                      ("chunks_nopdl", {"BDE2VID_VOXEL_CHUNK_MB": "1", "BDE2VID_PDL": "0"}),
                      ("chunks_memset", {"BDE2VID_VOXEL_CHUNK_MB": "1", "BDE2VID_VOXEL_ZERO_IN_KERNEL": "0"})):
        for k in ("BDE2VID_VOXEL_CHUNK_MB", "BDE2VID_PDL", "BDE2VID_VOXEL_ZERO_IN_KERNEL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        buf = torch.full((T, B, 5, Hp, Wp), float("nan"), device=DEV)
        ops.voxelize_seq_into(*args, _dev(off), 5, H, W, pt, pl, Hp, Wp, buf[0, 1], B * ge, min_events=3)
        torch.cuda.synchronize()
        assert torch.isnan(buf[:, 0]).all()                # the other sequence's grids between the windows are untouched
        results[name] = buf[:, 1].cpu().numpy()
    ref = results["one_chunk"]
    assert np.isfinite(ref).all() and not ref[6].any() and not ref[11].any() and ref[0].any()
    for w in (0, 5, 22):
        a, b = int(off[w]), int(off[w + 1])
        full = np.zeros((5, Hp, Wp), np.float32)
        full[:, pt:pt + H, pl:pl + W] = O.loader_voxel(ev["xs"][a:b], ev["ys"][a:b], ev["ts"][a:b], ev["ps"][a:b], 5, (H, W))
        mass = np.zeros((5, Hp, Wp), np.float32)
        mass[:, pt:pt + H, pl:pl + W] = O.voxel_abs_mass(xs[a:b], ys[a:b], ts[a:b], ps[a:b], 5, (H, W))
        _value_gate(ref[w], full, mass)
    for name in ("chunks", "chunks_nopdl", "chunks_memset"):
        got = results[name]
        assert np.array_equal(got == 0, ref == 0), name     # same zero pattern: nothing left unzeroed, nothing lost
        assert np.abs(got - ref).max() <= 1e-5, name        # only the accumulation order of colliding events differs
