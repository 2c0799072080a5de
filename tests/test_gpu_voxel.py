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


@pytest.mark.parametrize("algo", [1, 2, 3, 4])
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


@pytest.mark.parametrize("algo", [1, 2, 3, 4])
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


@pytest.mark.parametrize("algo", [1, 2, 3, 4])
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


@pytest.mark.parametrize("algo", [1, 2, 3, 4])
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
    for algo in (1, 2, 3, 4):
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
