"""CPU: the oracle port (oracle/oracle_torch.py) against the committed reference outputs."""
import hashlib

import numpy as np
import pytest
import torch

from bde2vid_b200 import synth
from conftest import load_golden
from oracle import oracle_torch as O
from oracle.make_golden import MODEL_CASES, sub, voxel_inputs


def test_voxel_small_bit_exact():
    g = load_golden("voxel_small")
    for name in ("a", "b"):
        H, W, N, T, sid = [int(v) for v in g["meta_" + name]]
        ev = synth.gen_events(sid, T, H, W, N)
        for w in range(T):
            xs, ys, ts, ps = synth.to_loader_format(ev, w)
            mine = O.voxel_grid(xs, ys, ts, ps, 5, (H, W))
            assert np.array_equal(mine, g["ref_%s_%d" % (name, w)])


@pytest.mark.parametrize("H,W,N,sid", [(180, 240, 15000, 0), (260, 346, 31500, 1)])
def test_voxel_checksums(manifest, H, W, N, sid):
    ev = synth.gen_events(sid, 2, H, W, N)
    for w in range(2):
        rec = manifest["voxel_checksums"]["%dx%d_N%d_seq%d_w%d" % (H, W, N, sid, w)]
        xs, ys, ts, ps = synth.to_loader_format(ev, w)
        mine = O.voxel_grid(xs, ys, ts, ps, 5, (H, W))
        assert rec["oracle_bit_equal"]
        assert hashlib.sha256(mine.tobytes()).hexdigest() == rec["sha256"]
        assert int(np.count_nonzero(mine)) == rec["nonzero"]


def test_voxel_known_answers():
    xs = np.array([1., 2., 3., 0.], np.float32); ys = np.array([0., 1., 2., 3.], np.float32)
    ts = np.array([0., 0.3125, 0.5, 1.0], np.float32); ps = np.array([1., -1., 1., 1.], np.float32)
    v = O.voxel_grid(xs, ys, ts, ps, 5, (4, 4))
    assert v[0, 0, 1] == 1.0                      # first event: weight 1 in bin 0
    assert v[4, 3, 0] == 1.0 and v[:4, 3, 0].sum() == 0   # last event only in bin B-1
    assert v[1, 1, 2] == -0.75 and v[2, 1, 2] == -0.25    # t_norm = 1.25
    assert v[2, 2, 3] == 1.0
    b, tn = O.voxel_bin_indices(ts, 5)
    assert list(b) == [0, 1, 2, 4]


def test_croper(manifest):
    for key, rec in manifest["croper"].items():
        wh, e = key.split("_e")
        w, h = [int(v) for v in wh.split("x")]
        p = O.croper_params(w, h, int(e))
        assert [p["Hp"], p["Wp"]] == [rec["Hp"], rec["Wp"]]
        assert list(p["pad"]) == rec["pad"] and list(p["crop"]) == rec["crop"]


@pytest.mark.parametrize("name", ["bde2vid_56x80_T3_q0", "bde2vid_64x96_T5_buf5"])
def test_bde2vid_forward_matches_reference(manifest, name):
    H, W, T, N, over, wseed, sid = MODEL_CASES[name]
    g = load_golden(name)
    cfg = O.full_cfg(over)
    sd = synth.init_state_dict(cfg, wseed, stress=True)
    vox, _ = voxel_inputs(sid, T, H, W, N)
    taps = {}
    with torch.no_grad():
        out = O.bde2vid_forward(sd, cfg, vox, taps=taps)
    frames = torch.cat(out, 0).numpy()
    # bit-equal in the container that wrote the fixtures; allow fp32 reassociation on other hosts
    assert np.abs(frames - g["ref_frames"]).max() <= 2e-6
    assert np.abs(sub(taps["head"][0]) - g["tap_head0"]).max() <= 1e-5
    assert np.abs(sub(taps["level2"][T - 1]) - g["tap_level2_last"]).max() <= 1e-4


def test_e2vid_forward_matches_reference(manifest):
    g = load_golden("e2vid_64x96_B2_T3")
    rec = manifest["e2vid_64x96_B2_T3"]
    gen = torch.Generator().manual_seed(rec["seed"])
    sd = {}
    for k, shape in zip(rec["keys"], rec["shapes"]):
        v = torch.rand(shape, generator=gen) * 2 - 1
        sd[k] = v / max(1, v[0].numel()) ** 0.5
    xs = [torch.randn(2, 5, 64, 96, generator=gen) for _ in range(3)]
    st = None
    with torch.no_grad():
        for i, x in enumerate(xs):
            img, st = O.e2vid_recurrent_forward(sd, x, st)
            assert np.abs(img.numpy() - g["ref_frames"][i]).max() <= 2e-6
