"""TEST INFRASTRUCTURE ONLY -- write tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

Every array named ``ref_*`` in the fixtures is an output of the reference's own code
(events_to_voxel_torch, BDE2VID.forward, E2VIDRecurrent.forward, Croper), imported through
oracle/ref_shim.py.  Arrays named ``tap_*`` are intermediate tensors of the oracle port taken
in the same run *after* its final output was checked bit-equal to the reference's.
Inputs are regenerated from seeds by the tests (bde2vid_b200/synth.py), so fixtures stay small.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim, oracle_torch as O  # noqa: E402
from bde2vid_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = {
    # name: (H, W, T, N events/window, cfg overrides, weight seed, event seq_id)
    "bde2vid_64x96_T6": (64, 96, 6, 2500, {}, 0, 7),
    "bde2vid_72x104_T4": (72, 104, 4, 3000, {}, 1, 8),
    "bde2vid_64x96_T5_buf5": (64, 96, 5, 2500, dict(buffer_index=[-2, -1, 0, 1, 2], q_idx=2, depths=[2, 0, 2]), 2, 9),
    "bde2vid_56x80_T3_q0": (56, 80, 3, 2000, dict(buffer_index=[0, 1], q_idx=0, depths=[1, 0, 1]), 3, 10),
}


# architecture variants of the generator constructor (...V5.py:19-98); weights = synth.random_state_dict_like(template, seed)
# name: (H, W, T, N, cfg overrides, weight seed, event seq_id)
VARIANT_CASES = {
    "var_gru_64x96_T4": (64, 96, 4, 2500, dict(recurrent_block_type="convgru", depths=[1, 0, 1]), 21, 41),
    "var_concat_64x96_T3": (64, 96, 3, 2500, dict(skip_type="concat", depths=[1, 0, 1]), 22, 42),
    "var_bn_64x96_T3": (64, 96, 3, 2500, dict(norm="BN", depths=[1, 0, 1]), 23, 43),
    "var_in_56x80_T3": (56, 80, 3, 2000, dict(norm="IN", depths=[1, 0, 1]), 24, 44),
    "var_nwin_64x96_T3": (64, 96, 3, 2500, dict(nwindow_size=(3, 3), depths=[2, 0, 2]), 25, 45),
    "var_tail_64x96_T4": (64, 96, 4, 2500, dict(depths=[1, 0, 0]), 26, 46),
    "var_norc_64x96_T3": (64, 96, 3, 2500, dict(useRC=False, depths=[1, 0, 1]), 27, 47),
    "var_all_64x96_T3": (64, 96, 3, 2500, dict(recurrent_block_type="convgru", skip_type="concat", norm="BN",
                                               nwindow_size=(2, 2), depths=[2, 0, 0], buffer_index=[-1, 0], q_idx=1,
                                               activation=dict(type="Identity")), 28, 48),
}


def gen_cfg(over):
    """Constructor kwargs of the generator for a variant (type + every argument of DEFAULT_CFG)."""
    return O.full_cfg(over)


def voxel_inputs(seq_id, T, H, W, N, num_encoders=3):
    """Voxel grids of a synthetic sequence via the oracle voxeliser, padded like the driver does."""
    ev = synth.gen_events(seq_id, T, H, W, N)
    prm = O.croper_params(W, H, num_encoders)
    out = []
    for w in range(T):
        xs, ys, ts, ps = synth.to_loader_format(ev, w)
        v = torch.from_numpy(O.voxel_grid(xs, ys, ts, ps, 5, (H, W)))[None]
        out.append(O.pad_voxel(v, prm))
    return out, prm


def sub(t):
    """Strided subsample of a [B,C,H,W] tensor to keep fixtures small."""
    return t[:, ::4, ::2, ::2].contiguous().numpy()


def main():
    os.makedirs(GOLD, exist_ok=True)
    R = ref_shim.reference_modules()
    manifest = {}

    # ---- voxeliser -------------------------------------------------------------------
    small = {}
    for name, (H, W, N, T, sid) in {"a": (48, 64, 3000, 2, 3), "b": (31, 45, 700, 3, 4)}.items():
        ev = synth.gen_events(sid, T, H, W, N)
        for w in range(T):
            xs, ys, ts, ps = synth.to_loader_format(ev, w)
            ref = R.events_to_voxel_torch(*(torch.from_numpy(a) for a in (xs, ys, ts, ps)), 5,
                                          sensor_size=(H, W)).numpy()
            mine = O.voxel_grid(xs, ys, ts, ps, 5, (H, W))
            assert np.array_equal(ref, mine)
            small["ref_%s_%d" % (name, w)] = ref
        small["meta_%s" % name] = np.array([H, W, N, T, sid])
    np.savez_compressed(os.path.join(GOLD, "voxel_small.npz"), **small)

    sums = {}
    for (H, W, N, sid) in [(180, 240, 15000, 0), (260, 346, 31500, 1), (720, 1280, 333333, 2)]:
        ev = synth.gen_events(sid, 2, H, W, N)
        for w in range(2):
            xs, ys, ts, ps = synth.to_loader_format(ev, w)
            ref = R.events_to_voxel_torch(*(torch.from_numpy(a) for a in (xs, ys, ts, ps)), 5,
                                          sensor_size=(H, W)).numpy()
            mine = O.voxel_grid(xs, ys, ts, ps, 5, (H, W))
            # torch's CPU index_put_(accumulate=True) changes its summation order for large N
            # (observed at N=333333: <=1 ulp differences in ~70 cells), so bit-equality with the
            # event-order restatement holds only for the smaller windows; record which.
            exact = bool(np.array_equal(ref, mine))
            assert float(np.abs(ref - mine).max()) <= 5e-7
            sums["%dx%d_N%d_seq%d_w%d" % (H, W, N, sid, w)] = dict(
                sha256=hashlib.sha256(ref.tobytes()).hexdigest() if exact else None,
                oracle_bit_equal=exact,
                bin_sums=[float(s) for s in ref.astype(np.float64).sum(axis=(1, 2))],
                abs_sum=float(np.abs(ref).astype(np.float64).sum()),
                nonzero=int(np.count_nonzero(ref)))
    manifest["voxel_checksums"] = sums

    # analytic known answers, confirmed on the reference
    xs = torch.tensor([1., 2., 3., 0.]); ys = torch.tensor([0., 1., 2., 3.])
    ts = torch.tensor([0., 0.3125, 0.5, 1.0]); ps = torch.tensor([1., -1., 1., 1.])
    ref = R.events_to_voxel_torch(xs, ys, ts, ps, 5, sensor_size=(4, 4)).numpy()
    assert ref[0, 0, 1] == 1.0 and ref[4, 3, 0] == 1.0 and ref[:4, 3, 0].sum() == 0
    assert ref[1, 1, 2] == -0.75 and ref[2, 1, 2] == -0.25 and ref[2, 2, 3] == 1.0
    manifest["voxel_known_answers"] = "t_norm=1.25 -> (.75,.25); first event bin0 w=1; last event bin B-1 only"

    # ---- croper ----------------------------------------------------------------------
    crop = {}
    for (w, h, ne) in [(346, 260, 3), (240, 180, 3), (1280, 720, 3), (346, 260, 4), (96, 64, 3), (641, 481, 3)]:
        c = R.Croper(ne)
        c.update_params(w, h)
        crop["%dx%d_e%d" % (w, h, ne)] = dict(Hp=c.height_crop_size, Wp=c.width_crop_size,
                                                pad=[c.padding_left, c.padding_right, c.padding_top, c.padding_bottom],
                                                crop=[c.iy0, c.iy1, c.ix0, c.ix1])
        p = O.croper_params(w, h, ne)
        assert list(p["pad"]) == crop["%dx%d_e%d" % (w, h, ne)]["pad"] and list(p["crop"]) == crop["%dx%d_e%d" % (w, h, ne)]["crop"]
    manifest["croper"] = crop

    # ---- BDE2VID ---------------------------------------------------------------------
    for name, (H, W, T, N, over, wseed, sid) in MODEL_CASES.items():
        cfg = O.full_cfg(over)
        sd = synth.init_state_dict(cfg, wseed, stress=True)
        model = R.BDE2VID(generator=dict(cfg)).eval()
        model.load_state_dict(sd, strict=True)
        vox, prm = voxel_inputs(sid, T, H, W, N)
        with ref_shim.cpu_mode(), torch.no_grad():
            ref = model([{"events": v} for v in vox])
            taps = {}
            mine = O.bde2vid_forward(sd, cfg, vox, taps=taps)
        err = max(float((a - b).abs().max()) for a, b in zip(ref, mine))
        assert err == 0.0, (name, err)
        arrays = {"ref_frames": torch.cat(ref, 0).numpy()}
        arrays["tap_head0"] = sub(taps["head"][0])
        for l in range(cfg["num_encoders"]):
            arrays["tap_merged%d_t1" % l] = sub(taps["merged%d" % l][1])
            arrays["tap_level%d_t1" % l] = sub(taps["level%d" % l][1])
            arrays["tap_level%d_last" % l] = sub(taps["level%d" % l][T - 1])
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
        manifest[name] = dict(H=H, W=W, T=T, N=N, cfg_overrides=over, weight_seed=wseed, seq_id=sid,
                              oracle_vs_reference_maxabs=err,
                              frame_mean=float(torch.cat(ref, 0).mean()))
        print(name, "ok", arrays["ref_frames"].shape, manifest[name]["frame_mean"])

    # ---- BDE2VID architecture variants ---------------------------------------------------
    for name, (H, W, T, N, over, wseed, sid) in VARIANT_CASES.items():
        cfg = gen_cfg(over)
        model = R.BDE2VID(generator=dict(cfg)).eval()
        sd = synth.random_state_dict_like(model.state_dict(), wseed)
        model.load_state_dict(sd, strict=True)
        vox, prm = voxel_inputs(sid, T, H, W, N)
        with ref_shim.cpu_mode(), torch.no_grad():
            ref = model([{"events": v} for v in vox])
            mine = O.bde2vid_forward(sd, cfg, vox)
        err = max(float((a - b).abs().max()) for a, b in zip(ref, mine))
        assert err == 0.0, (name, err)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), ref_frames=torch.cat(ref, 0).numpy())
        manifest[name] = dict(H=H, W=W, T=T, N=N, cfg_overrides={k: (list(v) if isinstance(v, tuple) else v) for k, v in over.items()},
                              weight_seed=wseed, seq_id=sid, oracle_vs_reference_maxabs=err, n_keys=len(sd),
                              keys_sha256=hashlib.sha256("\n".join(sorted(sd.keys())).encode()).hexdigest(),
                              frame_mean=float(torch.cat(ref, 0).mean()))
        print(name, "ok", manifest[name]["frame_mean"])

    # ---- E2VIDRecurrent with ConvGRU, FireNet --------------------------------
    from model.e2vid.model import FireNet
    g = torch.Generator().manual_seed(6)
    e = R.E2VIDRecurrent({"num_bins": 5, "recurrent_block_type": "convgru", "num_encoders": 3}).eval()
    esd = synth.random_state_dict_like(e.state_dict(), 31)
    e.load_state_dict(esd, strict=True)
    xs = [torch.randn(2, 5, 64, 96, generator=g) for _ in range(3)]
    frames, st = [], None
    with torch.no_grad():
        for x in xs:
            r = e({"events": x})["image"]
            a, st = O.e2vid_recurrent_forward(esd, x, st, num_encoders=3)
            assert float((r - a).abs().max()) == 0.0
            frames.append(r)
    np.savez_compressed(os.path.join(GOLD, "e2vid_gru_64x96_B2_T3.npz"), ref_frames=torch.stack(frames).numpy())
    manifest["e2vid_gru_64x96_B2_T3"] = dict(seed=31, input_seed=6, n_keys=len(esd))
    fn = FireNet().eval()
    fsd = synth.random_state_dict_like(fn.state_dict(), 32)
    fn.load_state_dict(fsd, strict=True)
    frames, st = [], None
    with torch.no_grad():
        for x in xs:
            r = fn({"events": x})["image"]
            a, st = O.firenet_forward(fsd, x, st)
            assert float((r - a).abs().max()) == 0.0
            frames.append(r)
    np.savez_compressed(os.path.join(GOLD, "firenet_64x96_B2_T3.npz"), ref_frames=torch.stack(frames).numpy())
    manifest["firenet_64x96_B2_T3"] = dict(seed=32, input_seed=6, n_keys=len(fsd), keys=sorted(fsd.keys()))

    # ---- loader-side transforms: LegacyNorm / RobustNorm / hot-pixel mask -------
    from utils_func.data_augmentation import LegacyNorm
    from utils_func.utils import RobustNorm
    import events_contrast_maximization.utils.event_utils as EU
    ev = synth.gen_events(51, 2, 48, 64, 3000)
    for w in range(2):
        xs_, ys_, ts_, ps_ = synth.to_loader_format(ev, w)
        v = torch.from_numpy(O.voxel_grid(xs_, ys_, ts_, ps_, 5, (48, 64)))
        assert torch.equal(LegacyNorm()(v.clone()), O.legacy_norm(v.clone()))
        for lo, hi in ((0, 95), (1, 99), (5, 50)):
            assert torch.equal(RobustNorm(lo, hi)(v.clone()), O.robust_norm(v.clone(), lo, hi))
    a0, a1 = int(ev["offsets"][0]), int(ev["offsets"][1])
    pm = ev["ps"][a0:a1] * 2.0 - 1.0
    ref_mask = EU.get_hot_event_mask(ev["xs"][a0:a1].astype(np.int64), ev["ys"][a0:a1].astype(np.int64), pm, (48, 64), num_hot=30)
    assert np.array_equal(ref_mask, O.hot_event_mask(ev["xs"][a0:a1], ev["ys"][a0:a1], pm, (48, 64), 30))
    manifest["loader_transforms"] = "LegacyNorm / RobustNorm / get_hot_event_mask: oracle bit-equal to the reference classes"

    # ---- E2VIDRecurrent (config 3 twin) ----------------------------------------------
    torch.manual_seed(0)
    e = R.E2VIDRecurrent({"num_bins": 5}).eval()
    g = torch.Generator().manual_seed(5)
    esd = {k: (torch.rand(v.shape, generator=g) * 2 - 1) / max(1, v[0].numel()) ** 0.5 for k, v in e.state_dict().items()}
    e.load_state_dict(esd, strict=True)
    xs = [torch.randn(2, 5, 64, 96, generator=g) for _ in range(3)]
    frames = []
    st = None
    with torch.no_grad():
        for x in xs:
            r = e({"events": x})["image"]
            a, st = O.e2vid_recurrent_forward(esd, x, st)
            assert float((r - a).abs().max()) == 0.0
            frames.append(r)
    np.savez_compressed(os.path.join(GOLD, "e2vid_64x96_B2_T3.npz"), ref_frames=torch.stack(frames).numpy())
    manifest["e2vid_64x96_B2_T3"] = dict(seed=5, keys=list(esd.keys()), shapes=[list(v.shape) for v in esd.values()])

    manifest["torch"] = torch.__version__
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
