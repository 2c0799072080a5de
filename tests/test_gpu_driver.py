"""GPU: the headless driver (bde2vid_b200/eval_seq.py = eval_models_seq.py:99-282 without GUI / LPIPS) on a synthetic
sequence file and a synthetic mmengine-style checkpoint, against the oracle run of the same flow."""
import json
import os

import numpy as np
import pytest
import torch

from bde2vid_b200 import synth
from oracle import oracle_torch as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

CFG_STR = """model = dict(
    type='BDE2VID',
    generator=dict(
        type='BDE2VIDCrossscalePropogationV5',
        num_bins=5, basechannels=32, num_encoders=3, ks=5, num_res_blocks=2, norm=None,
        recurrent_block_type='convlstm', useRC=True, skip_type='sum',
        buffer_index=[-1, 0, 1], q_idx=1, window_size=(7, 7), nwindow_size=None,
        depths=[1, 0, 1], num_heads=16, losses=[]))
"""


def _write_inputs(tmp_path, T=7, H=60, W=90, N=2500):
    ns = {}
    exec(CFG_STR, ns)
    cfg = O.full_cfg(ns["model"]["generator"])
    sd = synth.init_state_dict(cfg, 3, stress=True)
    wdir, ddir = tmp_path / "weights", tmp_path / "data"
    (ddir / "SYN").mkdir(parents=True)
    wdir.mkdir()
    torch.save({"meta": {"cfg": CFG_STR}, "state_dict": sd}, str(wdir / "epoch_1.pth"))
    ev = synth.gen_events(17, T, H, W, N)
    off = ev["offsets"].copy()
    off[3] = off[2] + 2                       # a 2-event window -> zero voxel grid (loader contract)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (T, H, W), dtype=np.uint8)
    np.savez(str(ddir / "SYN" / "seq0.npz"), xs=ev["xs"], ys=ev["ys"], ts=ev["ts"], ps=ev["ps"], event_idx=off[1:], frames=frames)
    return cfg, sd, dict(ev, offsets=off), frames, str(wdir), str(ddir)


def _oracle_metrics(cfg, sd, ev, frames, chunks, H, W, legacy=False):
    """The reference flow on the CPU: loader voxels (+ LegacyNorm) -> pad -> model per chunk -> crop -> mse / ssim."""
    prm = O.croper_params(W, H, 3)
    off = ev["offsets"]
    T = len(off) - 1
    vox = []
    for w in range(T):
        a, b = int(off[w]), int(off[w + 1])
        g = torch.from_numpy(O.loader_voxel(ev["xs"][a:b], ev["ys"][a:b], ev["ts"][a:b], ev["ps"][a:b], 5, (H, W)))
        if legacy:
            g = O.legacy_norm(g)
        vox.append(O.pad_voxel(g[None], prm))
    preds = []
    with torch.no_grad():
        for t0 in range(0, T, chunks):
            preds += O.bde2vid_forward(sd, cfg, vox[t0:t0 + chunks])
    preds = [O.crop_image(p, prm).reshape(H, W) for p in preds]
    gt = torch.from_numpy(frames).float() / 255
    mse = [O.mse(p, g) for p, g in zip(preds, gt)]
    ssim = [O.ssim_uniform7(p, g, data_range=2.0) for p, g in zip(preds, gt)]
    return mse, ssim, torch.stack(preds, 0)


@pytest.mark.parametrize("legacy", [False, True])
def test_eval_sequence_matches_reference_flow(tmp_path, legacy):
    from bde2vid_b200 import eval_seq
    from bde2vid_b200.model import load_checkpoint
    H, W = 60, 90
    cfg, sd, ev, frames, wdir, ddir = _write_inputs(tmp_path)
    model = load_checkpoint(os.path.join(wdir, "epoch_1.pth"), device=DEV)     # eval_models_seq.py:41-60,:86
    seq = eval_seq.load_sequence(os.path.join(ddir, "SYN", "seq0.npz"))
    result, detail, got = eval_seq.eval_sequence(model, seq, torch.device(DEV), subseq_L=4, normalize=legacy, return_frames=True)
    mse, ssim, ref = _oracle_metrics(cfg, sd, ev, frames, 4, H, W, legacy)
    assert float((got.cpu() - ref).abs().max()) <= 2e-3
    assert np.abs(np.array(detail["mse"]) - np.array(mse)).max() <= 1e-3
    assert np.abs(np.array(detail["ssim"]) - np.array(ssim)).max() <= 1e-3
    assert abs(result["mse"] - float(np.mean(mse))) <= 1e-3 and abs(result["ssim"] - float(np.mean(ssim))) <= 1e-3


def test_eval_model_alldata_writes_reference_schema(tmp_path):
    from bde2vid_b200 import eval_seq
    cfg, sd, ev, frames, wdir, ddir = _write_inputs(tmp_path, T=5)
    out = str(tmp_path / "out")
    res = eval_seq.eval_model_alldata(["SYN/seq0.npz"], os.path.join(wdir, "epoch_1.pth"), ddir, out, torch.device(DEV),
                                      subseq_L=1000, datatype="SYN")
    assert set(res) == {"SYN"} and set(res["SYN"]["seq0"]) == {"mse", "ssim"}
    rf = os.path.join(out, "epoch_1_L1000_SYN.txt")
    with open(rf) as f:
        assert json.load(f) == res
    with open(rf.replace(".txt", "_detail.txt")) as f:
        d = json.load(f)
    assert len(d["SYN"]["seq0"]["mse"]) == 5
    # resume rule (eval_models_seq.py:110-112): an existing result file skips the checkpoint
    assert eval_seq.eval_model_alldata(["SYN/seq0.npz"], os.path.join(wdir, "epoch_1.pth"), ddir, out, torch.device(DEV),
                                       subseq_L=1000, datatype="SYN") is None
