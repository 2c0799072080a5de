"""CPU, build container only: the oracle port against the live, unmodified reference."""
import numpy as np
import pytest
import torch

from bde2vid_b200 import synth
from oracle import oracle_torch as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


def test_voxel_vs_reference():
    R = ref_shim.reference_modules()
    ev = synth.gen_events(11, 2, 90, 120, 4000)
    for w in range(2):
        xs, ys, ts, ps = synth.to_loader_format(ev, w)
        ref = R.events_to_voxel_torch(*(torch.from_numpy(a) for a in (xs, ys, ts, ps)), 5, sensor_size=(90, 120)).numpy()
        assert np.array_equal(ref, O.voxel_grid(xs, ys, ts, ps, 5, (90, 120)))


def test_bde2vid_vs_reference_and_strict_load():
    R = ref_shim.reference_modules()
    cfg = O.full_cfg()
    sd = synth.init_state_dict(cfg, 5, stress=True)
    model = R.BDE2VID(generator=dict(cfg)).eval()
    model.load_state_dict(sd, strict=True)       # our synthetic checkpoints carry the reference's exact key set
    g = torch.Generator().manual_seed(3)
    vox = [torch.randn(1, 5, 56, 64, generator=g) for _ in range(3)]
    with ref_shim.cpu_mode(), torch.no_grad():
        ref = model([{"events": v} for v in vox])
        mine = O.bde2vid_forward(sd, cfg, vox)
    assert max(float((a - b).abs().max()) for a, b in zip(ref, mine)) == 0.0
