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


@pytest.mark.parametrize("over", [dict(recurrent_block_type="convgru", skip_type="concat", depths=[1, 0, 0], num_res_blocks=1),
                                  dict(norm="BN", nwindow_size=(2, 3), depths=[1, 0, 1], useRC=False)])
def test_variants_vs_reference_live(over):
    """Fresh variant cfgs (not the ones in tests/golden): strict load of our container's keys into the reference model,
    oracle bit-equal to the reference forward."""
    from bde2vid_b200.model import BDE2VID
    R = ref_shim.reference_modules()
    cfg = O.full_cfg(over)
    ours = BDE2VID(generator=dict(cfg))
    sd = synth.random_state_dict_like(ours.state_dict(), 77)
    model = R.BDE2VID(generator=dict(cfg)).eval()
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(4)
    vox = [torch.randn(1, 5, 56, 64, generator=g) for _ in range(3)]
    with ref_shim.cpu_mode(), torch.no_grad():
        ref = model([{"events": v} for v in vox])
        mine = O.bde2vid_forward(sd, cfg, vox)
    assert max(float((a - b).abs().max()) for a, b in zip(ref, mine)) == 0.0


def test_loader_transforms_vs_reference_live():
    ref_shim.install()
    from utils_func.data_augmentation import LegacyNorm
    from utils_func.utils import RobustNorm
    ev = synth.gen_events(61, 1, 40, 56, 2500)
    v = torch.from_numpy(O.voxel_grid(*synth.to_loader_format(ev, 0), 5, (40, 56)))
    assert torch.equal(LegacyNorm()(v.clone()), O.legacy_norm(v.clone()))
    assert torch.equal(RobustNorm(2, 98)(v.clone()), O.robust_norm(v.clone(), 2, 98))
