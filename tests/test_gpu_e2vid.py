"""GPU parity: drop-in E2VIDRecurrent (config 3: ConvLSTM recurrent UNet sharing the BDE2VID conv/gate kernels)
against the reference's committed frames (tests/golden/e2vid_64x96_B2_T3.npz) and the oracle port."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle_torch as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"fp32": 2e-4, "bf16": 2e-3}


def golden_state_dict(rec):
    gen = torch.Generator().manual_seed(rec["seed"])
    sd = {}
    for k, shape in zip(rec["keys"], rec["shapes"]):
        v = torch.rand(shape, generator=gen) * 2 - 1
        sd[k] = v / max(1, v[0].numel()) ** 0.5
    return sd, gen


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_e2vid_golden_frames(manifest, precision):
    from bde2vid_b200.e2vid import E2VIDRecurrent
    g = load_golden("e2vid_64x96_B2_T3")
    sd, gen = golden_state_dict(manifest["e2vid_64x96_B2_T3"])
    xs = [torch.randn(2, 5, 64, 96, generator=gen) for _ in range(3)]
    model = E2VIDRecurrent({"num_bins": 5})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    model.unetrecurrent.precision = precision
    model.reset_states()
    with torch.no_grad():
        for i, x in enumerate(xs):
            img = model({"events": x.to(DEV)})["image"]
            assert img.shape == (2, 1, 64, 96) and img.dtype == torch.float32
            err = np.abs(img.cpu().numpy() - g["ref_frames"][i]).max()
            print("e2vid step", i, precision, "max-abs", float(err))
            assert err <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_e2vid_states_and_reset(manifest, precision):
    """UNetRecurrent.forward(x, prev_states) -> (img, states): explicit state passing (unet.py:167-200), states
    accepted back in either our layout or the reference's (h, c) NCHW tuples, reset_states() restarts from zeros."""
    from bde2vid_b200.e2vid import E2VIDRecurrent
    sd, gen = golden_state_dict(manifest["e2vid_64x96_B2_T3"])
    model = E2VIDRecurrent({"num_bins": 5})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    model.unetrecurrent.precision = precision
    xs = [torch.randn(1, 5, 48, 80, generator=gen) for _ in range(4)]
    with torch.no_grad():
        st_ref, st = None, None
        for i, x in enumerate(xs):
            ref, st_ref = O.e2vid_recurrent_forward(sd, x, st_ref)
            if i == 2:   # hand the states back as reference-layout tensors
                st = [(s.h, s.c) for s in st]
            img, st = model.unetrecurrent(x.to(DEV), st)
            assert float((img.cpu() - ref).abs().max()) <= TOL[precision]
            for a, b in zip(st, st_ref):
                h, c = a
                assert h.shape == b[0].shape and c.shape == b[1].shape
                assert float((c.cpu() - b[1]).abs().max()) <= (1e-4 if precision == "fp32" else 3e-2)
        model.reset_states()
        a = model({"events": xs[0].to(DEV)})["image"]
        model.reset_states()
        b = model({"events": xs[0].to(DEV)})["image"]
        assert torch.equal(a, b)
        assert float((a.cpu() - O.e2vid_recurrent_forward(sd, xs[0], None)[0]).abs().max()) <= TOL[precision]


def test_e2vid_rejects_cpu_and_unsupported():
    from bde2vid_b200.e2vid import E2VIDRecurrent
    with pytest.raises(NotImplementedError):
        E2VIDRecurrent({"num_bins": 5, "norm": "BN"})
    m = E2VIDRecurrent({"num_bins": 5}).eval()
    with pytest.raises(RuntimeError):
        m({"events": torch.zeros(1, 5, 32, 32)})


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_e2vid_convgru_golden_frames(manifest, precision):
    """E2VIDRecurrent with recurrent_block_type='convgru' (model/e2vid/model.py:89-92): ConvGRU as two fused convolutions
    (BDE_EPI_GRU_UR / BDE_EPI_GRU_OUT) against the frames of the unmodified reference; states round-trip in the reference's
    layout (a single NCHW tensor per level)."""
    from bde2vid_b200 import synth
    from bde2vid_b200.e2vid import E2VIDRecurrent
    rec = manifest["e2vid_gru_64x96_B2_T3"]
    g = load_golden("e2vid_gru_64x96_B2_T3")
    gen = torch.Generator().manual_seed(rec["input_seed"])
    xs = [torch.randn(2, 5, 64, 96, generator=gen) for _ in range(3)]
    model = E2VIDRecurrent({"num_bins": 5, "recurrent_block_type": "convgru", "num_encoders": 3})
    sd = synth.random_state_dict_like(model.state_dict(), rec["seed"])
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    model.unetrecurrent.precision = precision
    st = None
    with torch.no_grad():
        for i, x in enumerate(xs):
            if i == 2:
                st = [s.h for s in st]                      # hand the states back as reference-layout tensors
            img, st = model.unetrecurrent(x.to(DEV), st)
            err = np.abs(img.cpu().numpy() - g["ref_frames"][i]).max()
            print("e2vid gru step", i, precision, "max-abs", float(err))
            assert err <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_firenet_golden_frames(manifest, precision):
    """FireNet (model/e2vid/model.py:119-172) on the same kernels against the reference's frames."""
    from bde2vid_b200 import synth
    from bde2vid_b200.e2vid import FireNet
    rec = manifest["firenet_64x96_B2_T3"]
    g = load_golden("firenet_64x96_B2_T3")
    gen = torch.Generator().manual_seed(rec["input_seed"])
    xs = [torch.randn(2, 5, 64, 96, generator=gen) for _ in range(3)]
    model = FireNet()
    sd = synth.random_state_dict_like(model.state_dict(), rec["seed"])
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(DEV)
    model.precision = precision
    model.reset_states()
    with torch.no_grad():
        for i, x in enumerate(xs):
            img = model({"events": x.to(DEV)})["image"]
            assert img.shape == (2, 1, 64, 96)
            err = np.abs(img.cpu().numpy() - g["ref_frames"][i]).max()
            scale = max(1.0, float(np.abs(g["ref_frames"][i]).max()))
            print("firenet step", i, precision, "max-abs", float(err), "scale", scale)
            assert err <= (2e-4 if precision == "fp32" else 1e-2) * scale     # raw (no sigmoid) output
        model.reset_states()
        again = model({"events": xs[0].to(DEV)})["image"]
        assert np.abs(again.cpu().numpy() - g["ref_frames"][0]).max() <= (2e-4 if precision == "fp32" else 1e-2) * scale
