"""Reference-facing voxeliser API on top of the CUDA kernels.

``events_to_voxel_torch`` keeps the signature and return layout of
events_contrast_maximization/utils/event_utils.py:466-509 (one window -> float32 [B, H, W]);
``voxelize_sequence`` is the batched form the fused path uses (CSR windows -> padded grids).
"""
import torch

from . import ops


def _as_cuda_f32(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()


def events_to_voxel_torch(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240), temporal_bilinear=True, check_bounds=True):
    """Turn one window of events into a voxel grid with temporal bilinear interpolation.

    Same arguments as the reference.  ``device`` (or the events' device) must be a CUDA device:
    this implementation has no CPU path.  Events outside the sensor raise IndexError like the
    reference's ``index_put_`` does; that check reads one counter back from the device (a stream
    synchronisation per call) -- ``check_bounds=False`` (an extension) skips it and drops such events."""
    if not temporal_bilinear:
        raise NotImplementedError("temporal_bilinear=False is broken in the reference itself "
                                  "(event_utils.py:500-503 uses undefined names) and is not provided")
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    if device is None:
        device = xs.device if torch.is_tensor(xs) else torch.device("cuda")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("bde2vid_b200.events_to_voxel_torch runs on CUDA only (no CPU fallback)")
    xs, ys, ts, ps = (_as_cuda_f32(t, device) for t in (xs, ys, ts, ps))
    n = xs.numel()
    if n == 0:
        raise IndexError("events_to_voxel_torch: empty event window (the reference fails on ts[-1] too)")
    offsets = torch.tensor([0, n], dtype=torch.int64, device=device)
    H, W = sensor_size
    if not check_bounds:
        return ops.voxelize_seq(xs, ys, ts, ps, offsets, B, H, W)[0]
    oob = torch.zeros(1, dtype=torch.int32, device=device)
    out = ops.voxelize_seq(xs, ys, ts, ps, offsets, B, H, W, oob_count=oob)
    n_oob = int(oob.item())
    if n_oob != 0:
        raise IndexError("events_to_voxel_torch: %d events outside the %dx%d sensor" % (n_oob, H, W))
    return out[0]


def voxelize_sequence(xs, ys, ts, ps, offsets, num_bins, sensor_size, crop=None, algo=0, out=None, oob_count=None, min_events=1):
    """All windows of a sequence in one launch.  ``crop`` (a ``Croper`` with params set) selects the
    zero-padded output geometry; returns float32 [T, num_bins, Hp, Wp]."""
    H, W = sensor_size
    if crop is None:
        return ops.voxelize_seq(xs, ys, ts, ps, offsets, num_bins, H, W, out=out, oob_count=oob_count, algo=algo, min_events=min_events)
    return ops.voxelize_seq(xs, ys, ts, ps, offsets, num_bins, H, W, crop.padding_top, crop.padding_left,
                            crop.height_crop_size, crop.width_crop_size, out=out, oob_count=oob_count, algo=algo, min_events=min_events)
