"""Evaluation metrics on the device + the path's only collective.

The reference computes MSE / SSIM / LPIPS per frame on the host (eval_models_seq.py:242-258 -> evaluate/metrics.py) and
averages them per file (:278-282).  Here one kernel launch produces (mse, ssim) for every frame of a sequence, the sums
stay on the GPU until a single all-reduce across the ranks that share the evaluation.  LPIPS needs AlexNet weights that
are not available offline and stays out of scope.

``data_range``: the reference calls skimage ``structural_similarity(y_input, y_target)`` on float32 arrays WITHOUT a
``data_range`` (evaluate/metrics.py:59-63).  scikit-image < 0.22 then takes it from the dtype: for floating-point images
``dmin, dmax = dtype_range[float32] = (-1, 1)`` so data_range = 2 (C1 = (0.01*2)^2, C2 = (0.03*2)^2); scikit-image >= 0.22
refuses float inputs without an explicit data_range.  The reference-equivalent default is therefore 2.0
(``REFERENCE_DATA_RANGE``); pass 1.0 for the conventional [0, 1]-image SSIM.
"""
import torch

from . import ops
from .dist import finalize_means, reduce_metric_sums

REFERENCE_DATA_RANGE = 2.0


def sequence_metric_sums(frames, gts, croper, data_range=REFERENCE_DATA_RANGE):
    """frames: list of T tensors [1, 1, Hp, Wp] (model output, padded) or one tensor [T, Hp, Wp]; gts: [T, H, W] float32.
    Returns {'mse': sum, 'ssim': sum, 'n': T} for this sequence (python floats; one device -> host read)."""
    if isinstance(frames, (list, tuple)):
        pred = torch.stack([f.reshape(f.shape[-2], f.shape[-1]) for f in frames], 0)
    else:
        pred = frames
    pred = pred.to(torch.float32).contiguous()
    gts = gts.to(device=pred.device, dtype=torch.float32).contiguous()
    T, Hp, Wp = pred.shape
    H, W = gts.shape[-2:]
    y0 = Hp // 2 - H // 2        # Croper.crop: rows [cy - floor(H/2), cy + ceil(H/2))  (inference_utils.py:26-32,112-114)
    x0 = Wp // 2 - W // 2
    if croper is not None:
        assert (croper.height_crop_size, croper.width_crop_size) == (Hp, Wp)
    per_frame = ops.frame_metrics(pred, gts.reshape(T, H, W), y0, x0, data_range)
    s = per_frame.sum(0).tolist()
    return {"mse": s[0], "ssim": s[1], "n": float(T)}, per_frame


def evaluate(sequences, data_range=REFERENCE_DATA_RANGE):
    """sequences: iterable of (frames, gts, croper) owned by THIS rank.  Returns the per-frame means over all ranks
    (one all-reduce of the metric sums; eval_models_seq.py:278-282)."""
    total = {"mse": 0.0, "ssim": 0.0, "n": 0.0}
    for frames, gts, croper in sequences:
        s, _ = sequence_metric_sums(frames, gts, croper, data_range)
        for k in total:
            total[k] += s[k]
    return finalize_means(reduce_metric_sums(total))
