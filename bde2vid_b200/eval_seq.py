"""Headless evaluation driver: the loop of the reference's ``eval_models_seq.py`` (:99-144 eval_model_alldata, :147-282
eval_model) without the GUI / LPIPS parts, on the fused events -> frames path.

What is kept from the reference driver
  * checkpoint loading: ``Config.fromstring(meta.cfg).model`` -> ``MODELS.build`` -> strict ``load_state_dict``
    (eval_models_seq.py:41-60, :86) via ``model.load_checkpoint``;
  * per file: one window of events per ground-truth frame ("between_frames": window i = events [event_idx[i-1],
    event_idx[i]), h5_dataset.py:448-455), fewer than 3 events -> zero voxel grid (:219-221), optional LegacyNorm
    (``args.normalize``, eval_models_seq.py:159-161) and hot-pixel filter (``filter_hot_events``, h5_dataset.py:163-172);
    ``Croper(num_encoders)`` pad / crop (:197-205, :242); the sequence is cut into chunks of ``subseq_L`` frames (1000) and
    every chunk is one ``model(...)`` call with freshly reset states (:216-219, bde2vid.py:31); ``max_length``
    (the "pause" robustness experiment, :184-189, is not reproduced);
  * per frame MSE and SSIM against ``frame / 255`` (:253-258, evaluate/metrics.py:42-65), per-file means (:278-282), the
    result JSON ``{dataset: {file: {metric: mean}}}`` plus the ``_detail`` JSON of per-frame values (:137-144) and the
    "skip if the result file exists" resume rule (:110-121).
What differs: voxelisation happens on the GPU from the raw event arrays (no DataLoader workers, no per-window H2D of dense
grids), metrics are computed on the device (2 doubles per frame leave the GPU), files are sharded over the ranks of a
``torch.distributed`` job (one process per GPU) and the per-file rows are gathered on rank 0.  LPIPS ('p_loss') needs
AlexNet weights that are not available offline and is not computed.

Sequence files: ``.npz`` with ``xs`` int16, ``ys`` int16, ``ts`` float64, ``ps`` bool (events_contrast_maximization/tools/
event_packagers.py:44-47), ``event_idx`` int64 [T] (the ``event_idx`` attribute of every image, h5_dataset.py:448-455),
``frames`` uint8 [T, H, W]; or the same layout in an ``.h5`` file when h5py is importable (it is not in this image).
"""
import argparse
import collections
import glob
import json
import os

import numpy as np
import torch

from . import metrics as M
from . import ops
from .croper import Croper
from .dist import shard_units


def load_sequence(path):
    """-> dict(xs, ys, ts, ps, event_idx, frames, sensor_resolution) of numpy arrays in the on-disk dtypes."""
    if path.endswith(".npz"):
        z = np.load(path)
        seq = {k: z[k] for k in ("xs", "ys", "ts", "ps", "event_idx", "frames")}
    elif path.endswith(".h5"):
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("reading %s needs h5py, which is not installed; convert the file to the .npz layout "
                               "(see bde2vid_b200/eval_seq.py)" % path) from e
        with h5py.File(path, "r") as f:
            names = sorted(f["images"].keys())
            seq = dict(xs=f["events/xs"][:], ys=f["events/ys"][:], ts=f["events/ts"][:], ps=f["events/ps"][:],
                       event_idx=np.array([f["images"][n].attrs["event_idx"] for n in names], dtype=np.int64),
                       frames=np.stack([f["images"][n][:] for n in names], 0))
    else:
        raise ValueError("unknown sequence file type: %s" % path)
    seq["xs"] = np.ascontiguousarray(seq["xs"], dtype=np.int16)
    seq["ys"] = np.ascontiguousarray(seq["ys"], dtype=np.int16)
    seq["ts"] = np.ascontiguousarray(seq["ts"], dtype=np.float64)
    seq["ps"] = np.ascontiguousarray(seq["ps"]).astype(bool)
    seq["event_idx"] = np.asarray(seq["event_idx"], dtype=np.int64)
    seq["sensor_resolution"] = tuple(int(v) for v in seq["frames"].shape[-2:])
    return seq


def frame_windows(event_idx):
    """compute_frame_indices (h5_dataset.py:448-455): window i = [end of window i-1, event_idx[i]) as CSR offsets [T+1]."""
    return np.concatenate([np.zeros(1, np.int64), np.asarray(event_idx, np.int64)])


@torch.no_grad()
def eval_sequence(model, seq, device, subseq_L=1000, normalize=False, filter_hot_events=False, max_length=None,
                  data_range=M.REFERENCE_DATA_RANGE, return_frames=False):
    """One file: events -> frames -> per-frame metrics (eval_models_seq.py:147-282).  Returns (result, detail) with
    result = {'mse': mean, 'ssim': mean} and detail = {'mse': [...], 'ssim': [...]} (+ the frames when asked)."""
    H, W = seq["sensor_resolution"]
    off = frame_windows(seq["event_idx"])
    T = len(off) - 1
    if max_length is not None:
        T = min(T, int(max_length))
    off = off[:T + 1].copy()
    n_ev = int(off[-1])
    dev_ev = [torch.from_numpy(seq[k][:n_ev]).to(device) for k in ("xs", "ys", "ts")] + \
             [torch.from_numpy(seq["ps"][:n_ev].view(np.uint8)).to(device)]
    hot = None
    if filter_hot_events:
        # h5_dataset.py:163-169: mask from the events of the first 0.2 s, num_hot = 1 % of the pixels
        t0 = seq["ts"][0]
        hot_num = min(int(np.searchsorted(seq["ts"], t0 + 0.2, side="left")), len(seq["ts"]))
        hx, hy = torch.from_numpy(seq["xs"][:hot_num]).to(device), torch.from_numpy(seq["ys"][:hot_num]).to(device)
        hp = torch.from_numpy(seq["ps"][:hot_num].view(np.uint8)).to(device)
        hot = ops.hot_pixel_mask(hx, hy, hp, H, W, int(H * W * 0.01))
    try:
        num_encoders = model.num_encoders
    except AttributeError:
        num_encoders = 3                                        # eval_models_seq.py:197-201
    crop = Croper(num_encoders)
    crop.update_params(W, H)
    gts = torch.from_numpy(np.ascontiguousarray(seq["frames"][:T])).to(device).float() / 255    # transform_frame, h5_dataset.py:372
    L = T if subseq_L is None else int(subseq_L)
    per_frame, frames_out = [], []
    for t0 in range(0, T, L):
        t1 = min(T, t0 + L)
        a, b = int(off[t0]), int(off[t1])
        sub_off = torch.from_numpy(off[t0:t1 + 1] - off[t0]).to(device)
        chunk = [e[a:b] for e in dev_ev]
        if b == a:                                              # a chunk without any event: keep the arrays non-empty
            chunk = [torch.zeros(1, dtype=e.dtype, device=device) for e in dev_ev]
            sub_off = torch.zeros(t1 - t0 + 1, dtype=torch.int64, device=device)
        frames = model.reconstruct_events(*chunk, sub_off, (H, W), num_encoders=num_encoders,
                                          normalize="legacy" if normalize else None, hot_mask=hot)
        pred = torch.cat(frames, 0).reshape(t1 - t0, H, W)
        per_frame.append(ops.frame_metrics(pred.contiguous(), gts[t0:t1].contiguous(), 0, 0, data_range))
        if return_frames:
            frames_out.append(pred)
    pf = torch.cat(per_frame, 0).cpu()                          # [T, 2] float64: the only D2H of the evaluation
    detail = {"mse": pf[:, 0].tolist(), "ssim": pf[:, 1].tolist()}
    result = {k: float(sum(v) / max(1, len(v))) for k, v in detail.items()}
    if return_frames:
        return result, detail, torch.cat(frames_out, 0)
    return result, detail


def eval_files(datafiles, eval_fn, costs=None):
    """Shard ``datafiles`` over the ranks of the (optional) process group, run ``eval_fn(datafile) -> (result, detail)`` on
    this rank's share and gather every row on all ranks.  Returns (results, detail_results) in the reference's nesting
    ``{dataset: {file: ...}}`` (eval_models_seq.py:122-135) plus the frame-weighted overall means (one all-reduce)."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank, world = (dist.get_rank(), dist.get_world_size()) if distributed else (0, 1)
    mine = shard_units(costs if costs is not None else [1.0] * len(datafiles), world)[rank]
    rows = []
    for i in mine:
        result, detail = eval_fn(datafiles[i])
        rows.append((i, result, detail))
    if distributed:
        gathered = [None] * world
        dist.all_gather_object(gathered, rows)
        rows = [r for part in gathered for r in part]
    rows.sort(key=lambda r: r[0])
    results = collections.defaultdict(dict)
    details = collections.defaultdict(dict)
    sums = collections.defaultdict(float)
    for i, result, detail in rows:
        dataset, fname = os.path.split(datafiles[i])
        dataset = os.path.basename(dataset) or "."
        fname = fname.rsplit(".", 1)[0]
        results[dataset][fname] = result
        details[dataset][fname] = detail
    # overall per-frame means: metric sums of THIS rank's files, one all-reduce (the path's only collective)
    for i, result, detail in rows:
        if i in mine:
            for k, v in detail.items():
                sums[k] += float(sum(v))
            sums["n"] += float(len(next(iter(detail.values())))) if detail else 0.0
    overall = M.finalize_means(M.reduce_metric_sums(dict(sums))) if sums else {}
    return dict(results), dict(details), overall


def eval_model_alldata(datafiles, checkpoint_file, data_dir, out_dir, device, subseq_L=1000, normalize=False,
                       filter_hot_events=False, max_length=None, datatype="data"):
    """eval_models_seq.py:99-144 for one checkpoint: result / detail JSON files named like the reference's, written by
    rank 0; returns the result dict (None when the result file already exists: the reference's resume rule)."""
    import torch.distributed as dist
    from .model import load_checkpoint
    name = os.path.split(checkpoint_file)[-1].split(".")[0]
    result_file = "%s_L%s_%s.txt" % (name, subseq_L, datatype) if subseq_L is not None else "%s_%s.txt" % (name, datatype)
    result_file = os.path.join(out_dir, result_file)
    if os.path.exists(result_file):
        print("skiping %s" % checkpoint_file)
        return None
    model = load_checkpoint(checkpoint_file, device=device)

    def one(datafile):
        seq = load_sequence(os.path.join(data_dir, datafile))
        return eval_sequence(model, seq, device, subseq_L, normalize, filter_hot_events, max_length)

    results, details, overall = eval_files(datafiles, one)
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
        with open(result_file, "w") as fp:
            json.dump(results, fp)
        with open(result_file.replace(".txt", "_detail.txt"), "w") as fp:
            json.dump(details, fp)
        with open(result_file.replace(".txt", "_overall.txt"), "w") as fp:
            json.dump(overall, fp)
        print("results writed to %s" % result_file)
    return results


def main(argv=None):
    """``python -m bde2vid_b200.eval_seq --weights_dir W --data_dir D`` (the reference's flags, eval_models_seq.py:293-298);
    under torchrun every rank takes a share of the files."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights_dir", required=True)
    ap.add_argument("--data_dir", required=True, help="directory with <dataset>/<file>.npz (and optionally eval_data.txt)")
    ap.add_argument("--output_dir", default=None)
    ap.add_argument("--datatype", default="all", help="dataset name filter, as args.datatype of the reference")
    ap.add_argument("--subseq_L", type=int, default=1000)
    ap.add_argument("--normalize", action="store_true")
    ap.add_argument("--filter_hot_events", action="store_true")
    args = ap.parse_args(argv)
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    listing = os.path.join(args.data_dir, "eval_data.txt")
    if os.path.exists(listing):
        with open(listing) as f:
            files = [ln.strip() for ln in f if ln.strip()]
    else:
        files = sorted(os.path.relpath(p, args.data_dir) for p in glob.glob(os.path.join(args.data_dir, "*", "*.npz")))
    if args.datatype != "all":
        files = [f for f in files if f.split(os.sep)[0] == args.datatype]
    out_dir = args.output_dir or args.weights_dir
    for ckpt in sorted(glob.glob(os.path.join(args.weights_dir, "*.pth"))):
        eval_model_alldata(files, ckpt, args.data_dir, out_dir, device, args.subseq_L, args.normalize,
                           args.filter_hot_events, None, args.datatype)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
