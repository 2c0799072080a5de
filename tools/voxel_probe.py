"""Timing of the voxeliser algorithms: uniform-random and spatially CLUSTERED synthetic streams (moving Gaussian blobs -- real event
streams concentrate on edges), the bench shape (346x260, 100 windows x 31 500 events) and the Gen4 shape (1280x720, 64 windows x
333 333 events), loader-format (16 B / event) and raw on-disk-format (13 B / event) ingest.  `--ncu-shape` runs only the 720x1280
launch (for an ncu --set full capture of its DRAM traffic)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bde2vid_b200 import ops, synth  # noqa: E402

DEV = "cuda"


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_shape(H, W, NEV, T, Hp, Wp, pt, pl, algos, streams=("uniform", "clustered")):
    for kind in streams:
        ev = (synth.gen_events if kind == "uniform" else synth.gen_events_clustered)(0, T, H, W, NEV)
        f32 = [torch.from_numpy(a).to(DEV) for a in synth.to_loader_format_seq(ev)]
        raw = [torch.from_numpy(np.ascontiguousarray(ev[k])).to(DEV) for k in ("xs", "ys", "ts")] + \
              [torch.from_numpy(ev["ps"].view(np.uint8)).to(DEV), torch.from_numpy(ev["offsets"]).to(DEV)]
        out = torch.empty(T, 5, Hp, Wp, device=DEV)
        ref = None
        for name, algo, env in algos:
            os.environ.pop("BDE2VID_VOXEL_SCANALL", None)
            os.environ.update(env)
            for fmt, args, bpe in (("f32", f32, 16), ("raw", raw, 13)):
                if fmt == "raw" and algo not in (2, 5):
                    continue
                if fmt == "f32":
                    fn = lambda: ops.voxelize_seq(*args, 5, H, W, pt, pl, Hp, Wp, out=out, algo=algo, min_events=3)  # noqa: E731
                else:
                    fn = lambda: ops.voxelize_raw(*args, 5, H, W, pt, pl, Hp, Wp, out=out, algo=algo, min_events=3)  # noqa: E731
                ms = timed(fn)
                if ref is None:
                    ref = out.clone()
                err = float((out - ref).abs().max())
                gbs = (bpe * NEV + 4 * 5 * H * W) * T / ms / 1e6
                print("%dx%d T=%d %-9s %-40s %-3s %.3f ms  %5.0f GB/s algorithmic  max |diff| %.1e" % (W, H, T, kind, name, fmt, ms, gbs, err))


if "--ncu-shape" in sys.argv:
    for (H, W, N, T, Hp, Wp, pt, pl) in ((260, 346, 31500, 100, 264, 352, 2, 3), (720, 1280, 333333, 64, 720, 1280, 0, 0)):
        ev = synth.gen_events(0, T, H, W, N)
        f32 = [torch.from_numpy(a).to(DEV) for a in synth.to_loader_format_seq(ev)]
        out = torch.empty(T, 5, Hp, Wp, device=DEV)
        for _ in range(2):
            ops.voxelize_seq(*f32, 5, H, W, pt, pl, Hp, Wp, out=out, min_events=3)
        torch.cuda.synchronize()
    sys.exit(0)

if "--pipeline2" in sys.argv:
    base = {"BDE2VID_VOXEL_ZERO_IN_KERNEL": "1", "BDE2VID_PDL": "1"}
    mk = lambda mb, cps: ("zero in kernel, %d MB, %d CTAs/SM" % (mb, cps), 2, dict(base, BDE2VID_VOXEL_CHUNK_MB=str(mb), BDE2VID_VOXEL_CTAS_PER_SM=str(cps)))  # noqa: E731
    run_shape(260, 346, 31500, 100, 264, 352, 2, 3, [mk(mb, cps) for mb in (28, 32, 36, 40) for cps in (2, 3, 4, 5, 6)], streams=("uniform",))
    run_shape(720, 1280, 333333, 64, 720, 1280, 0, 0, [mk(mb, cps) for mb in (20, 40, 60) for cps in (4, 6, 8, 12, 16)], streams=("uniform",))
    sys.exit(0)
if "--pipeline" in sys.argv:
    # chunk pipeline of the default algorithm: in-kernel zeroing of the next chunk on / off, chunk size, CTAs per SM
    base = {"BDE2VID_VOXEL_ZERO_IN_KERNEL": "1", "BDE2VID_VOXEL_CHUNK_MB": "32", "BDE2VID_VOXEL_CTAS_PER_SM": "8", "BDE2VID_PDL": "1"}
    variants = [("memset per chunk, 48 MB (round-2 form)", dict(base, BDE2VID_VOXEL_ZERO_IN_KERNEL="0", BDE2VID_VOXEL_CHUNK_MB="48"))]
    for mb in (16, 24, 32, 48):
        for cps in (4, 8):
            variants.append(("zero in kernel, %d MB, %d CTAs/SM" % (mb, cps), dict(base, BDE2VID_VOXEL_CHUNK_MB=str(mb), BDE2VID_VOXEL_CTAS_PER_SM=str(cps))))
    variants.append(("zero in kernel, 32 MB, 8 CTAs/SM, no PDL", dict(base, BDE2VID_PDL="0")))
    algos = tuple((n, 2, env) for n, env in variants)
    run_shape(260, 346, 31500, 100, 264, 352, 2, 3, algos, streams=("uniform",))
    run_shape(720, 1280, 333333, 64, 720, 1280, 0, 0, algos[:1] + algos[5:7], streams=("uniform",))
    sys.exit(0)

ALGOS = (("2 memset + global RED (default)", 2, {}), ("5 = 2 + warp aggregation", 5, {}), ("1 row-band tiles + warp aggregation", 1, {}),
         ("3 cluster DSMEM reductions", 3, {}), ("4 cluster zero + global RED", 4, {}))
run_shape(260, 346, 31500, 100, 264, 352, 2, 3, ALGOS)
run_shape(720, 1280, 333333, 64, 720, 1280, 0, 0, ALGOS[:2])
