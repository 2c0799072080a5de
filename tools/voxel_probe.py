"""Timing of the voxeliser algorithms on the bench workload (346x260, 100 windows x 31,500 events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bde2vid_b200 import ops, synth  # noqa: E402

H, W, NEV, T = 260, 346, 31500, 100
ev = synth.gen_events(0, T, H, W, NEV)
xs, ys, ts, ps, off = [torch.from_numpy(a).cuda() for a in synth.to_loader_format_seq(ev)]
out = torch.empty(T, 5, 264, 352, device="cuda")
ref = None
for name, algo, env in (("row-band + warp aggregation", 1, {}), ("global atomics", 2, {}), ("cluster, remote reductions", 3, {}),
                        ("cluster, scan-all + local atomics", 3, {"BDE2VID_VOXEL_SCANALL": "1"}),
                        ("cluster, zero + global reductions", 4, {})):
    os.environ.pop("BDE2VID_VOXEL_SCANALL", None)
    os.environ.update(env)
    for _ in range(3):
        ops.voxelize_seq(xs, ys, ts, ps, off, 5, H, W, 2, 3, 264, 352, out=out, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.voxelize_seq(xs, ys, ts, ps, off, 5, H, W, 2, 3, 264, 352, out=out, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if ref is None:
        ref = out.clone()
    err = float((out - ref).abs().max())
    print("%-36s %.3f ms  %.0f GB/s  max |diff| vs algo 1 %.2e" % (name, ms, (16 * NEV + 4 * 5 * H * W) * T / ms / 1e6, err))
