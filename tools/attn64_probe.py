"""Time of one launch of the level-1 (C = 64, head_dim 4) fused window attention at the bench's batch sizes.  Run against
the ablation builds of tools/attn64_probe.sh (BDE2VID_LIB) it tells which unit binds the kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bde2vid_b200 import ops  # noqa: E402
from bde2vid_b200.engine import window_token_map  # noqa: E402

DEV = "cuda"
g = torch.Generator().manual_seed(1)
h, w, Cc, heads, D, q_ind = 132, 176, 64, 16, 3, 1
for B in (1, 4, 8):
    P = B * h * w
    tm, _ = window_token_map(B, h, w, (7, 7), False, DEV)
    nwin = tm.shape[0]
    frames = [(torch.randn(P, Cc, generator=g)).to(DEV) for _ in range(D)]
    wqkv = (torch.randn(3 * Cc, Cc, generator=g) / 8).to(torch.bfloat16).to(DEV)
    bqkv = (torch.randn(3 * Cc, generator=g) * 0.1).to(DEV)
    tbl = (torch.randn(heads, D, 169, generator=g) * 0.5).to(DEV)
    wproj = (torch.randn(Cc, Cc, generator=g) / 8).to(torch.bfloat16).to(DEV)
    bproj = (torch.randn(Cc, generator=g) * 0.1).to(DEV)
    xs = frames[q_ind].clone()
    fr = list(frames)
    fr[q_ind] = xs

    def run():
        ops.window_attention_fused(fr, q_ind, tm.view(-1), nwin, Cc, heads, wqkv, bqkv, tbl, wproj, bproj, xs=xs)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("%-40s B=%d windows=%d  %.1f us per launch" % (os.environ.get("BDE2VID_LIB", "default build").split("/")[-1], B, nwin,
                                                       e0.elapsed_time(e1) * 1e3 / 20))
