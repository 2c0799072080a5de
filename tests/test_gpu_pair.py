"""Two GPUs, ONE sequence (engine.PairSplit): the frames of the split run must be bit-identical to
the single-GPU run of the same sequence.  Needs two CUDA devices (skipped on a one-GPU box)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, T, graph, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from bde2vid_b200 import synth
    from bde2vid_b200.model import BDE2VID
    from oracle import oracle_torch as O
    cfg = O.full_cfg({})
    gen_cfg = {k: cfg[k] for k in ("type", "num_bins", "basechannels", "num_encoders", "ks", "num_res_blocks",
                                    "buffer_index", "q_idx", "depths", "num_heads", "losses")}
    model = BDE2VID(generator=gen_cfg)
    model.load_state_dict(synth.init_state_dict(cfg, 0, stress=True), strict=True)
    model = model.eval().to(dev)
    H, W, N = 260, 346, 8000
    ev = synth.gen_events(3, T, H, W, N)
    seq = [torch.from_numpy(a).to(dev) for a in synth.to_loader_format_seq(ev)]
    with torch.no_grad():
        ref = torch.cat(model.reconstruct_events(*seq, (H, W)), 0)
        for _ in range(3):          # the third call replays the captured graph when graph=True
            got = torch.cat(model.reconstruct_events_pair(*seq, (H, W), pair_rank=rank, graph=graph), 0)
    torch.cuda.synchronize()
    ret[rank] = (bool(torch.equal(ref, got)), float((ref - got).abs().max()), float(ref.mean()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("graph", [False, True])
def test_pair_split_bit_identical(graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if graph and os.environ.get("BDE2VID_PAIR_GRAPH", "0") != "1":
        pytest.skip("CUDA-graph capture of the NCCL calls is opt-in (BDE2VID_PAIR_GRAPH=1): process-group teardown hung with it")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29551 + int(graph), 19, graph, ret), nprocs=2, join=True)
    for r in (0, 1):
        same, err, mean = ret[r]
        assert same, "rank %d: pair-split frames differ from the single-GPU frames (max-abs %g)" % (r, err)
        assert 0.0 < mean < 1.0
    assert ret[0][2] == ret[1][2]
