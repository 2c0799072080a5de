"""The other BASELINE.json configurations, run as ``python bench.py --config <name>`` (same timing rules and JSON keys as
the headline line; the driver only runs the default config, these lines are kept under profiles/).

  e2vid16  configs[2]: E2VIDRecurrent (model/e2vid, ConvLSTM) random-init, batch 16, 272x352, states carried over 100 steps
  gen4     configs[4]: BDE2VID at Prophesee Gen4 1280x720, 333 333 events per window (10 Mev/s), T = 100, one stream per GPU
  shard64  configs[3]: a FIXED job of 64 independent 346x260 sequences sharded over the ranks (dist.shard_units),
                       device MSE / SSIM per frame and the metric all-reduce INSIDE the timed region (strong scaling)
"""
import os
import time

import numpy as np
import torch

from bde2vid_b200 import synth


def _common(dev, world):
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    return barrier, timed


def _line(metric, value, unit, world, steps, warm, ms, scaling, dtype, config, e2e, launches, clocks, extra):
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    line.update(extra)
    return line


# --------------------------------------------------------------------------------------------------------------------
def run_e2vid16(args, rank, world, dev):
    import bench
    from bde2vid_b200.e2vid import E2VIDRecurrent
    from oracle import oracle_torch as O
    barrier, timed = _common(dev, world)
    B, Hh, Ww, steps_per_seq = 16, 272, 352, 100
    torch.manual_seed(0)
    model = E2VIDRecurrent({"num_bins": 5})
    sd = synth.random_state_dict_like(model.state_dict(), 0)
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(dev)
    model.unetrecurrent.precision = args.precision
    g = torch.Generator().manual_seed(1 + rank)
    n_in = 8
    host = [(torch.randn(B, 5, Hh, Ww, generator=g) * (torch.rand(B, 5, Hh, Ww, generator=g) < 0.3)).pin_memory() for _ in range(n_in)]
    resident = [x.to(dev) for x in host]
    out_host = torch.empty(B, 1, Hh, Ww).pin_memory()

    def seq_resident(i):
        model.reset_states()
        for t in range(steps_per_seq):
            img = model({"events": resident[t % n_in]})["image"]
        return img

    def seq_e2e(i):
        model.reset_states()
        for t in range(steps_per_seq):
            img = model({"events": host[t % n_in].to(dev, non_blocking=True)})["image"]
            out_host.copy_(img, non_blocking=True)

    warm = max(3, args.warmup)
    with torch.no_grad():
        for i in range(min(warm, 3)):
            seq_resident(i)
        sampler = bench.ClockSampler(dev.index or 0)
        sampler.start()
        ms = timed(seq_resident, args.steps)
        clocks = sampler.stop()
        seq_e2e(0)
        ms_e2e = timed(seq_e2e, args.steps)
    frames = world * args.steps * steps_per_seq * B
    fps, fps_e2e = frames / (ms * 1e-3), frames / (ms_e2e * 1e-3)
    # parity + CPU baseline on a bounded sample: 2 recurrent steps at batch 16 through the oracle port
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        st, refs = None, []
        with torch.no_grad():
            for t in range(2):
                r, st = O.e2vid_recurrent_forward(sd, host[t], st)
                refs.append(r)
        dt = time.perf_counter() - t0
        cpu = {"value": 2 * B / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "2 of 100 recurrent steps at batch 16, 272x352, oracle port of model/e2vid (functional torch fp32), %.1f s" % dt}
        model.reset_states()
        with torch.no_grad():
            errs = [float((model({"events": resident[t]})["image"].cpu() - refs[t]).abs().max()) for t in range(2)]
        parity = {"max_abs": max(errs), "steps": 2, "vs": "port", "gate": 2e-3}
    gf = 113.32   # algorithmic GFLOP per frame of E2VIDRecurrent default at 272x352
    return _line("E2VIDRecurrent frames/s (272x352, batch 16, 5-bin)", fps, "frames/s", world, args.steps, warm, ms, "weak",
                 "bf16" if args.precision.startswith("bf16") else "f32",
                 {"workload": "BASELINE.json configs[2]: E2VIDRecurrent({'num_bins': 5}) random-init, batch 16, 272x352, 100 steps per "
                              "sequence with the ConvLSTM states carried, one step = one 100-step sequence of 16 streams",
                  "l2_policy": "no flush: a step streams ~1 GB of activations per recurrent step"},
                 {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": steps_per_seq * B * 5 * Hh * Ww * 4,
                  "d2h_bytes_per_step": steps_per_seq * B * Hh * Ww * 4, "ms_per_step": ms_e2e / args.steps},
                 args.steps * steps_per_seq * 20, clocks,
                 {"cpu_baseline": cpu, "parity": parity, "tflops_algorithmic": fps / world * gf / 1e3, "precision": args.precision,
                  "roofline": {"bound": "tensor", "achieved": fps / world * gf / 1e3, "peak": bench.peaks()["tc_sustained"],
                               "unit": "TFLOP/s", "frac": fps / world * gf / 1e3 / bench.peaks()["tc_sustained"], "traffic": None,
                               "note": "whole-step algorithmic FLOP rate (all kernels of the step), not one kernel"}})


# --------------------------------------------------------------------------------------------------------------------
def run_gen4(args, rank, world, dev):
    import bench
    from bde2vid_b200 import ops
    from bde2vid_b200.model import MODELS
    barrier, timed = _common(dev, world)
    Hh, Ww, N, T = 720, 1280, 333333, args.windows
    model = MODELS.build(bench.cfg_dict())
    model.load_state_dict(synth.init_state_dict(bench.cfg_dict()["generator"], 0), strict=True)
    model = model.eval().to(dev)
    model.generator.precision = args.precision
    n_seq = 2
    host = []
    for i in range(n_seq):
        ev = synth.gen_events(rank + i * world, T, Hh, Ww, N)
        host.append([torch.from_numpy(a).pin_memory() for a in (ev["xs"], ev["ys"], ev["ts"], ev["ps"].view(np.uint8), ev["offsets"])])
    resident = [[a.to(dev) for a in h] for h in host]
    out_host = torch.empty(T, 1, 1, Hh, Ww).pin_memory()

    def step_resident(i):
        return model.reconstruct_events(*resident[i % n_seq], (Hh, Ww))

    def step_e2e(i):
        fr = model.reconstruct_events(*host[i % n_seq], (Hh, Ww))
        out_host.copy_(torch.stack(fr, 0), non_blocking=True)

    warm = max(3, args.warmup)
    with torch.no_grad():
        for i in range(warm):
            step_resident(i)
        sampler = bench.ClockSampler(dev.index or 0)
        sampler.start()
        ms = timed(step_resident, args.steps)
        clocks = sampler.stop()
        step_e2e(0)
        ms_e2e = timed(step_e2e, args.steps)
    fps = world * args.steps * T / (ms * 1e-3)
    fps_e2e = world * args.steps * T / (ms_e2e * 1e-3)
    plan = model.generator.engine().plan(T, 1, Hh, Ww)
    # voxeliser roofline at this shape: raw 13 B / event ingest + one write of the grid
    pk = bench.peaks()
    xs, ys, ts, ps, off = resident[0]
    vout = torch.empty(T, 5, Hh, Ww, device=dev)
    for _ in range(3):
        ops.voxelize_raw(xs, ys, ts, ps, off, 5, Hh, Ww, out=vout)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ops.voxelize_raw(xs, ys, ts, ps, off, 5, Hh, Ww, out=vout)
    e.record()
    torch.cuda.synchronize()
    vms = s.elapsed_time(e) / 5
    vbytes = (13 * N + 4 * 5 * Hh * Ww) * T
    gf = 1600.62
    return _line("reconstructed frames/s (1280x720, 5-bin)", fps, "frames/s", world, args.steps, warm, ms, "weak",
                 "bf16" if args.precision.startswith("bf16") else "f32",
                 {"workload": "BASELINE.json configs[4]: BDE2VID assumed-cfg at Prophesee Gen4 1280x720, T=%d windows x %d events (10 Mev/s at "
                              "30 windows/s), raw int16/float64/bool event ingest, one sequence in flight per GPU" % (T, N),
                  "l2_policy": "no flush: one step streams >20 GB of activations"},
                 {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": 13 * N * T + 8 * (T + 1), "d2h_bytes_per_step": T * Hh * Ww * 4,
                  "ms_per_step": ms_e2e / args.steps},
                 plan.launches * args.steps, clocks,
                 {"cpu_baseline": None, "tflops_algorithmic": fps / world * gf / 1e3, "precision": args.precision,
                  "roofline_voxeliser": {"kernel": "voxel_atomic_kernel<raw> (13 B/event ingest)", "bound": "hbm",
                                         "achieved": vbytes / (vms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                         "frac": vbytes / (vms * 1e-3) / 1e9 / pk["hbm"], "traffic": None, "ms_per_launch": vms,
                                         "bytes_per_launch": vbytes},
                  "roofline": {"bound": "tensor", "achieved": fps / world * gf / 1e3, "peak": pk["tc_sustained"], "unit": "TFLOP/s",
                               "frac": fps / world * gf / 1e3 / pk["tc_sustained"], "traffic": None,
                               "note": "whole-step algorithmic FLOP rate (all kernels), not one kernel"}})


# --------------------------------------------------------------------------------------------------------------------
def run_shard64(args, rank, world, dev):
    import bench
    from bde2vid_b200 import metrics
    from bde2vid_b200.croper import Croper
    from bde2vid_b200.dist import shard_units
    from bde2vid_b200.model import MODELS
    barrier, timed = _common(dev, world)
    Hh, Ww, N, T, n_total, NB = 260, 346, 31500, args.windows, 64, 4
    model = MODELS.build(bench.cfg_dict())
    model.load_state_dict(synth.init_state_dict(bench.cfg_dict()["generator"], 0), strict=True)
    model = model.eval().to(dev)
    model.generator.precision = args.precision
    mine = shard_units([T * 264 * 352] * n_total, world)[rank]           # independent units -> ranks
    # a small pool of distinct event sequences / ground-truth frames cycled over the rank's units (host memory bound)
    pool = 8
    host, gts = [], []
    for i in range(pool):
        ev = synth.gen_events(1000 * rank + i, T, Hh, Ww, N)
        host.append([torch.from_numpy(a).pin_memory() for a in synth.to_loader_format_seq(ev)])
        gts.append(torch.rand(T, Hh, Ww, generator=torch.Generator().manual_seed(i)).to(dev))
    crop = Croper(3)
    crop.update_params(Ww, Hh)
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    result = {}

    def job(i):
        """The whole fixed job of this rank: every unit reconstructed (batches of NB), per-frame MSE / SSIM on the device,
        then ONE all-reduce of the metric sums over the ranks (eval_models_seq.py:278-282)."""
        total = {"mse": 0.0, "ssim": 0.0, "n": 0.0}
        main = torch.cuda.current_stream()
        sums = []
        for k in range(0, len(mine), NB):
            units = mine[k:k + NB]
            st = streams[(k // NB) % 2]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                frames = model.reconstruct_events_batch([host[u % pool] for u in units], (Hh, Ww), slot=(k // NB) % 2)
                for u, fr in zip(units, frames):
                    pred = torch.cat(fr, 0).reshape(T, Hh, Ww)
                    sums.append(metrics.ops.frame_metrics(pred.contiguous(), gts[u % pool], 0, 0, metrics.REFERENCE_DATA_RANGE).sum(0))
        for st in streams:
            main.wait_stream(st)
        tot = torch.stack(sums, 0).sum(0).tolist() if sums else [0.0, 0.0]
        total = {"mse": tot[0], "ssim": tot[1], "n": float(len(mine) * T)}
        result["means"] = metrics.finalize_means(metrics.reduce_metric_sums(total))

    warm = max(1, min(args.warmup, 2))
    with torch.no_grad():
        for i in range(warm):
            job(i)
        sampler = bench.ClockSampler(dev.index or 0)
        sampler.start()
        ms = timed(job, args.steps)
        clocks = sampler.stop()
    fps = args.steps * n_total * T / (ms * 1e-3)
    plan = model.generator.engine().plan(T, NB, 264, 352, 0)
    return _line("reconstructed frames/s (346x260, 5-bin), fixed 64-sequence job", fps, "frames/s", world, args.steps, warm, ms, "strong",
                 "bf16" if args.precision.startswith("bf16") else "f32",
                 {"workload": "BASELINE.json configs[3]: 64 independent 346x260 sequences (T=%d x %d events) sharded over %d rank(s) by "
                              "dist.shard_units, %d per call, device MSE/SSIM per frame + one metric all-reduce inside the timed region; "
                              "events start in pinned HOST memory (H2D inside the timed region)" % (T, N, world, NB),
                  "sequences_per_rank": len(mine), "l2_policy": "no flush: >2 GB of activations per call"},
                 {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": n_total * (16 * N * T + 8 * (T + 1)), "d2h_bytes_per_step": 32,
                  "note": "this config is end to end by construction (host events in, 4 doubles out)"},
                 plan.launches * ((len(mine) + NB - 1) // NB) * args.steps, clocks,
                 {"cpu_baseline": None, "metric_means": result.get("means"), "precision": args.precision,
                  "tflops_algorithmic": fps / world * 163.34 / 1e3})


def run_pair(args, rank, world, dev):
    """ONE 346x260 sequence on ONE GPU vs split over a PAIR of GPUs (engine.PairSplit).  Needs 2 ranks.
    value = frames/s of the pair-split sequence; the single-GPU rate of the same sequence (every rank measures its own) is in
    `single_gpu`.  Strong scaling of one sequence's latency -- the only case the sequence-sharded path cannot speed up."""
    import bench
    from bde2vid_b200.model import MODELS
    assert world == 2, "--config pair needs exactly two ranks (torchrun --nproc-per-node 2)"
    barrier, timed = _common(dev, world)
    Hh, Ww, N, T = 260, 346, 31500, args.windows
    model = MODELS.build(bench.cfg_dict())
    model.load_state_dict(synth.init_state_dict(bench.cfg_dict()["generator"], 0), strict=True)
    model = model.eval().to(dev)
    model.generator.precision = args.precision
    ev = synth.gen_events(0, T, Hh, Ww, N)
    seq = [torch.from_numpy(a).to(dev) for a in synth.to_loader_format_seq(ev)]
    out = {}

    def single(i):
        out["single"] = model.reconstruct_events(*seq, (Hh, Ww))

    def pair_eager(i):
        out["pair"] = model.reconstruct_events_pair(*seq, (Hh, Ww), pair_rank=rank, graph=False)

    def pair_graph(i):
        out["pair_graph"] = model.reconstruct_events_pair(*seq, (Hh, Ww), pair_rank=rank, graph=True)

    res = {}
    warm = max(3, args.warmup)
    with torch.no_grad():
        sampler = bench.ClockSampler(dev.index or 0)
        # the CUDA-graph form (NCCL calls captured) is opt-in: it measured 1736 vs 1706 frames/s eager, but tearing the process
        # group down while the captured graph is alive hung torch.distributed (profiles/r02_bench_pair_2gpu.json)
        variants = [("single", single), ("pair", pair_eager)]
        if os.environ.get("BDE2VID_PAIR_GRAPH", "0") == "1":
            variants.append(("pair_graph", pair_graph))
        res["pair_graph"] = None
        for name, fn in variants:
            try:
                for i in range(warm):
                    fn(i)
                if name == "pair":
                    sampler.start()
                res[name] = timed(fn, args.steps)
                if name == "pair":
                    clocks = sampler.stop()
            except Exception as e:          # graph capture of the NCCL calls is the one thing that may be refused
                if name != "pair_graph":
                    raise
                res[name] = None
                out["pair_graph_error"] = repr(e)[:200]
    same = bool(torch.equal(torch.cat(out["single"], 0), torch.cat(out["pair"], 0)))
    fps = {k: (None if v is None else args.steps * T / (v * 1e-3)) for k, v in res.items()}
    best = max(v for k, v in fps.items() if k != "single" and v is not None)
    eng = model.generator.engine()
    plan = [q for k, q in eng.plans.items() if isinstance(k[4], tuple)][0]
    exch = sum(2 * d["hf"].numel() * 2 for d in plan.lv if "hf" in d) + plan.img.numel() * 4
    return _line("reconstructed frames/s (346x260, 5-bin), ONE sequence split over two GPUs", best, "frames/s", world, args.steps, warm,
                 min(v for k, v in res.items() if k != "single" and v is not None), "strong",
                 "bf16" if args.precision.startswith("bf16") else "f32",
                 {"workload": "SURVEY 8(f4): one 346x260 sequence, T=%d x %d events; rank 0 runs the forward recurrent chains, rank 1 the "
                              "backward ones, hidden states exchanged per level (NCCL broadcast), attention on both, decoder chunks "
                              "alternate, frames by one all-reduce" % (T, N),
                  "l2_policy": "no flush: one sequence streams >2 GB of activations"},
                 {"value": best, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                  "note": "events resident; this config measures the device-side latency of one sequence"},
                 plan.launches * args.steps, clocks,
                 {"cpu_baseline": None, "precision": args.precision, "single_gpu": fps["single"], "pair_eager": fps["pair"],
                  "pair_cuda_graph": fps["pair_graph"], "pair_graph_error": out.get("pair_graph_error"),
                  "bit_identical_to_single_gpu": same, "nccl_bytes_per_sequence": exch})
