#!/bin/bash
# two GPUs of one box: the default (weak-scaling) bench under torchrun, the fixed-size sharded job, the reference arm
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2-gpu bench exit $?"; tail -3 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json | tail -1 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --config shard64 > gpurun_out/bench_2gpu_shard64.json 2> gpurun_out/bench_2gpu_shard64.err; echo "2-gpu shard64 exit $?"; tail -3 gpurun_out/bench_2gpu_shard64.err; cat gpurun_out/bench_2gpu_shard64.json | tail -1 | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "2-gpu reference arm exit $?"; tail -1 gpurun_out/bench_2gpu_ref.json | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_driver.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -2
