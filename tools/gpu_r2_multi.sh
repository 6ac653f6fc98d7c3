#!/bin/bash
# multi-GPU checks on N GPUs of one box ($1 = N, $2 = tag): the two-device parity test, the bench line at N ranks, config 5 (seq64) at N ranks
N=${1:-2}; tag=${2:-n2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_${tag}.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or reference_build" 2>&1 | tail -5 | tee gpurun_out/pytest_${tag}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    2> gpurun_out/bench_${tag}.err | tail -1 | tee gpurun_out/bench_${tag}.json | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload seq64 --steps 5 --warmup 3 \
    2> gpurun_out/bench_seq_${tag}.err | tail -1 | tee gpurun_out/bench_seq_${tag}.json | cut -c1-300
tail -n 3 gpurun_out/bench_${tag}.err; tail -n 3 gpurun_out/bench_seq_${tag}.err
