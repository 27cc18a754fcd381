#!/bin/bash
# 2 GPUs: the multi-GPU tests (NCCL all-gather and the peer-store gather, rows compared bit for bit) and
# the driver's --gpus 2 command
mkdir -p gpurun_out/r2m2; O=gpurun_out/r2m2
nvidia-smi -L > $O/gpus.txt
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_distributed_gloo.py -q -rs -m "gpu or not gpu" ) > $O/pytest_multi.log 2>&1
tail -8 $O/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_gpus2.json 2> $O/bench_gpus2.err
tail -c 3000 $O/bench_gpus2.json; tail -3 $O/bench_gpus2.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_gpus1.json 2> $O/bench_gpus1.err
tail -c 1200 $O/bench_gpus1.json
