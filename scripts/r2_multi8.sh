#!/bin/bash
# 8 GPUs: the driver's --gpus N command at N = 8 and 4 (strong headline + weak + gather legs in one line),
# and the multi-GPU tests
mkdir -p gpurun_out/r2m8; O=gpurun_out/r2m8
nvidia-smi -L > $O/gpus.txt
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_gpus$N.json 2> $O/bench_gpus$N.err
  tail -c 400 $O/bench_gpus$N.json; tail -2 $O/bench_gpus$N.err
done
timeout 300 python -m pytest tests/test_gpu_multi.py -q -rs 2>&1 | tail -3 > $O/pytest_multi.log; cat $O/pytest_multi.log
