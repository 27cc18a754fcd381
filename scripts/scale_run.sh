N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/s2_scale_weak_$N.json 2> gpurun_out/s2_scale_weak_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 --scaling strong > gpurun_out/s2_scale_strong_$N.json 2> gpurun_out/s2_scale_strong_$N.err
tail -c 600 gpurun_out/s2_scale_weak_$N.json; tail -3 gpurun_out/s2_scale_weak_$N.err
