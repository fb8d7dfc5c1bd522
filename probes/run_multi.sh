#!/bin/bash
# usage: run_multi.sh N  -- config-5 sweep and the default bench on N GPUs of the box (torchrun, one rank per GPU)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-e2e --cpu-seconds 0.2 > gpurun_out/r02_sweep_n$N.json 2> gpurun_out/r02_sweep_n$N.err
tail -2 gpurun_out/r02_sweep_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -2 gpurun_out/r02_bench_n$N.err
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
