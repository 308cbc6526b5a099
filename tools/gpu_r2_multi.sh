#!/bin/bash
# multi-GPU lines the driver does not run: strong scaling of config 3 (4096 samples in total) and the config-5 sweep
N=$1
mkdir -p gpurun_out; L=gpurun_out/r2_multi_$N.log; : > $L
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
( time timeout 900 $RUN --steps 1 --warmup 3 --scaling strong ) > gpurun_out/r02_bench_strong_n$N.json 2>> $L
tail -c 2500 gpurun_out/r02_bench_strong_n$N.json >> $L
if [ "$N" = "8" ]; then
( time timeout 600 $RUN --workload sweep ) > gpurun_out/r02_bench_sweep_n$N.json 2>> $L
tail -c 1500 gpurun_out/r02_bench_sweep_n$N.json >> $L
fi
cat $L | tail -c 6000
