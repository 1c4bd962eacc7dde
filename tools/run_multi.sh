#!/bin/bash
# Runs on an 8-GPU box (gpurun --gpus 8): the bench at N = 8, the host->device copy matrix at 8 / 4 / 2 concurrent ranks
# (NUMA-bound and not), the reference arm on that box's host cores.  Results land in gpurun_out/m.
O=gpurun_out/m
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" > $O/lscpu.txt 2>&1
$T --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo bench8=$?
for n in 8 4 2; do
  $T --nproc-per-node $n --master-port 2952$n tools/h2d_matrix.py > $O/h2d_matrix_${n}.json 2> $O/h2d_matrix_${n}.err; echo h2d$n=$?
done
CD_NO_NUMA_BIND=1 $T --nproc-per-node 8 --master-port 29531 tools/h2d_matrix.py > $O/h2d_matrix_8_unbound.json 2> $O/h2d_matrix_8_unbound.err; echo h2d8u=$?
$T --nproc-per-node 8 --master-port 29541 bench.py --impl reference --gpus 8 > $O/bench_ref_8.json 2> $O/bench_ref_8.err; echo ref8=$?
tail -c 600 $O/bench_8gpu.json
