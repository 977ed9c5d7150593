#!/bin/bash
# Round-2 job X (eight B200s): the bench at N = 8 in the round-end state (headline mode only, short).
mkdir -p gpurun_out
O=gpurun_out
T=${1:-x}
PG_BENCH_WATCHDOG=90 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline --no-parity-gate --kernel-steps 4 --gather-batches 2 --modes hbm20 > $O/r2${T}_bench_n8.log 2> $O/r2${T}_bench_n8.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_n8.log | head -c 1600; echo; grep "^\[bench" $O/r2${T}_bench_n8.err | tail -8; tail -3 $O/r2${T}_bench_n8.err
cp $O/bench_detail_n8.json $O/r2${T}_bench_detail_n8.json
date +%s
