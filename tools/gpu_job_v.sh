#!/bin/bash
# Round-2 job V (four B200s): BASELINE config 3 (GraphSAGE-mean, dg partition = 4, one partition per GPU) at N = 4.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-v}
nvidia-smi -L
echo "== bench config 3, n4"; date +%s
PG_BENCH_WATCHDOG=120 timeout 230 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --config 3 --steps 40 --warmup 6 --no-cpu-baseline --kernel-steps 4 --gather-batches 4 > $O/r2${T}_bench_cfg3_n4.log 2> $O/r2${T}_bench_cfg3_n4.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_cfg3_n4.log | head -c 1800; echo; grep "^\[bench" $O/r2${T}_bench_cfg3_n4.err | tail -16; tail -5 $O/r2${T}_bench_cfg3_n4.err
cp $O/bench_detail_cfg3_n4.json $O/r2${T}_bench_detail_cfg3_n4.json
date +%s
