#!/bin/bash
# Round-2 job U (two B200s): two-rank tests (LL all-reduce + Adam with the counter-advancing launch, peer cache tier) and the
# bench at N = 2 in the round-end state.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-u}
nvidia-smi -L
echo "== two-rank tests"; date +%s
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_gather.py -x -q -k "two_ranks or peer" > $O/r2${T}_pytest_2gpu.log 2>&1
echo "rc=$?"; tail -n 4 $O/r2${T}_pytest_2gpu.log
echo "== bench n2"; date +%s
PG_BENCH_WATCHDOG=200 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --modes hbm20,vtx20,peer20 > $O/r2${T}_bench_n2.log 2> $O/r2${T}_bench_n2.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_n2.log | head -c 1800; echo; grep "timed region done" $O/r2${T}_bench_n2.err
cp $O/bench_detail_n2.json $O/r2${T}_bench_detail_n2.json
date +%s
