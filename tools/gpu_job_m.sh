#!/bin/bash
# Round-2 job M (one B200): forward epilogue on TMA tensor stores + counter-advancing all-reduce/Adam launch.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-m}
echo "== micro dense"; date +%s
timeout 200 python tools/micro_dense.py --iters 30 --fwd-variants u > $O/r2${T}_micro_dense.json 2> $O/r2${T}_micro_dense.err
echo "rc=$?"; cat $O/r2${T}_micro_dense.json; tail -3 $O/r2${T}_micro_dense.err
echo "== dense / engine tests"; date +%s
timeout 900 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_engine.py tests/test_gpu_models.py -x -q > $O/r2${T}_pytest.log 2>&1
echo "rc=$?"; tail -n 3 $O/r2${T}_pytest.log
echo "== bench"; date +%s
PG_BENCH_WATCHDOG=150 timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --kernel-steps 8 --gather-batches 2 > $O/r2${T}_bench.log 2> $O/r2${T}_bench.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench.log | head -c 1500; echo; grep "timed region done" $O/r2${T}_bench.err
cp $O/bench_detail_n1.json $O/r2${T}_bench_detail_n1.json
date +%s
