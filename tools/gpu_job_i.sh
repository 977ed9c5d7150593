#!/bin/bash
# Round-2 job I (one B200): engine tests + short bench after an engine / kernel change.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-i}
echo "== engine tests"; date +%s
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_aggregate.py -x -q > $O/r2${T}_pytest.log 2>&1
echo "rc=$?"; tail -n 3 $O/r2${T}_pytest.log
timeout 200 python tools/micro_dense.py --iters 30 --fwd-variants u > $O/r2${T}_micro_dense.json 2> $O/r2${T}_micro_dense.err
echo "rc=$?"; cat $O/r2${T}_micro_dense.json; tail -3 $O/r2${T}_micro_dense.err
B="python bench.py --steps 40 --warmup 6 --no-cpu-baseline --kernel-steps 4 --gather-batches 2"
echo "== bench"; date +%s
timeout 400 $B > $O/r2${T}_bench.log 2> $O/r2${T}_bench.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench.log | head -c 1500; echo; tail -3 $O/r2${T}_bench.err
cp $O/bench_detail_n1.json $O/r2${T}_bench_detail_n1.json
date +%s
