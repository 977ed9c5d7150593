#!/bin/bash
# Round-2 job J (one B200): dynamic-schedule tcgen05 dW check, full GPU test suite, the driver's default bench run, reference arm.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-j}
echo "== micro dense"; date +%s
timeout 200 python tools/micro_dense.py --iters 30 --fwd-variants u > $O/r2${T}_micro_dense.json 2> $O/r2${T}_micro_dense.err
echo "rc=$?"; cat $O/r2${T}_micro_dense.json; tail -3 $O/r2${T}_micro_dense.err
echo "== pytest gpu"; date +%s
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2${T}_pytest_gpu.log 2>&1
echo "rc=$?"; tail -n 4 $O/r2${T}_pytest_gpu.log
echo "== bench (driver default)"; date +%s
timeout 900 python bench.py > $O/r2${T}_bench_n1.log 2> $O/r2${T}_bench_n1.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_n1.log | head -c 1500; echo; tail -3 $O/r2${T}_bench_n1.err
cp $O/bench_detail_n1.json $O/r2${T}_bench_detail_n1.json
echo "== reference arm"; date +%s
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2${T}_bench_reference.log 2> $O/r2${T}_bench_reference.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_reference.log | head -c 1200; echo
date +%s
