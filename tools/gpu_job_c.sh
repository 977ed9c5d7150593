#!/bin/bash
# Round-2 job C (one B200): tcgen05 forward with the bulk-store epilogue, engine timelines, L2 hint variants.
mkdir -p gpurun_out
O=gpurun_out
echo "== micro dense"; date +%s
timeout 200 python tools/micro_dense.py --iters 30 > $O/r2c_micro_dense.json 2> $O/r2c_micro_dense.err
echo "rc=$?"; cat $O/r2c_micro_dense.json; tail -3 $O/r2c_micro_dense.err
echo "== dense + engine tests"; date +%s
timeout 900 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_engine.py -x -q > $O/r2c_pytest.log 2>&1
echo "rc=$?"; tail -n 3 $O/r2c_pytest.log
echo "== engine breakdown"; date +%s
timeout 300 python tools/engine_breakdown.py 60 hbm20 > $O/r2c_breakdown.json 2> $O/r2c_breakdown.err
echo "rc=$?"; head -c 600 $O/r2c_breakdown.json; echo
PG_ENGINE_SPLIT=1 timeout 300 python tools/engine_breakdown.py 60 hbm20 > $O/r2c_breakdown_split.json 2> $O/r2c_breakdown_split.err
echo "rc=$?"; head -c 600 $O/r2c_breakdown_split.json; echo
echo "== hot sweep 2"; date +%s
timeout 400 python tools/micro_fused.py --quick --modes hbm20 --iters 24 --sweep 2:200 --hot-sweep off,8n,24n,48n,96n,24 > $O/r2c_hot_sweep.json 2> $O/r2c_hot_sweep.err
echo "rc=$?"; cat $O/r2c_hot_sweep.json
date +%s
