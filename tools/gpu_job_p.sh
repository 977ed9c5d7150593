#!/bin/bash
# Round-2 job P (one B200): dW with the merged N = 64 product + cluster reduction: correctness, ncu durations, tests, bench A/B.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-p}
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
echo "== micro dense (correctness)"; date +%s
for c in 1 4; do
PG_DW_CLUSTER=$c timeout 150 python tools/micro_dense.py --iters 10 --fwd-variants u > $O/r2${T}_micro_dense_c$c.json 2> $O/r2${T}_micro_dense_c$c.err
echo "rc=$?"; python - <<P
import json
d=json.load(open("$O/r2${T}_micro_dense_c$c.json"))
print({k:v for k,v in d.items() if "err" in k or "ok" in k})
P
tail -3 $O/r2${T}_micro_dense_c$c.err
done
timeout 120 $NCU -k regex:"fwd_umma" --log-file $O/r2${T}_fwd.csv python tools/micro_dense.py --iters 6 --only fwd --fwd-variants u > /dev/null 2>&1
echo "fwd rc=$?"; grep fwd_umma $O/r2${T}_fwd.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
for c in 1 2 4; do
  PG_DW_CLUSTER=$c timeout 120 $NCU -k regex:"dw_umma" --log-file $O/r2${T}_dw_c$c.csv python tools/micro_dense.py --iters 6 --only bwd > /dev/null 2>&1
  echo "dw cluster=$c rc=$?"; grep dw_umma $O/r2${T}_dw_c$c.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
done
echo "== dense / engine tests"; date +%s
timeout 900 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_engine.py tests/test_gpu_models.py -x -q > $O/r2${T}_pytest.log 2>&1
echo "rc=$?"; tail -n 3 $O/r2${T}_pytest.log
B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-parity-gate --kernel-steps 4 --gather-batches 2 --modes hbm20"
for c in 1 2 4; do
  echo "== bench cluster $c"; date +%s
  PG_DW_CLUSTER=$c PG_BENCH_WATCHDOG=100 timeout 200 $B > $O/r2${T}_bench_c$c.log 2> $O/r2${T}_bench_c$c.err
  echo "rc=$?"; tail -n 1 $O/r2${T}_bench_c$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_ms'])"
done
date +%s
