#!/bin/bash
# Round-2 job S (one B200): dW epilogue with direct per-lane reductions: correctness, ncu duration, tests, bench.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-s}
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 150 python tools/micro_dense.py --iters 10 --only bwd > $O/r2${T}_micro_dense.json 2> $O/r2${T}_micro_dense.err
echo "rc=$?"; cat $O/r2${T}_micro_dense.json; echo; tail -3 $O/r2${T}_micro_dense.err
for n in 36864 18944; do
timeout 120 $NCU -k regex:"dw_umma" --log-file $O/r2${T}_dw_n$n.csv python tools/micro_dense.py --iters 6 --only bwd --n $n > /dev/null 2>&1
echo "dw n=$n rc=$?"; grep dw_umma $O/r2${T}_dw_n$n.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
done
echo "== dense / engine tests"; date +%s
timeout 900 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_engine.py tests/test_gpu_models.py -x -q > $O/r2${T}_pytest.log 2>&1
echo "rc=$?"; tail -n 3 $O/r2${T}_pytest.log
echo "== bench"; date +%s
PG_BENCH_WATCHDOG=150 timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --kernel-steps 8 --gather-batches 2 > $O/r2${T}_bench.log 2> $O/r2${T}_bench.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench.log | head -c 1500; echo; grep "timed region done" $O/r2${T}_bench.err
cp $O/bench_detail_n1.json $O/r2${T}_bench_detail_n1.json
python - <<P
import json
d=json.load(open('$O/bench_detail_n1.json'))
k=d['kernels']
print({n.split('(')[0]:(round(v['avg_ms'],4), round(v.get('avg_ms_in_pipeline',0),4)) for n,v in k.items()})
P
date +%s
