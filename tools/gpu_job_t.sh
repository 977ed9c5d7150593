#!/bin/bash
# Round-2 job T (one B200): validation of the round-end state — dW timing, full GPU test suite, smoke(), the driver's bench
# command, reference arm, launch list of the timed region.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-t}
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 120 $NCU -k regex:"umma" --log-file $O/r2${T}_dense.csv python tools/micro_dense.py --iters 6 --fwd-variants u > $O/r2${T}_micro_dense.json 2>/dev/null
echo "rc=$?"; for k in fwd_umma dw_umma; do echo -n "$k: "; grep $k $O/r2${T}_dense.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo; done
python - <<P
import json
d=json.load(open("$O/r2${T}_micro_dense.json"))
print({k:v for k,v in d.items() if "err" in k or "ok" in k})
P
echo "== pytest gpu"; date +%s
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2${T}_pytest_gpu.log 2>&1
echo "rc=$?"; tail -n 4 $O/r2${T}_pytest_gpu.log
echo "== smoke"; date +%s
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2${T}_smoke.log 2>&1
echo "rc=$?"; tail -n 2 $O/r2${T}_smoke.log
echo "== bench (driver command)"; date +%s
PG_BENCH_WATCHDOG=200 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2${T}_bench_n1.log 2> $O/r2${T}_bench_n1.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_n1.log | head -c 1500; echo
cp $O/bench_detail_n1.json $O/r2${T}_bench_detail_n1.json
echo "== bench (default)"; date +%s
PG_BENCH_WATCHDOG=200 timeout 600 python bench.py > $O/r2${T}_bench_default.log 2> $O/r2${T}_bench_default.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_default.log | head -c 1500; echo
cp $O/bench_detail_n1.json $O/r2${T}_bench_default_detail_n1.json
echo "== reference arm"; date +%s
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2${T}_bench_reference.log 2> $O/r2${T}_bench_reference.err
echo "rc=$?"; tail -n 1 $O/r2${T}_bench_reference.log | head -c 1200; echo
echo "== ncu launch list, timed region only"; date +%s
PG_BENCH_CUDA_PROFILER=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
  --log-file $O/r2${T}_launches_engine.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-parity-gate --kernel-steps 2 \
  --gather-batches 2 --modes hbm20 > $O/r2${T}_ncu_list.log 2>&1
echo "rc=$?"
date +%s
