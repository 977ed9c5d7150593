#!/bin/bash
# Round-2 job E (one B200): split schedule x SMs left out of the aggregation grid.
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-parity-gate --kernel-steps 4 --gather-batches 2 --modes hbm20"
for cfg in "1 8" "1 24" "1 32" "1 40" "1 48" "0 40" "0 48" "0 64"; do
  set -- $cfg
  echo "== split $1 reserve $2"; date +%s
  PG_ENGINE_SPLIT=$1 PG_AGG_RESERVE_SMS=$2 timeout 300 $B > $O/r2e_bench_s$1_r$2.log 2> $O/r2e_bench_s$1_r$2.err
  echo "rc=$?"; tail -n 1 $O/r2e_bench_s$1_r$2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_ms'])"
done
date +%s
