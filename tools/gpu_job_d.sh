#!/bin/bash
# Round-2 job D (one B200): dW kernel fixes; SMs left out of the aggregation grid so that the head can run beside it.
mkdir -p gpurun_out
O=gpurun_out
echo "== micro dense"; date +%s
timeout 200 python tools/micro_dense.py --iters 30 --fwd-variants u > $O/r2d_micro_dense.json 2> $O/r2d_micro_dense.err
echo "rc=$?"; cat $O/r2d_micro_dense.json; tail -3 $O/r2d_micro_dense.err
B="python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-parity-gate --kernel-steps 4 --gather-batches 2 --modes hbm20"
for r in 0 8 16 24 32; do
  echo "== reserve $r"; date +%s
  PG_AGG_RESERVE_SMS=$r timeout 300 $B > $O/r2d_bench_res$r.log 2> $O/r2d_bench_res$r.err
  echo "rc=$?"; tail -n 1 $O/r2d_bench_res$r.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_ms'])"
done
echo "== reserve 16 + split"; date +%s
PG_ENGINE_SPLIT=1 PG_AGG_RESERVE_SMS=16 timeout 300 $B > $O/r2d_bench_res16_split.log 2> $O/r2d_bench_res16_split.err
echo "rc=$?"; tail -n 1 $O/r2d_bench_res16_split.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_ms'])"
echo "== breakdown reserve 16"; date +%s
PG_AGG_RESERVE_SMS=16 timeout 300 python tools/engine_breakdown.py 60 hbm20 > $O/r2d_breakdown_res16.json 2> $O/r2d_breakdown_res16.err
echo "rc=$?"
date +%s
