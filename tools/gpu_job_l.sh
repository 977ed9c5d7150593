#!/bin/bash
# Round-2 job L (one B200): SM-reserve sweep with the dynamic dW, launch list + ncu full captures of the final kernels, stage timeline.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-l}
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-parity-gate --kernel-steps 4 --gather-batches 2 --modes hbm20"
for r in 16 24 32 40; do
  echo "== reserve $r"; date +%s
  PG_ENGINE_RESERVE_SMS=$r timeout 200 $B > $O/r2${T}_bench_res$r.log 2> $O/r2${T}_bench_res$r.err
  echo "rc=$?"; tail -n 1 $O/r2${T}_bench_res$r.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_ms'])"
done
echo "== ncu launch list, timed region only"; date +%s
PG_BENCH_CUDA_PROFILER=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
  --log-file $O/r2${T}_launches_engine.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-parity-gate --kernel-steps 2 \
  --gather-batches 2 --modes hbm20 > $O/r2${T}_ncu_list.log 2>&1
echo "rc=$?"
echo "== ncu full: fused aggregation (in-step grid)"; date +%s
PG_BENCH_CUDA_PROFILER=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"agg_rows_tma" --launch-skip 3 -c 2 -f -o $O/r2${T}_full_agg \
  python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-parity-gate --kernel-steps 2 --gather-batches 2 --modes hbm20 > $O/r2${T}_ncu_agg.log 2>&1
echo "rc=$?"
echo "== ncu full: dense kernels"; date +%s
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dw_umma" --launch-skip 4 -c 2 -f -o $O/r2${T}_full_dw_umma \
  python tools/micro_dense.py --iters 4 --only bwd > $O/r2${T}_ncu_dw.log 2>&1
echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fwd_umma" --launch-skip 4 -c 2 -f -o $O/r2${T}_full_fwd_umma \
  python tools/micro_dense.py --iters 4 --only fwd --fwd-variants u > $O/r2${T}_ncu_fwd.log 2>&1
echo "rc=$?"
echo "== breakdown"; date +%s
timeout 300 python tools/engine_breakdown.py 60 hbm20 > $O/r2${T}_breakdown.json 2> $O/r2${T}_breakdown.err
echo "rc=$?"
date +%s
