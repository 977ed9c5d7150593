#!/bin/bash
# Round-2 job B (one B200): L2 reuse hint sweep, lagged loss read-back, launch list of the timed region only.
mkdir -p gpurun_out
O=gpurun_out
echo "== fused + engine tests"; date +%s
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_engine.py tests/test_gpu_models.py -x -q > $O/r2b_pytest.log 2>&1
echo "rc=$?"; tail -n 3 $O/r2b_pytest.log
echo "== hot sweep"; date +%s
timeout 400 python tools/micro_fused.py --quick --modes hbm20 --iters 24 --sweep 2:200 --hot-sweep off,0,16,32,48,64,96 > $O/r2b_hot_sweep.json 2> $O/r2b_hot_sweep.err
echo "rc=$?"; cat $O/r2b_hot_sweep.json
echo "== bench n1 (hints on)"; date +%s
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2b_bench_n1.log 2> $O/r2b_bench_n1.err
echo "rc=$?"; tail -n 1 $O/r2b_bench_n1.log | head -c 1600; echo
cp $O/bench_detail_n1.json $O/r2b_bench_detail_n1.json 2>/dev/null
echo "== bench n1 (hints off)"; date +%s
PG_AGG_L2HINT=0 PG_CACHE_HOT_MB=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --modes hbm20 > $O/r2b_bench_n1_nohint.log 2> $O/r2b_bench_n1_nohint.err
echo "rc=$?"; tail -n 1 $O/r2b_bench_n1_nohint.log | head -c 1600; echo
echo "== ncu launch list, timed region only"; date +%s
PG_BENCH_CUDA_PROFILER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
  --log-file $O/r2b_launches_engine.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-parity-gate --kernel-steps 2 \
  --gather-batches 2 --modes hbm20 > $O/r2b_ncu_list.log 2>&1
echo "rc=$?"
echo "== ncu full: fused aggregation with the hint"; date +%s
PG_BENCH_CUDA_PROFILER=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"agg_rows_tma" --launch-skip 3 -c 2 -f -o $O/r2b_full_agg \
  python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-parity-gate --kernel-steps 2 --gather-batches 2 --modes hbm20 > $O/r2b_ncu_agg.log 2>&1
echo "rc=$?"
date +%s
