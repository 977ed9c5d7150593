#!/bin/bash
# Round-2 measurement job (one B200): headline bench, GPU tests, reference arm, ncu launch list, ncu full captures.
# Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand afterwards.
mkdir -p gpurun_out
O=gpurun_out
(nvidia-smi; lscpu | head -20) > $O/r2_machine.txt 2>&1

echo "== bench n1"; date +%s
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.log 2> $O/r2_bench_n1.err
echo "rc=$?"; tail -n 1 $O/r2_bench_n1.log | head -c 1600; echo
cp $O/bench_detail_n1.json $O/r2_bench_detail_n1.json 2>/dev/null

echo "== pytest gpu"; date +%s
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1
echo "rc=$?"; tail -n 4 $O/r2_pytest_gpu.log

echo "== reference arm"; date +%s
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.log 2> $O/r2_bench_reference.err
echo "rc=$?"; tail -n 1 $O/r2_bench_reference.log | head -c 1200; echo

echo "== ncu launch list"; date +%s
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 400 --csv --log-file $O/r2_launches_engine.csv \
  python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-parity-gate --kernel-steps 2 --gather-batches 2 --modes hbm20 > $O/r2_ncu_list.log 2>&1
echo "rc=$?"

echo "== ncu full: dense kernels"; date +%s
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fwd_umma" --launch-skip 4 -c 2 -f -o $O/r2_full_fwd_umma \
  python tools/micro_dense.py --iters 4 --only fwd --fwd-variants u > $O/r2_ncu_fwd.log 2>&1
echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dw_umma" --launch-skip 4 -c 2 -f -o $O/r2_full_dw_umma \
  python tools/micro_dense.py --iters 4 --only bwd > $O/r2_ncu_dw.log 2>&1
echo "rc=$?"
timeout 120 python tools/micro_dense.py --iters 30 > $O/r2_micro_dense.json 2> $O/r2_micro_dense.err
echo "rc=$?"; cat $O/r2_micro_dense.json

echo "== ncu full: fused aggregation in the engine"; date +%s
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"agg_rows_tma" --launch-skip 8 -c 2 -f -o $O/r2_full_agg \
  python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-parity-gate --kernel-steps 2 --gather-batches 2 --modes hbm20 > $O/r2_ncu_agg.log 2>&1
echo "rc=$?"
date +%s
ls -la $O
