#!/bin/bash
# Round-2 job G (two B200s): two-rank tests (LL all-reduce + Adam, peer cache tier), bench at N = 1 and N = 2, engine timeline.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
echo "== two-rank tests"; date +%s
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_gather.py -x -q -k "two_ranks or peer or engine" > $O/r2g_pytest_2gpu.log 2>&1
echo "rc=$?"; tail -n 4 $O/r2g_pytest_2gpu.log
echo "== bench n1"; date +%s
timeout 600 python bench.py --steps 40 --warmup 6 --no-cpu-baseline > $O/r2g_bench_n1.log 2> $O/r2g_bench_n1.err
echo "rc=$?"; tail -n 1 $O/r2g_bench_n1.log | head -c 1500; echo
cp $O/bench_detail_n1.json $O/r2g_bench_detail_n1.json
echo "== bench n2"; date +%s
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 6 --no-cpu-baseline > $O/r2g_bench_n2.log 2> $O/r2g_bench_n2.err
echo "rc=$?"; tail -n 1 $O/r2g_bench_n2.log | head -c 1500; echo
cp $O/bench_detail_n2.json $O/r2g_bench_detail_n2.json
echo "== breakdown (defaults: split + 24 SMs)"; date +%s
timeout 300 python tools/engine_breakdown.py 60 hbm20 > $O/r2g_breakdown.json 2> $O/r2g_breakdown.err
echo "rc=$?"
date +%s
