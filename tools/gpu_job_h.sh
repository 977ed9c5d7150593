#!/bin/bash
# Round-2 job H (one B200): sampler kernels fenced off the aggregation's SMs by a shared-memory request.
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-parity-gate --kernel-steps 4 --gather-batches 2 --modes hbm20"
for cfg in "0 24 2" "32768 24 2" "32768 32 2" "32768 16 2" "32768 24 1" "32768 40 2"; do
  set -- $cfg
  echo "== pad $1 reserve $2 samplers $3"; date +%s
  PG_SAMPLE_SMEM_PAD=$1 PG_ENGINE_RESERVE_SMS=$2 PG_ENGINE_SAMPLERS=$3 timeout 300 $B > $O/r2h_bench_p$1_r$2_s$3.log 2> $O/r2h_bench_p$1_r$2_s$3.err
  echo "rc=$?"; tail -n 1 $O/r2h_bench_p$1_r$2_s$3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_ms'])"
  python - <<P
import json
d=json.load(open('$O/bench_detail_n1.json'))
k=d['kernels']
print({n.split('(')[0]:(round(v['avg_ms'],4), round(v.get('avg_ms_in_pipeline',0),4)) for n,v in k.items()})
P
done
date +%s
