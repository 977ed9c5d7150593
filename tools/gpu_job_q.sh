#!/bin/bash
# Round-2 job Q (one B200): dW templated on the cluster size (unrolled reduction loop): durations by ncu over n, bench A/B.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-q}
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
PG_DW_CLUSTER=1 timeout 150 python tools/micro_dense.py --iters 10 --only bwd > $O/r2${T}_micro_dense.json 2> $O/r2${T}_micro_dense.err
echo "rc=$?"; cat $O/r2${T}_micro_dense.json; echo
for c in 1 2 4; do
  PG_DW_CLUSTER=$c timeout 120 $NCU -k regex:"dw_umma" --log-file $O/r2${T}_dw_c$c.csv python tools/micro_dense.py --iters 6 --only bwd > /dev/null 2>&1
  echo "dw cluster=$c rc=$?"; grep dw_umma $O/r2${T}_dw_c$c.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
done
for n in 18944 75776; do
  PG_DW_CLUSTER=4 timeout 120 $NCU -k regex:"dw_umma" --log-file $O/r2${T}_dw_n$n.csv python tools/micro_dense.py --iters 6 --only bwd --n $n > /dev/null 2>&1
  echo "dw c=4 n=$n rc=$?"; grep dw_umma $O/r2${T}_dw_n$n.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
done
B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-parity-gate --kernel-steps 4 --gather-batches 2 --modes hbm20"
for c in 1 4 2; do
  echo "== bench cluster $c"; date +%s
  PG_DW_CLUSTER=$c PG_BENCH_WATCHDOG=100 timeout 200 $B > $O/r2${T}_bench_c$c.log 2> $O/r2${T}_bench_c$c.err
  echo "rc=$?"; tail -n 1 $O/r2${T}_bench_c$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_ms'])"
  python - <<P
import json
d=json.load(open('$O/bench_detail_n1.json'))
k=d['kernels']
print({n.split('(')[0]:(round(v['avg_ms'],4), round(v.get('avg_ms_in_pipeline',0),4)) for n,v in k.items()})
P
done
date +%s
