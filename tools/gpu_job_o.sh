#!/bin/bash
# Round-2 job O (one B200): kernel durations by ncu (gpu__time_duration only) — fwd at 1 / 2 / 4 tile waves, dW at cluster 1 / 2 / 4.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-o}
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
for n in 18944 37888 75776; do
  timeout 120 $NCU -k regex:"fwd_umma" --log-file $O/r2${T}_fwd_n$n.csv python tools/micro_dense.py --iters 6 --only fwd --fwd-variants u --n $n > /dev/null 2>&1
  echo "fwd n=$n rc=$?"; grep fwd_umma $O/r2${T}_fwd_n$n.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
done
for c in 1 2 4; do
  PG_DW_CLUSTER=$c timeout 120 $NCU -k regex:"dw_umma" --log-file $O/r2${T}_dw_c$c.csv python tools/micro_dense.py --iters 6 --only bwd > /dev/null 2>&1
  echo "dw cluster=$c rc=$?"; grep dw_umma $O/r2${T}_dw_c$c.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
done
date +%s
