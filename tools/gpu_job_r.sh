#!/bin/bash
# Round-2 job R (one B200): where the fixed 17 us of the dW kernel go — epilogue pieces switched off (timing only).
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r}
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
for d in 0 1 2; do
  PG_DW_DEBUG=$d PG_DW_CLUSTER=1 timeout 120 $NCU -k regex:"dw_umma" --log-file $O/r2${T}_dw_d$d.csv python tools/micro_dense.py --iters 6 --only bwd > /dev/null 2>&1
  echo "dw debug=$d rc=$?"; grep dw_umma $O/r2${T}_dw_d$d.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
done
PG_DW_DEBUG=2 PG_DW_CLUSTER=1 timeout 120 $NCU -k regex:"dw_umma" --log-file $O/r2${T}_dw_d2_n.csv python tools/micro_dense.py --iters 6 --only bwd --n 18944 > /dev/null 2>&1
echo "dw debug=2 n=18944 rc=$?"; grep dw_umma $O/r2${T}_dw_d2_n.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -5 | tr '\n' ' '; echo
date +%s
