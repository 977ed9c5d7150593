#!/bin/bash
# Round-2 job N (one B200): forward with the N = 64 merged product, dW cluster reduction (cluster 1 / 2 / 4).
mkdir -p gpurun_out
O=gpurun_out
T=${1:-n}
for c in 1 2 4; do
  echo "== micro dense, dW cluster $c"; date +%s
  PG_DW_CLUSTER=$c timeout 150 python tools/micro_dense.py --iters 30 --fwd-variants u > $O/r2${T}_micro_dense_c$c.json 2> $O/r2${T}_micro_dense_c$c.err
  echo "rc=$?"; python - <<P
import json
d=json.load(open("$O/r2${T}_micro_dense_c$c.json"))
print({k:v for k,v in d.items() if k.startswith("fwd_variant_u") or k.startswith("bwd_umma") or k=="fwd_max_err_vs_fp64"})
P
  tail -3 $O/r2${T}_micro_dense_c$c.err
done
date +%s
