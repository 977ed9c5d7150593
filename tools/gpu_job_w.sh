#!/bin/bash
# Round-2 job W (one B200): ncu full captures of the round-end tcgen05 kernels.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-w}
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"dw_umma" --launch-skip 10 -c 2 -f -o $O/r2${T}_full_dw_umma \
  python tools/micro_dense.py --iters 2 --only bwd > $O/r2${T}_ncu_dw.log 2>&1
echo "rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"fwd_umma" --launch-skip 10 -c 2 -f -o $O/r2${T}_full_fwd_umma \
  python tools/micro_dense.py --iters 2 --only fwd --fwd-variants u > $O/r2${T}_ncu_fwd.log 2>&1
echo "rc=$?"
timeout 100 python tools/micro_dense.py --iters 20 --fwd-variants u > $O/r2${T}_micro_dense.json 2> $O/r2${T}_micro_dense.err
echo "rc=$?"; cat $O/r2${T}_micro_dense.json; echo
date +%s
