#!/bin/bash
# Round-2 job F (one B200): where do the tcgen05 NodeUpdate kernels spend their time? rows = 1, 2, 4 tiles per CTA, with / without dropout output.
mkdir -p gpurun_out
O=gpurun_out
for n in 18944 37888 75776; do for p in 0.2 0.0; do
  echo "== n $n p $p"
  timeout 120 python tools/micro_dense.py --iters 30 --fwd-variants u --only fwd --n $n --p $p 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['fwd_variant_u_us'], d['cublas_fp32_linear_us'])"
done; done
for n in 18944 37888 75776; do
  echo "== bwd n $n"
  timeout 120 python tools/micro_dense.py --iters 30 --only bwd --n $n 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['bwd_umma_3xtf32_us'], d['bwd_mma_3xtf32_us'])"
done
