for w in 8 12 16; do for i in 1 2; do for b in 200 222; do
PG_AGG_WARPS=$w PG_AGG_ILP=$i timeout 200 python tools/micro_fused.py --iters 12 --sweep 2:$b 2>/dev/null | grep -E "fused_only_(nodrop_)?depth" | head -2 | tr -d '\n' | sed "s/^/w=$w ilp=$i /"; echo; done; done; done
