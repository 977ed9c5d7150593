for cfg in "200000 16 16" "100000 16 16" "100000 8 16" "120000 12 16" "150000 16 16" "200000 16 32" "100000 8 32"; do set -- $cfg
echo "agg_smem=$1 agg_warps=$2 dense_tile=$3"; PG_AGG_SMEM=$1 PG_AGG_WARPS=$2 PG_DENSE_TILE=$3 timeout 250 python tools/engine_breakdown.py 150 hbm20 2>/dev/null | tail -1; done
