show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["roofline"]["achieved"],1))'
for k in 512 1000 2000 4000; do
  echo "=== GDL_BN_FUSE_MIN_K=$k"; GDL_BN_FUSE_MIN_K=$k timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --workloads headline 2>/dev/null | python -c "$show"
done
