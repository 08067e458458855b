# A/B: k = 1 specialised grad kernel vs the generic one
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "TransH 100 fb15k Adam 1" "TransE 50 fb15k SGD 1" "TransD 100 fb15k Adam 1" "TransE 200 fb15k Adam 1"; do
  echo "== $cfg"; OKB200_GRAD_GENERIC=1 python tools/train_bench.py $cfg 2>&1 | tail -1 | cut -c1-250
  python tools/train_bench.py $cfg 2>&1 | tail -1 | cut -c1-250
done
show='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]
print("us/step %.2f  chunked %.2f | per-launch us:"%(d["ms_per_step"]*1e3, d["training_loop_chunked"]["ms_per_step"]*1e3), {k:round(v*1e3,2) for k,v in r["per_launch_ms"].items()}, "frac %.3f"%r["frac"], r["kernel"])'
B="python bench.py --steps 500 --warmup 10 --no-cpu-baseline --lp-queries 64"
echo "== bench generic"; OKB200_GRAD_GENERIC=1 $B | python -c "$show"
echo "== bench k1"; $B | python -c "$show"
