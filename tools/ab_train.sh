# A/B of runtime flags on the bench workload (not the driver's bench)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
show='
import json,sys
d=json.loads(sys.stdin.read()); r=d["roofline"]
print("us/step %.2f  chunked %.2f  e2e %.3g | per-launch us:"%(d["ms_per_step"]*1e3, d["training_loop_chunked"]["ms_per_step"]*1e3, d["e2e"]["value"]), {k:round(v*1e3,2) for k,v in r["per_launch_ms"].items()})'
B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --lp-queries 64"
echo "== main"; $B | python -c "$show"
echo "== no L2 prefetch"; OKB200_L2_PREFETCH=0 $B | python -c "$show"
echo "== simple adam"; OKB200_ADAM_SIMPLE=1 $B | python -c "$show"
echo "== simple adam, no prefetch"; OKB200_L2_PREFETCH=0 OKB200_ADAM_SIMPLE=1 $B | python -c "$show"
echo "== no PDL"; OKB200_PDL=0 $B | python -c "$show"
for a in "TransD 100 wn18 Adam 10" "TransE 200 fb15k Adam 1"; do python tools/train_bench.py $a 2>&1 | tail -1; OKB200_ADAM_SIMPLE=1 python tools/train_bench.py $a 2>&1 | tail -1;  OKB200_L2_PREFETCH=0 python tools/train_bench.py $a 2>&1 | tail -1; done
