# A/B: several warps per positive in the generic grad kernel
python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -1
for cfg in "TransD 100 wn18 Adam 10" "TransE 100 wn18 SGD 25" "TransH 100 fb15k Adam 3 400"; do
  echo "== $cfg"; OKB200_GRAD_SINGLE_WARP=1 python tools/train_bench.py $cfg 2>&1 | tail -1 | cut -c1-240; python tools/train_bench.py $cfg 2>&1 | tail -1 | cut -c1-240
done
