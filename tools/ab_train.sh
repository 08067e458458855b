# A/B of the Adam pass variants
python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -2
for cfg in "TransH 100 fb15k Adam 1" "TransD 100 wn18 Adam 10" "TransE 200 fb15k Adam 1" "TransE 50 fb15k Adam 1"; do
  echo "== $cfg"; OKB200_ADAM_LEGACY=1 python tools/train_bench.py $cfg 2>&1 | tail -1 | cut -c1-260
  for v in main t5 t4; do
    if [ $v = main ]; then unset OKB200_LIB; else export OKB200_LIB=$PWD/openkeonspark_b200/variants/libokb200_$v.so; fi
    python tools/train_bench.py $cfg 2>&1 | tail -1 | cut -c1-260
  done; unset OKB200_LIB
done
