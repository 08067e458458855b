# A/B of build variants (tools/build_variant.sh) on the chunked train loop, warm L2
for v in main nograd b888 b444 nogradb888; do
  if [ $v = main ]; then unset OKB200_LIB; else export OKB200_LIB=$PWD/openkeonspark_b200/variants/libokb200_$v.so; fi
  echo "== $v"; python tools/train_bench.py TransH 100 fb15k Adam 1 2>&1 | tail -1
done
