#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): every ncu command is preceded by a plain run of the same command line.
#   1. launch list of the bench (gpu__time_duration.sum per launch; cold-cache, serialised: compare SHARES)
#   2. ncu --set full of the two train kernels of the headline config (-> roofline.traffic via tools/ncu_traffic.py)
#   3. ncu --set full of the ranking kernel (TransH D=100, 16,384 queries)
#   4. ncu --set full of the tcgen05 TransR candidate projection (tensor-pipe utilisation)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --lp-queries 256 --configs 2"
$BENCH > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv $BENCH > gpurun_out/r02_ncu_launch.log 2>&1
$BENCH > gpurun_out/r02_plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"adam_tile_kernel|grad_k1_kernel" -s 16 -c 4 -f -o gpurun_out/r02_train_full $BENCH > gpurun_out/r02_ncu_full_train.log 2>&1
LP="python tools/lp_bench.py TransH 100 8192 1"
$LP > gpurun_out/r02_plain_lp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rank_kernel -c 2 -f -o gpurun_out/r02_rank_full $LP > gpurun_out/r02_ncu_full_rank.log 2>&1
TC="python tools/lp_bench.py TransR 100 4096 1 tc"
$TC > gpurun_out/r02_plain_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transr_project_tc -c 2 -f -o gpurun_out/r02_transr_tc_full $TC > gpurun_out/r02_ncu_full_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
