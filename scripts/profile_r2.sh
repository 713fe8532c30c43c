#!/bin/bash
# Round-2 evidence, run under gpurun on one B200: every ncu pass only after the same command exited 0 without ncu.
set -u
O=gpurun_out
mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err || exit 1
./easy-ml_b200/host/nn_bench 8192 300 > $O/r2_nn_bench.json || exit 1
EML_FUSE=0 ./easy-ml_b200/host/nn_bench 8192 300 > $O/r2_nn_bench_unfused.json || exit 1
python scripts/ew_bench.py 8192 512 300 > $O/r2_ew_bench.txt 2>&1
# launch lists (the recipe's pass: cold caches, serialised launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2_launches_nn_step.csv \
    ./easy-ml_b200/host/nn_bench 8192 4 > $O/ncu_nn.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 500 --csv --log-file $O/r2_launches_nn_step_warm.csv \
    ./easy-ml_b200/host/nn_bench 8192 4 > $O/ncu_nn_warm.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > $O/ncu_bench.log 2>&1
# full captures: the headline kernel, the fused elementwise kernels of the NN step, the vectorised skinny product
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 3 -c 1 -o $O/r2_prof_dgemm_tma \
    python bench.py --steps 2 --warmup 3 --no-extra > $O/ncu_dgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"chain_(forward|backward)_kernel<float, 4" -c 4 -o $O/r2_prof_chain \
    ./easy-ml_b200/host/nn_bench 8192 2 > $O/ncu_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"skinny_rows|smallk_reg|tn_wide|quirk_cols|pair" -c 6 -o $O/r2_prof_nn_misc \
    ./easy-ml_b200/host/nn_bench 8192 2 > $O/ncu_misc.log 2>&1
for f in r2_prof_dgemm_tma r2_prof_chain r2_prof_nn_misc; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
  ncu -i $O/$f.ncu-rep --page details > $O/${f}_details.txt 2>/dev/null
done
ls -la $O | tail -30
