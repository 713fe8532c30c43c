#!/bin/bash
# Round-2 (late) evidence, run under gpurun on one B200: every ncu pass only after the same command exited 0 without ncu.
# Files are named r2c_*; profiles/README.md says what each one is.
set -u
O=gpurun_out
mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err || exit 1
./easy-ml_b200/host/nn_bench 8192 300 > $O/r2c_nn_bench.json || exit 1
EML_SIDE=0 ./easy-ml_b200/host/nn_bench 8192 300 > $O/r2c_nn_bench_one_stream.json || exit 1
EML_FUSE=0 ./easy-ml_b200/host/nn_bench 8192 300 > $O/r2c_nn_bench_unfused.json || exit 1
EML_EPILOGUE=1 ./easy-ml_b200/host/nn_bench 8192 300 > $O/r2c_nn_bench_epilogue_chain.json || exit 1
# what each kernel of the recorded step costs on its critical path (eml_graph_profile)
NN_PROFILE=1 ./easy-ml_b200/host/nn_bench 8192 100 2> $O/r2c_nn_critical_path.txt > /dev/null
# kernel completion timeline of eager steps (event after every launch; slower than a normal run)
EML_TRACE=1 ./easy-ml_b200/host/nn_bench 8192 20 2> $O/r2c_nn_timeline.txt > /dev/null
python scripts/e2e_pageable_vs_pinned.py > $O/r2c_e2e_pageable_vs_pinned.txt 2>&1
# launch lists (the recipe's pass; warm caches for the NN step: its kernels run back to back in real life)
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/r2c_launches_nn_step_warm.csv \
    ./easy-ml_b200/host/nn_bench 8192 4 > $O/ncu_nn_warm.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > $O/ncu_bench.log 2>&1
# full captures: the headline kernel again (same code as r2_prof_dgemm_tma), the CTA-pair kernel with the folded repair
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 3 -c 1 -o $O/r2c_prof_dgemm_tma -f \
    python bench.py --steps 2 --warmup 3 --no-extra > $O/ncu_dgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tf32x3_pair|zero_ranges|sgd_batch" -c 4 -o $O/r2c_prof_nn_pair -f \
    ./easy-ml_b200/host/nn_bench 8192 2 > $O/ncu_pair.log 2>&1
for f in r2c_prof_dgemm_tma r2c_prof_nn_pair; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
  ncu -i $O/$f.ncu-rep --page details > $O/${f}_details.txt 2>/dev/null
done
ls -la $O | tail -30
