# Round-1 profiling recipe (run under gpurun on one B200; profiles/README.md says what was kept).
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r1z.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_r1z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf32x3_pair -s 3 -c 1 -o gpurun_out/prof_tf32x3_pair_r1 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_tf32pair_r1z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 3 -c 1 -o gpurun_out/prof_dgemm_tma_r1 -f python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_dgemm_r1z.log 2>&1
# in-kernel split variant of the f32 GEMM on config 3's shape (after the same command ran clean without ncu)
EML_SGEMM_FUSED=1 python scripts/fused_one.py 1024 1024 1024 64 &&
EML_SGEMM_FUSED=1 ncu --set full --clock-control none --import-source on -k regex:tf32x3_fused -c 1 -o gpurun_out/fused_r2 -f python scripts/fused_one.py 1024 1024 1024 64 > gpurun_out/ncu_fused.log 2>&1
# f64 through int8 slices: launch list of one call, full capture of the main kernel, the headline line through it
python scripts/sliced_one.py 8192 8 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sliced_launches.csv python scripts/sliced_one.py 8192 8 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:sliced_kernel -s 1 -c 1 -o gpurun_out/sliced_r3 -f python scripts/sliced_one.py 8192 0 > gpurun_out/ncu_sliced.log 2>&1
python bench.py --steps 5 --warmup 3 --no-extra --f64-path sliced > gpurun_out/bench_sliced.json 2>/dev/null
# NN step launch list (C++ host mirror)
(cd easy-ml_b200/host && LD_LIBRARY_PATH=../lib ./nn_bench 8192 6 > ../../gpurun_out/nn_plain.log 2>&1 &&
 LD_LIBRARY_PATH=../lib ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ../../gpurun_out/nn_launches.csv ./nn_bench 8192 3 > ../../gpurun_out/nn_ncu.log 2>&1)
ls -la gpurun_out/ | tail -8
