# Round-1 profiling recipe (run under gpurun on one B200; profiles/README.md says what was kept).
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r1z.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_r1z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf32x3_pair -s 3 -c 1 -o gpurun_out/prof_tf32x3_pair_r1 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_tf32pair_r1z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 3 -c 1 -o gpurun_out/prof_dgemm_tma_r1 -f python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_dgemm_r1z.log 2>&1
ls -la gpurun_out/ | tail -8
