set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r1d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_r1d.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/plain_r1d2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 3 -c 1 -o gpurun_out/prof_dgemm_tma_r1 -f python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_dgemm_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf32x3_kernel -s 3 -c 1 -o gpurun_out/prof_tf32x3_r1 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_tf32_r1d.log 2>&1
ls -la gpurun_out/
