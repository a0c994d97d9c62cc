# Round-2 profile captures (run through gpurun on one B200); every ncu pass only after the same command exited 0 without ncu.
set -x
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-probe --no-sweep --no-selfcheck > gpurun_out/plain_bench_r02.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-probe --no-sweep --no-selfcheck > gpurun_out/ncu_bench_r02.log 2>&1
python scratch/ncu_target.py > gpurun_out/plain_t_r02.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_fe_|k_abs|k_gain|k_gl_iter|k_window" -c 20 -o gpurun_out/prof_r02 -f python scratch/ncu_target.py > gpurun_out/ncu_t_r02.log 2>&1
tail -3 gpurun_out/ncu_t_r02.log
