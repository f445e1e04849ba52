set -x
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
python tools/sanitize_cases.py 2>&1 | tail -8
python bench.py > gpurun_out/r2c_bench_cfg4_n1.json 2> gpurun_out/r2c_bench_cfg4_n1.err; tail -3 gpurun_out/r2c_bench_cfg4_n1.err
python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/r2c_bench_quadpole2d_dev.json 2>&1
python bench.py --workload quadpole_sweep --device-only --steps 1 --warmup 1 > gpurun_out/r2c_bench_sweep_1M.json 2>&1
python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2c_bench_quadpole_dev.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches_quadpole.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > gpurun_out/r2c_ncu_launch.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r2c_sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/r2c_sanitizer_memcheck.log
