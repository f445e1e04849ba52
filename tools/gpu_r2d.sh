set -x
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15
python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2d_bench_quadpole_dev.json 2>&1; tail -c 600 gpurun_out/r2d_bench_quadpole_dev.json
python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/r2d_bench_quadpole2d_dev.json 2>&1
python bench.py --workload pendulum --device-only --steps 5 --warmup 3 > gpurun_out/r2d_bench_pendulum_dev.json 2>&1
