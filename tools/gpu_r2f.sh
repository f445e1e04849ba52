set -x
python -m pytest tests/test_gpu_umma.py tests/test_gpu_baseline_shapes.py tests/test_gpu_host_api.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2f_bench_quadpole.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2f_launches_quadpole.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1
python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/r2f_bench_quadpole2d.json 2>&1
