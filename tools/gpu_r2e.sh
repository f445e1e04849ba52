set -x
python -m pytest tests/test_gpu_umma.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for np in 4 2; do
TG_TCW_NP=$np python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2e_bench_quadpole_np$np.json 2>&1
TG_TCW_NP=$np ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2e_launches_np$np.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1
done
TG_TCW_NP=2 python -m pytest tests/test_gpu_umma.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
