set -x
python -m pytest tests/test_gpu_umma.py tests/test_gpu_baseline_shapes.py -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^\[grad" | tail -6
python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -s -p no:cacheprovider -k "gradient and (w128 or w256)" 2>&1 | grep "grad-parity"
for ncv in 12 16; do
TG_TCW_NCV=$ncv python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2h_bench_quadpole_ncv${ncv}.json 2>&1
TG_TCW_NCV=$ncv python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/r2h_bench_quadpole2d_ncv${ncv}.json 2>&1
done
TG_TCW_NCV=16 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2h_launches_ncv16.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1
