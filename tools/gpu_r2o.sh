set -x
for np in 2 4; do
TG_TCW_NP128=$np python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/r2o_bench_quadpole2d_np$np.json 2>&1
done
TG_TCW_NP128=4 python -m pytest tests/test_gpu_umma.py tests/test_gpu_baseline_shapes.py -m gpu -q -p no:cacheprovider 2>&1 | grep -v "grad-parity" | tail -3
TG_TCW_NP128=4 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2o_launches_np4.csv python bench.py --workload quadpole2d --device-only --steps 1 --warmup 1 > /dev/null 2>&1
TG_TCW_NP128=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2o_launches_np2.csv python bench.py --workload quadpole2d --device-only --steps 1 --warmup 1 > /dev/null 2>&1
