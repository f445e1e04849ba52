set -x
python -m pytest tests/test_gpu_umma.py tests/test_gpu_baseline_shapes.py -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^\[grad" | tail -6
python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -s -p no:cacheprovider -k "gradient and (w128 or w256)" 2>&1 | grep "grad-parity"
for ncv in 8 12; do for tr in 1 0; do
TG_TCW_NCV=$ncv TG_TCW_TRUNC=$tr python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2g_bench_quadpole_ncv${ncv}_tr${tr}.json 2>&1
done; done
TG_TCW_NCV=12 TG_TCW_TRUNC=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2g_launches_ncv12_tr1.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1
TG_TCW_NCV=8 TG_TCW_TRUNC=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2g_launches_ncv8_tr1.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1
