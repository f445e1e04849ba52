set -x
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2m_bench_quadpole.json 2>&1
python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/r2m_bench_quadpole2d.json 2>&1
