set -x
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2s_bench_cfg4_n1.json 2> gpurun_out/r2s_bench_cfg4_n1.err; tail -3 gpurun_out/r2s_bench_cfg4_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_reference.json 2>&1
