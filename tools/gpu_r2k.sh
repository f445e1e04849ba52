set -x
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2k_bench_cfg4_n1.json 2> gpurun_out/r2k_bench_cfg4_n1.err; tail -3 gpurun_out/r2k_bench_cfg4_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2k_bench_reference.json 2>&1
python bench.py --workload quadpole_sweep --device-only --steps 1 --warmup 1 > gpurun_out/r2k_bench_sweep_1M.json 2>&1
python bench.py --workload quadpole_sweep --sweep-envs-per-gpu 4194304 --device-only --steps 1 --warmup 1 > gpurun_out/r2k_bench_sweep_4M.json 2>&1
python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2k_launches_quadpole.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > gpurun_out/r2k_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rollout_tc256|adv_grpo" -c 2 -o gpurun_out/r2k_prof_k1k2 -f python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > gpurun_out/r2k_ncu_k1k2.log 2>&1
