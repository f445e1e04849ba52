set -x
python bench.py --workload quadpole2d_cfg3 --device-only --steps 2 --warmup 2 > gpurun_out/r2n_bench_quadpole2d_cfg3.json 2>&1; tail -c 300 gpurun_out/r2n_bench_quadpole2d_cfg3.json
python bench.py --workload quadpole2d_ppo --device-only --steps 2 --warmup 2 > gpurun_out/r2n_bench_quadpole2d_ppo.json 2>&1; tail -c 300 gpurun_out/r2n_bench_quadpole2d_ppo.json
python bench.py --workload cartpole --device-only --steps 3 --warmup 2 > gpurun_out/r2n_bench_cartpole.json 2>&1; tail -c 300 gpurun_out/r2n_bench_cartpole.json
