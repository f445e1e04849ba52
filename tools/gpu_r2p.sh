set -x
for ncv in 0 22; do
TG_TCW_NCV=$ncv python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/r2p_bench_quadpole_ncv$ncv.json 2>&1
TG_TCW_NCV=$ncv python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/r2p_bench_quadpole2d_ncv$ncv.json 2>&1
done
