cp tools/ab/libB.so trajopt_grpo_b200/libtrajopt_grpo_b200.so
timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_baseline_shapes.py tests/test_gpu_host_api.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12
for rep in 1 2; do for v in A B; do
cp tools/ab/lib$v.so trajopt_grpo_b200/libtrajopt_grpo_b200.so
timeout 300 python bench.py --workload quadpole --device-only --steps 2 --warmup 2 > gpurun_out/ab_quadpole_${v}_$rep.json 2>&1
timeout 300 python bench.py --workload quadpole2d --device-only --steps 2 --warmup 2 > gpurun_out/ab_quadpole2d_${v}_$rep.json 2>&1
done; done
cp tools/ab/libB.so trajopt_grpo_b200/libtrajopt_grpo_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ab_launches_B.csv python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1
