#!/bin/bash
# What a round-end validation runs on the GPU box (from the repo root, under gpurun):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh'
set -x
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/validate_bench_n1.json 2> gpurun_out/validate_bench_n1.err; tail -3 gpurun_out/validate_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/validate_bench_reference.json 2>&1
# launch list + full capture of the update / rollout / advantage kernels at one eighth of the default shard
python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/validate_launches_quadpole.csv \
      python bench.py --workload quadpole --device-only --steps 1 --warmup 1 > gpurun_out/validate_ncu_launch.log 2>&1
