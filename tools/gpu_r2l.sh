set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --workload pendulum --device-only --steps 20 --warmup 5 > gpurun_out/r2l_bench_pendulum_n8.json 2> gpurun_out/r2l_bench_pendulum_n8.err
tail -3 gpurun_out/r2l_bench_pendulum_n8.err
TG_PEER_ALLREDUCE=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 8 --workload pendulum --device-only --steps 20 --warmup 5 > gpurun_out/r2l_bench_pendulum_n8_nccl.json 2> /dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/r2l_bench_cfg4_n8.json 2> gpurun_out/r2l_bench_cfg4_n8.err
tail -3 gpurun_out/r2l_bench_cfg4_n8.err
