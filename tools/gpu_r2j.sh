set -x
python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
for peer in 1 0; do
TG_PEER_ALLREDUCE=$peer python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload pendulum --device-only --steps 20 --warmup 5 > gpurun_out/r2j_bench_pendulum_n2_peer$peer.json 2> gpurun_out/r2j_bench_pendulum_n2_peer$peer.err
tail -c 400 gpurun_out/r2j_bench_pendulum_n2_peer$peer.json | head -c 400; tail -3 gpurun_out/r2j_bench_pendulum_n2_peer$peer.err
done
python bench.py --workload pendulum --device-only --steps 20 --warmup 5 > gpurun_out/r2j_bench_pendulum_n1.json 2>&1
