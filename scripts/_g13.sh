timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/r2_bench_n8.err | grep '^{' > gpurun_out/r2_bench_n8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tests/dist_check.py 2>&1 | tail -2 > gpurun_out/r2_dist8.log
cat gpurun_out/r2_dist8.log
