timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/r2_bench_n4.err | grep '^{' > gpurun_out/r2_bench_n4.json
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n4.json'));print(d['value'],d['ms_per_step'],d['parity']['ok'],d['e2e']['value'])"
