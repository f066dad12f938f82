mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_ddp_gpu.py -q --timeout 400 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; tail -3 gpurun_out/bench_n2b.err | cut -c1-300; python -c "
import json; r=json.load(open('gpurun_out/bench_n2b.json')); print('N2', r['ms_per_step'], r['value'], r['e2e']['value'])"
