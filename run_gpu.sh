mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -q --timeout 400 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err; wc -l gpurun_out/bench_n2c.json; python -c "
import json; r=json.loads(open('gpurun_out/bench_n2c.json').read().strip().splitlines()[-1]); print('N2', r['ms_per_step'], r['value'], r['e2e']['value'], r['config']['cuda_graph'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
