mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_ddp_gpu.py tests/test_conv_gpu.py -q --timeout 400 -k "ddp or two_rank or wgrad" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-400; tail -4 gpurun_out/bench_n2.err | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1b.json 2>/dev/null; python -c "
import json; r=json.load(open('gpurun_out/bench_n1b.json')); print('N1', r['ms_per_step'], r['value'], json.dumps(r['roofline']['per_kind']))"
