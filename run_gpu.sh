mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_efficientnet_gpu.py tests/test_backbone_gpu.py -q --timeout 300 --tb=short --maxfail=8 2>&1 | tail -40 | cut -c1-230
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1m.json 2> gpurun_out/bench_n1m.err; python -c "
import json; r=json.load(open('gpurun_out/bench_n1m.json')); print('N1', r['ms_per_step'], r['value'], r['e2e']['value'], r['config']['cuda_graph'], r['roofline']['frac'], r['tree_roofline']['launch_us'])"
