mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8
timeout 200 python scripts/prof_bn.py > gpurun_out/prof_bn2.txt 2>&1; tail -24 gpurun_out/prof_bn2.txt
timeout 200 python scripts/prof_conv.py --time > gpurun_out/prof_conv2.txt 2>&1; grep -E "wgrad|totals" gpurun_out/prof_conv2.txt
NBDT_B200_WGRAD_STREAMK=1 timeout 200 python scripts/prof_conv.py --time --main 2>&1 | grep -E "wgrad|totals"
timeout 300 python bench_tree.py > gpurun_out/tree_sweep_v3.jsonl 2> gpurun_out/tree_sweep_v3.err; tail -3 gpurun_out/tree_sweep_v3.err; python -c "
import sys,json
for l in open('gpurun_out/tree_sweep_v3.jsonl'):
    r=json.loads(l); print(r['dataset'],r['rows'],r['C'],r['dtype'],r['kind'],'%.3f ms %.0f GB/s %.1f%%'%(r['ms'],r['gbs'],100*r['frac']))" | head -14
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; tail -3 gpurun_out/bench_n1c.err; python -c "
import json; r=json.load(open('gpurun_out/bench_n1c.json')); print('N1', r['ms_per_step'], r['value'], r['e2e']['value'], r['gpu_launches'], json.dumps(r['roofline']['per_kind']), r['config']['cuda_graph'])"
