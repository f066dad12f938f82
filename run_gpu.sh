mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --maxfail=12 --tb=short 2>&1 | tail -40
timeout 300 python bench_tree.py --quick > gpurun_out/tree_sweep_v6.jsonl 2> gpurun_out/tree_sweep_v6.err; tail -3 gpurun_out/tree_sweep_v6.err; python -c "
import sys,json
for l in open('gpurun_out/tree_sweep_v6.jsonl'):
    r=json.loads(l); print(r['dataset'],r['rows'],r['C'],r['dtype'],r['kind'],'%.3f ms %.0f GB/s %.1f%%'%(r['ms'],r['gbs'],100*r['frac']))" 2>/dev/null
