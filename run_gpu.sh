mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4
timeout 200 python scripts/prof_bn.py > gpurun_out/prof_bn3.txt 2>&1; tail -24 gpurun_out/prof_bn3.txt
timeout 300 python bench_tree.py > gpurun_out/tree_sweep_v4.jsonl 2> gpurun_out/tree_sweep_v4.err; tail -3 gpurun_out/tree_sweep_v4.err; python -c "
import sys,json
for l in open('gpurun_out/tree_sweep_v4.jsonl'):
    r=json.loads(l); print(r['dataset'],r['rows'],r['C'],r['dtype'],r['kind'],'%.3f ms %.0f GB/s %.1f%%'%(r['ms'],r['gbs'],100*r['frac']))" 2>/dev/null | sed -n 1,12p
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err; tail -3 gpurun_out/bench_n1d.err | cut -c1-300; python -c "
import json; r=json.load(open('gpurun_out/bench_n1d.json')); print('N1', r['ms_per_step'], r['value'], r['e2e']['value'], r['gpu_launches'], json.dumps(r['roofline']['per_kind']), r.get('cpu_baseline'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_spec -c 3 -f -o gpurun_out/r01_tree_spec_v2 python scripts/prof_tree.py > gpurun_out/ncu_tree2.log 2>&1; tail -2 gpurun_out/ncu_tree2.log
