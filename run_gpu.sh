mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -q --timeout 120 --tb=line 2>&1 | tail -12 | cut -c1-220
timeout 200 python scripts/prof_conv.py --time 2>&1 | tail -30
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --maxfail=5 --tb=line 2>&1 | tail -8 | cut -c1-220
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1g.json 2> gpurun_out/bench_n1g.err; python -c "
import json; r=json.load(open('gpurun_out/bench_n1g.json')); print('N1', r['ms_per_step'], r['value'], r['e2e']['value'], json.dumps(r['roofline']['per_kind']), r['roofline']['frac'])"
