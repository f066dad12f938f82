set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_backbone_gpu.py -q --timeout 300 2>&1 | tail -15 > gpurun_out/conv_backbone.log; tail -15 gpurun_out/conv_backbone.log
timeout 600 python scripts/diag_network.py > gpurun_out/diag.log 2>&1; tail -40 gpurun_out/diag.log
timeout 300 python -m pytest tests/test_tree_gpu.py tests/test_api_gpu.py -q --timeout 300 2>&1 | tail -5
timeout 300 python bench_tree.py --quick 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['dataset'], r['rows'], r['C'], r['dtype'], r['kind'], '%.3f ms %.0f GB/s frac %.3f' % (r['ms'], r['gbs'], r['frac']))
" | head -8
NBDT_B200_CONV_CLUSTER=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_nocluster.json 2> gpurun_out/bench3.err; python -c "
import json; r=json.load(open('gpurun_out/bench3_nocluster.json')); print('NOCLUSTER', r['ms_per_step'], r['value'], json.dumps(r['roofline']['per_kind']))"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2>> gpurun_out/bench3.err; python -c "
import json; r=json.load(open('gpurun_out/bench3.json')); print('CLUSTER2', r['ms_per_step'], r['value'], json.dumps(r['roofline']['per_kind']), r['roofline']['frac'], r['clocks'])"; tail -3 gpurun_out/bench3.err
