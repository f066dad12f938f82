mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_backbone_gpu.py -q --timeout 200 --tb=line 2>&1 | tail -4 | cut -c1-220
timeout 200 python scripts/prof_conv.py --time 2>&1 | grep -E "wgrad.*k3|totals"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1h.json 2> gpurun_out/bench_n1h.err; python -c "
import json; r=json.load(open('gpurun_out/bench_n1h.json')); print('N1', r['ms_per_step'], r['value'], r['e2e']['value'], json.dumps(r['roofline']['per_kind']), r['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench3.log 2>&1; tail -1 gpurun_out/ncu_bench3.log | cut -c1-100
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:tap_gemm|wgrad_kernel" -c 9 -f -o gpurun_out/r01_conv_v2 python scripts/prof_conv.py --main > gpurun_out/ncu_conv2.log 2>&1; tail -1 gpurun_out/ncu_conv2.log
