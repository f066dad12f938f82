mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_conv_gpu.py tests/test_efficientnet_gpu.py -q --timeout 300 --tb=short --maxfail=4 2>&1 | tail -8 | cut -c1-220
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1j.json 2> gpurun_out/bench_n1j.err; python -c "
import json; r=json.load(open('gpurun_out/bench_n1j.json')); print('N1', r['ms_per_step'], r['value'], r['e2e']['value'])"
timeout 300 python bench_effnet.py > gpurun_out/bench_effnet.json 2> gpurun_out/bench_effnet.err; tail -2 gpurun_out/bench_effnet.err | cut -c1-200; cat gpurun_out/bench_effnet.json | cut -c1-500
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_effnet.csv python bench_effnet.py --steps 1 --warmup 3 > gpurun_out/ncu_effnet.log 2>&1; tail -1 gpurun_out/ncu_effnet.log | cut -c1-100
