mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backbone_gpu.py -q --timeout 300 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; python -c "
import json; r=json.load(open('gpurun_out/bench6.json')); print('BENCH', r['ms_per_step'], r['value'], r['roofline']['frac'], r['roofline']['conv_share_of_step'], r['clocks'])"; tail -3 gpurun_out/bench6.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1; wc -l gpurun_out/launches3.csv
