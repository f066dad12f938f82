mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1; wc -l gpurun_out/launches2.csv
