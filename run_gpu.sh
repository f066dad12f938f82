mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench4.log 2>&1; tail -1 gpurun_out/ncu_bench4.log | cut -c1-100
