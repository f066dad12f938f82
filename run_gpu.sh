mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -q --timeout 120 --maxfail=3 --tb=short 2>&1 | tail -25
timeout 200 python scripts/prof_conv.py --time --main 2>&1 | tail -12
NBDT_B200_CONV_PAIR=0 timeout 200 python scripts/prof_conv.py --time --main 2>&1 | tail -12
