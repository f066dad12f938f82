set -x
mkdir -p gpurun_out
for t in test_pack_weights_padding test_conv_fwd test_conv_dgrad test_conv_wgrad; do
  timeout 300 python -m pytest tests/test_conv_gpu.py -q --timeout 120 -k $t 2>&1 | tail -25 > gpurun_out/conv_$t.log
  tail -12 gpurun_out/conv_$t.log
done
