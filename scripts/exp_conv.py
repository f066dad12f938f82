"""Perf experiments on the tap-GEMM kernel (timing only; outputs are garbage in exp modes)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from nbdt._b200 import conv_ops as C
    DEV = "cuda:0"
    for N, H, W, cin, cout, k in [(512, 32, 32, 160, 160, 3), (512, 32, 32, 64, 64, 3), (512, 32, 32, 128, 128, 3),
                                  (512, 32, 32, 256, 256, 3), (512, 16, 16, 320, 320, 3), (512, 8, 8, 640, 640, 3)]:
        x = torch.randn((N, H, W, cin), device=DEV).bfloat16()
        w = torch.randn((cout, cin, k, k), device=DEV) * 0.05
        wp, _ = C.pack_weights(w)
        for _ in range(2): C.conv_fwd(x, wp, k, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): C.conv_fwd(x, wp, k, 1)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flops = 2.0 * N * H * W * cout * cin * k * k
        stages = (N * H * W // 128) * (cout // (160 if cout % 160 == 0 else (128 if cout % 128 == 0 else 64))) * 9 * ((cin + 63) // 64)
        print(f"  {cin:4d}->{cout:4d} {H}x{W}: {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF  {ms * 1e6 / (stages / 148):7.1f} ns/stage/SM")
else:
    for cluster in ("1", "2"):
        for exp in (sys.argv[1:] or ["0", "15"]):
            env = dict(os.environ, NBDT_B200_CONV_EXP=exp, NBDT_B200_CONV_CLUSTER=cluster)
            print(f"== cluster={cluster} exp={exp} (1 skipB 2 skipA 4 skipStore 8 oneMMA)", flush=True)
            subprocess.run([sys.executable, __file__, "child"], env=env)
