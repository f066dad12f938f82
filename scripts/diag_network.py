"""Diagnostic (GPU): per-parameter gradient agreement of the B200 networks vs the fp32 eager oracle, next to
the agreement of a bf16-autocast eager run (the noise floor of bf16 activations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200")); sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from nbdt import models
from nbdt.loss import SoftTreeSupLoss
from oracle import torch_port as P

DEV = "cuda:0"
def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))

for arch, B in (("ResNet18", 16), ("ResNet18", 128), ("wrn", 16), ("wrn", 128)):
    torch.manual_seed(0)
    if arch == "ResNet18":
        net, ref, ref16 = models.ResNet18(num_classes=10).to(DEV), P.EagerResNet(num_classes=10).to(DEV), P.EagerResNet(num_classes=10).to(DEV)
        h = "induced-ResNet18"
    else:
        net, ref, ref16 = models.wrn28_10_cifar10().to(DEV), P.EagerWRN(num_classes=10).to(DEV), P.EagerWRN(num_classes=10).to(DEV)
        h = "induced-wrn28_10_cifar10"
    ref.load_state_dict(net.state_dict()); ref16.load_state_dict(net.state_dict())
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn((B, 3, 32, 32), device=DEV, generator=g); y = torch.randint(10, (B,), device=DEV, generator=g)
    crit = SoftTreeSupLoss(dataset="CIFAR10", criterion=nn.CrossEntropyLoss(), hierarchy=h)
    out = net(x); crit(out, y).backward()
    o2 = ref(x); crit(o2, y).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o3 = ref16(x)
    crit(o3.float(), y).backward()
    print(f"== {arch} B={B}: logits err ours {float((out-o2).abs().max()):.4f} autocast {float((o3.float()-o2).abs().max()):.4f} scale {float(o2.abs().max()):.3f}")
    rows = []
    for (n, p), (_, q), (_, r) in zip(net.named_parameters(), ref.named_parameters(), ref16.named_parameters()):
        rows.append((cos(p.grad, q.grad), cos(r.grad, q.grad), float(p.grad.norm() / (q.grad.norm() + 1e-30)), n))
    worst = sorted(rows)[:6]
    for c1, c2, ratio, n in worst:
        print(f"   ours {c1:.4f}  autocast {c2:.4f}  norm ratio {ratio:.3f}  {n}")
    print(f"   min cos ours {min(r[0] for r in rows):.4f}  autocast {min(r[1] for r in rows):.4f}; median ours {sorted(r[0] for r in rows)[len(rows)//2]:.4f}")
