"""Measured end-to-end error of the B200 backbones against the reference-generated goldens (what DESIGN.md §2 quotes).
usage: python scripts/measure_parity.py > profiles/r02_parity_measured.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-backed-decision-trees_b200"), ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from test_parity_gpu import autocast_floor, rel, train_step  # noqa: E402

out = {}
for arch, golden in (("ResNet18", "backbone_resnet18.npz"), ("ResNet50", "backbone_resnet50.npz")):
    g, net, logits, loss, x = train_step(arch, golden)
    norms = np.array([float(p.grad.double().norm()) for p in net.parameters()])
    ratio = norms / np.maximum(g["grad_norms"], 1e-12)
    big = g["grad_norms"] > 1e-3 * g["grad_norms"].max()
    cos = lambda a, b: float((a.flatten() @ b.flatten()) / (a.norm() * b.norm()))  # noqa: E731
    out[arch] = {"logits_rel_of_scale": rel(logits.detach().cpu(), g["train_logits"]),
                 "loss_rel": abs(loss.item() - float(g["train_loss"])) / abs(float(g["train_loss"])),
                 "grad_norm_ratio_min_med_max": [float(ratio[big].min()), float(np.median(ratio[big])), float(ratio[big].max())],
                 "cos_grad_linear": cos(net.linear.weight.grad.double().cpu(), torch.as_tensor(g["grad_linear_weight"]).double()),
                 "cos_grad_conv1": cos(net.conv1.weight.grad.double().cpu(), torch.as_tensor(g["grad_conv1_weight"]).double()),
                 "batch": int(x.shape[0])}
    fl, floss = autocast_floor(arch, g)
    out[arch]["eager_bf16_autocast_logits_rel_of_scale"] = rel(fl.cpu(), g["train_logits"])
    out[arch]["eager_bf16_autocast_loss_rel"] = abs(floss - float(g["train_loss"])) / abs(float(g["train_loss"]))
    if "eval_logits" in g.files:
        net.eval()
        with torch.no_grad():
            out[arch]["eval_logits_rel_of_scale"] = rel(net(x).cpu(), g["eval_logits"])
print(json.dumps(out, indent=1))
