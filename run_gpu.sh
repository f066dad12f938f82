set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 180 --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench_tree.py --quick > gpurun_out/tree_sweep.jsonl 2> gpurun_out/tree_sweep.err; python -c "
import sys, json
for l in open('gpurun_out/tree_sweep.jsonl'):
    r = json.loads(l); print(r['dataset'], r['rows'], r['C'], r['dtype'], r['kind'], '%.3f ms %.0f GB/s frac %.3f' % (r['ms'], r['gbs'], r['frac']))
"; tail -3 gpurun_out/tree_sweep.err
cat > /tmp/prof_tree.py <<'PY'
import sys, os
sys.path.insert(0, 'neural-backed-decision-trees_b200')
import torch
from nbdt.tree import Tree
from nbdt._b200 import ops
for ds, h, C, rows in (("CIFAR10", "induced", 10, 1 << 22), ("CIFAR100", "induced", 100, 1 << 19)):
    tree = Tree(ds, hierarchy=h)
    x = 3 * torch.randn(rows, C, device="cuda")
    y = torch.randint(C, (rows,), device="cuda")
    for _ in range(3):
        ops.tree_loss(tree, x, y, "soft", 1.0, 1.0)
    torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tree_kernel -c 6 -o gpurun_out/prof_tree -f python /tmp/prof_tree.py > gpurun_out/ncu_tree.log 2>&1; tail -3 gpurun_out/ncu_tree.log
