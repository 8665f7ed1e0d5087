import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FT3D_CONV"] = sys.argv[1] if len(sys.argv) > 1 else "f32"
from tests.test_gpu_graph import _host_batches, _trainer
from fusiontransformer_b200 import dataflow, ops
from fusiontransformer_b200.graph import StaticGeometry

mode = os.environ["FT3D_CONV"]
hb = _host_batches()[0]
nets = []
for i in range(3):
    net, body = _trainer(mode, optimize=False)
    if nets:
        net.load_state_dict(nets[0][0].state_dict())
    nets.append((net, body))
losses = []
for i, (net, body) in enumerate(nets):
    plan = dataflow.prepare_batch(hb, "cuda")
    if i == 2:
        st = StaticGeometry(plan)
        st.load(plan)
        ops.ROW_COUNTS = st.row_counts
        plan = st.as_plan()
    losses.append(body(plan).item())
    ops.ROW_COUNTS = {}
print("losses", losses)
ref = dict(nets[0][0].named_parameters())
gmax = max(p.grad.norm().item() for p in ref.values())
for j in (1, 2):
    rows = []
    for name, p in nets[j][0].named_parameters():
        e = (p.grad - ref[name].grad).norm().item() / max(ref[name].grad.norm().item(), 1e-4 * gmax)
        rows.append((e, name))
    rows.sort(reverse=True)
    print("exact-vs-exact" if j == 1 else "exact-vs-padded", ["%.2e %s" % r for r in rows[:6]], "median %.2e" % sorted(r[0] for r in rows)[len(rows) // 2])
