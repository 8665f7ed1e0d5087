import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FT3D_CONV"] = sys.argv[1] if len(sys.argv) > 1 else "f32"
from tests.test_gpu_graph import _host_batches, _trainer
import fusiontransformer_b200 as ft
from fusiontransformer_b200 import dataflow, ops
from fusiontransformer_b200.graph import StaticGeometry

mode = os.environ["FT3D_CONV"]
hb = _host_batches()[0]
net, _ = _trainer(mode, optimize=False)
g = torch.Generator(device="cuda").manual_seed(7)
plan = dataflow.prepare_batch(hb, "cuda")
n = plan.point_coords.shape[0]
img = torch.randn(n, 96, device="cuda", generator=g)
taps_e = {}
out_e = net(plan.extras["lidar"], img, taps=taps_e, plan=plan)["lidar_seg_logit"]
st = StaticGeometry(plan)
st.load(plan)
ops.ROW_COUNTS = st.row_counts
sp = st.as_plan()
P = st.n_points_cap
img_p = torch.zeros(P, 96, device="cuda")
img_p[:n] = img
img_p[n:] = 3.0          # garbage in the padding rows of an input
taps_p = {}
out_p = net(sp.extras["lidar"], img_p, taps=taps_p, plan=sp)["lidar_seg_logit"]
for k in taps_e:
    a, b = taps_e[k], taps_p[k]
    fa, fb = a.F, b.F
    m = fa.shape[0]
    d = (fa - fb[:m]).norm().item() / max(fa.norm().item(), 1e-12)
    pad = fb[m:].abs().max().item() if fb.shape[0] > m else 0.0
    print("%-4s rows %6d cap %6d rel diff %.2e  pad max %.2e" % (k, m, fb.shape[0], d, pad))
print("logits", ((out_e - out_p[:n]).norm() / out_e.norm()).item())
