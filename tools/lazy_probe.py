"""Wall time of ONE eager forward through the torchsparse-style API (no geometry prefetch, no CUDA graph): first call,
repeated calls on the same and on fresh SparseTensors (kernel maps and conv schedules rebuilt), fused blocks, and a
torch.profiler table.  KITTI-shaped scan (BASELINE configs[0]).   python tools/lazy_probe.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FT3D_CONV", "tc")
import fusiontransformer_b200 as ft
from fusiontransformer_b200 import dataflow
from fusiontransformer_b200.spvcnn import Net3DSeg
from fusiontransformer_b200.synthetic import make_scan
scan = make_scan("kitti", 41)
torch.manual_seed(5)
m = Net3DSeg(num_classes=20, dual_head=False, fusion="middle").cuda().eval()
db = dataflow.to_device(dataflow.host_batch_from_scans([scan]), torch.device("cuda", 0))
lidar, rc, bidx, labels, ginv, kept = dataflow.voxelize_batch(db)
img = torch.randn(lidar.C.shape[0], 96, device="cuda")
def run(x, tag):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        m(x, img)
    torch.cuda.synchronize(); print("%-28s %.2f ms" % (tag, 1e3 * (time.perf_counter() - t0)), flush=True)
run(lidar, "first call")
run(lidar, "same tensor again")
run(lidar, "same tensor again")
for i in range(3):
    run(ft.SparseTensor(lidar.F, lidar.C), "new tensor %d" % i)
ft.fuse(m)
run(ft.SparseTensor(lidar.F, lidar.C), "fused, new tensor")
run(ft.SparseTensor(lidar.F, lidar.C), "fused, new tensor")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(ft.SparseTensor(lidar.F, lidar.C), "profiled new tensor")
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
