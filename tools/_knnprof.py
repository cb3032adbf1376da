import sys, torch
sys.path.insert(0, '/root/repo')
from pgdvs_b200 import synthetic, ops
from pgdvs_b200.dyn_renderer import FilteredViews
from types import SimpleNamespace
from torch.profiler import ProfilerActivity, profile
dev = torch.device('cuda:0')
wl = synthetic.make_workload('c2_nvidia_seq', dev, n_views=24)
pairs, cams = wl.jobs(range(wl.n_views))
fv = FilteredViews(pairs, cams, wl.H, wl.W, dev, SimpleNamespace(dyn_pcl_outlier_knn=50, dyn_pcl_outlier_std_thres=0.1))
for _ in range(2):
    fv.filter()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fv.filter()
    torch.cuda.synchronize()
print("groups", fv.n_groups)
for evt in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]:
    print(f"{evt.key[:70]:70s} {evt.device_time_total/1e3:9.3f} ms x{evt.count}")
