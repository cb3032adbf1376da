import sys, torch
sys.path.insert(0, '/root/repo')
from pgdvs_b200 import synthetic
from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
dev = torch.device('cuda:0')
wl = synthetic.make_workload('c5_stress', dev, n_views=2, K=8, radius=0.01)
pairs, cams = wl.jobs(range(wl.n_views))
prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
for _ in range(2):
    out = render_prepared(prep, radius=wl.radius, points_per_pixel=wl.K, compositor='norm', static_rgb=wl.static_rgb, return_fragments=True)
torch.cuda.synchronize()
