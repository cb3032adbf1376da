"""GPU parity of the softmax-splatting path against the oracle restatement of
pgdvs/utils/softsplat.py and pgdvs/renderers/pgdvs_renderer_base.py:59-138.

Floating-point atomics make the summation order run-dependent (upstream too), so everything
here is tolerance-based: |delta| <= 2e-5 on splatted sums of O(1) values."""
import numpy as np
import pytest
import torch

from oracle import pgdvs_ref as ref

pytestmark = pytest.mark.gpu

SPLAT_ATOL = 2e-5


def _dev():
    return torch.device("cuda:0")


def _rand(shape, g, scale=1.0):
    return torch.rand(shape, generator=g) * scale


@pytest.mark.parametrize("N,C,H,W", [(1, 1, 5, 7), (2, 4, 24, 40), (3, 3, 31, 17)])
def test_softsplat_forward_matches_oracle(N, C, H, W):
    import pgdvs_b200
    g = torch.Generator().manual_seed(N * 100 + C)
    x = _rand((N, C, H, W), g)
    flow = torch.randn((N, 2, H, W), generator=g) * 3.0
    flow[0, :, 0, 0] = float("nan")       # non-finite targets are skipped (softsplat.py:360-361)
    flow[0, 0, 1, 1] = float("inf")
    flow[-1, :, 2, 2] = torch.tensor([-100.0, 3.0])  # far outside: no valid tap
    out = pgdvs_b200.softsplat.softsplat_forward(x.to(_dev()), flow.to(_dev())).cpu().numpy()
    exp = ref.softsplat_forward(x.numpy(), flow.numpy())
    np.testing.assert_allclose(out, exp, atol=SPLAT_ATOL, rtol=0)


@pytest.mark.parametrize("mode", ["sum", "avg", "linear", "soft", "soft-zeroeps", "linear-clipeps"])
def test_softsplat_modes_match_oracle(mode):
    import pgdvs_b200
    g = torch.Generator().manual_seed(5)
    N, C, H, W = 2, 3, 20, 28
    x = _rand((N, C, H, W), g)
    flow = torch.randn((N, 2, H, W), generator=g) * 2.0
    metric = None if mode in ("sum", "avg") else (_rand((N, 1, H, W), g) * 2 - 1.5)
    if mode.startswith("linear"):
        metric = metric.abs() + 0.1
    d = _dev()
    out = pgdvs_b200.softsplat.softsplat(x.to(d), flow.to(d), metric.to(d) if metric is not None else None, mode)
    exp = ref.softsplat(x, flow, metric, mode)
    np.testing.assert_allclose(out.cpu().numpy(), exp.numpy(), atol=1e-4, rtol=1e-4)


def test_sum_mode_conserves_mass():
    import pgdvs_b200
    g = torch.Generator().manual_seed(9)
    x = _rand((1, 2, 64, 96), g)
    flow = _rand((1, 2, 64, 96), g) * 0.8 + 0.1
    flow[:, :, -1, :] = 0
    flow[:, :, :, -1] = 0  # every tap in bounds -> the splat only moves mass around
    out = pgdvs_b200.softsplat.softsplat(x.to(_dev()), flow.to(_dev()), None, "sum")
    assert abs(float(out.sum().cpu()) - float(x.sum())) < 1e-2 * 1e-1 * x.numel() ** 0.5


def test_metric_matches_reference_golden(golden_dir):
    """back-warp + L1 metric recorded from the REAL PGDVSBaseRenderer (tests/golden/make_golden.py)."""
    import pgdvs_b200
    gd = np.load(golden_dir / "softsplat_metric.npz")
    rgb1, rgb2, flow12 = (torch.from_numpy(gd[k]) for k in ("rgb1", "rgb2", "flow12"))
    B, _, H, W = rgb1.shape
    d = _dev()
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous().to(d)  # noqa: E731
    _, _, metric = pgdvs_b200.softsplat.softsplat_dyn(
        rgb_1=cl(rgb1), dyn_mask_1=torch.ones(B, H, W, 1, device=d), rgb_2=cl(rgb2),
        flow_1_to_tgt=torch.zeros(B, H, W, 2, device=d), flow_12=cl(flow12), return_metric=True)
    np.testing.assert_allclose(metric.cpu().numpy(), gd["metric"], atol=2e-6, rtol=0)


@pytest.mark.parametrize("with_noise", [False, True])
def test_softsplat_dyn_matches_oracle(with_noise):
    import pgdvs_b200
    g = torch.Generator().manual_seed(21)
    B, H, W = 2, 36, 52
    r1, r2 = _rand((B, H, W, 3), g), _rand((B, H, W, 3), g)
    m = (torch.rand((B, H, W, 1), generator=g) < 0.55).float()
    m[1] = 0  # a view without dynamic content
    f12 = torch.randn((B, H, W, 2), generator=g) * 2.0
    r2 = 0.7 * r2 + 0.3 * r1  # partially consistent colours: metrics spread over (0, 0.5)
    ft = torch.randn((B, H, W, 2), generator=g) * 3.0 * m
    noise = torch.clamp(torch.randn((B, H, W, 3), generator=g), 0.0, 1.0) if with_noise else None
    d = _dev()
    rgb, mask, metric = pgdvs_b200.softsplat.softsplat_dyn(
        rgb_1=r1.to(d), dyn_mask_1=m.to(d), rgb_2=r2.to(d), flow_1_to_tgt=ft.to(d), flow_12=f12.to(d),
        alpha=100.0, noise=noise.to(d) if noise is not None else None, return_metric=True)
    e_rgb, e_mask, e_metric = ref.softsplat_dyn_render(rgb_1=r1, dyn_mask_1=m, rgb_2=r2, flow_1_to_tgt=ft,
                                                       flow_12=f12, noise=noise, alpha=100.0)
    np.testing.assert_allclose(metric.cpu().numpy(), e_metric.numpy(), atol=2e-6, rtol=0)
    agree = (mask.cpu() == e_mask)
    assert agree.float().mean() > 0.999  # the 1e-3 threshold may flip on a rounding
    assert float(mask[1].sum()) == 0 and float(rgb[1].abs().sum()) == 0
    same = agree.expand_as(e_rgb)
    # exp(-100 * metric) amplifies the last bits of the metric by up to 100x before normalisation
    np.testing.assert_allclose(rgb.cpu().numpy()[same.numpy()], e_rgb.numpy()[same.numpy()], atol=2e-3, rtol=0)
    assert float((rgb.cpu() - e_rgb).abs()[same].mean()) < 2e-5


def test_renderer_forward_softsplat_vs_oracle():
    """PGDVSDynamicRenderer.forward with dyn_render_type='softsplat' on a reference-shaped data dict
    vs the oracle pipeline (compute_dyn_pcl -> compute_projections flow -> softsplat branch)."""
    import pgdvs_b200
    from types import SimpleNamespace
    d = _dev()
    B, H, W = 2, 24, 40
    g = torch.Generator().manual_seed(13)
    data = {
        "rgb_src_temporal": torch.rand(B, 2, H, W, 3, generator=g),
        "depth_src_temporal": 2 + 3 * torch.rand(B, 2, H, W, 1, generator=g),
        "dyn_mask_src_temporal": (torch.rand(B, 2, H, W, 1, generator=g) < 0.7).float(),
        "flow_fwd": 1.5 * torch.randn(B, H, W, 2, generator=g),
        "flow_fwd_occ_mask": (torch.rand(B, H, W, 1, generator=g) < 0.1).float(),
        "time_src_temporal": torch.tensor([[0.0, 1.0], [4.0, 5.0]]),
        "time_tgt": torch.tensor([[0.25], [4.5]]),
    }
    Kc = torch.eye(4)
    Kc[0, 0] = Kc[1, 1] = 0.9 * W
    Kc[0, 2], Kc[1, 2] = W / 2, H / 2

    def flat(tx):
        c2w = torch.eye(4)
        c2w[:3, 3] = torch.tensor([tx, 0.01, 0.0])
        return torch.cat([torch.tensor([float(H), float(W)]), Kc.reshape(-1), c2w.reshape(-1)])

    data["flat_cam_src_temporal"] = torch.stack([torch.stack([flat(0.0), flat(0.05)]),
                                                 torch.stack([flat(0.1), flat(0.15)])])
    data["flat_cam_tgt"] = torch.stack([flat(0.02), flat(0.13)])
    data["dyn_mask_src_temporal"][1, 0] = 0  # empty-mask view
    cfg = SimpleNamespace(dyn_render_type="softsplat", dyn_render_use_flow_consistency=True,
                          dyn_pcl_remove_outlier=False)
    noise = torch.clamp(torch.randn(B, H, W, 3, generator=g), 0.0, 1.0)
    r = pgdvs_b200.PGDVSDynamicRenderer(softsplat_metric_abs_alpha=100.0)
    rgb, mask, info = r({k: v.to(d) for k, v in data.items()}, None, cfg, softsplat_noise=noise.to(d))
    assert rgb.shape == (B, 3, H, W) and mask.shape == (B, 1, H, W)
    assert float(mask[1].sum()) == 0 and float(rgb[1].abs().sum()) == 0
    # oracle: flow frame 1 -> target from compute_dyn_pcl + compute_projections (:470-503)
    fs = data["flat_cam_src_temporal"]
    o = ref.compute_dyn_pcl(
        dyn_mask_1=data["dyn_mask_src_temporal"][0, 0], rgb_1=data["rgb_src_temporal"][0, 0],
        depth_1=data["depth_src_temporal"][0, 0], flow_12=data["flow_fwd"][0],
        flow_12_occ_mask=data["flow_fwd_occ_mask"][0], rgb_2=data["rgb_src_temporal"][0, 1],
        depth_2=data["depth_src_temporal"][0, 1], K_1=fs[0, 0, 2:18].reshape(4, 4), c2w_1=fs[0, 0, 18:34].reshape(4, 4),
        K_2=fs[0, 1, 2:18].reshape(4, 4), c2w_2=fs[0, 1, 18:34].reshape(4, 4), time_1=torch.tensor(0.0),
        time_2=torch.tensor(1.0), time_tgt=torch.tensor(0.25), use_flow_consistency=True)
    uv_t, _ = ref.compute_projections(o["pcl"], data["flat_cam_tgt"][0])
    sp = o["src_pix"].long()
    flow_t = torch.zeros(H * W, 2)
    flow_t[sp] = uv_t - torch.stack([(sp % W).float(), (sp // W).float()], dim=1)
    valid = torch.zeros(H * W, 1)
    valid[sp] = 1.0
    e_rgb, e_mask, _ = ref.softsplat_dyn_render(
        rgb_1=data["rgb_src_temporal"][:1, 0], dyn_mask_1=valid.view(1, H, W, 1), rgb_2=data["rgb_src_temporal"][:1, 1],
        flow_1_to_tgt=flow_t.view(1, H, W, 2), flow_12=data["flow_fwd"][:1], noise=noise[:1], alpha=100.0)
    agree = (mask[:1].cpu() == e_mask)
    assert agree.float().mean() > 0.995
    diff = (rgb[:1].cpu() - e_rgb).abs()[agree.expand_as(e_rgb)]
    assert float(diff.mean()) < 1e-4 and float((diff < 5e-3).float().mean()) > 0.99
