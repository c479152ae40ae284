"""GPU parity: RoIAlign (C-ABI nuhtc_roi_align_fwd through the mmcv-surface mirrors) vs the CPU oracle.
Tolerance: 1e-5 abs in fp32 (BASELINE.json north_star); the literal kernel must be bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _rois(K, B, frame, seed, edge=True):
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(K, 2, generator=g) * frame
    wh = 4 + torch.rand(K, 2, generator=g) * 90
    r = torch.cat([torch.randint(0, B, (K, 1), generator=g).float(), ctr - wh / 2, ctr + wh / 2], 1)
    if edge:
        r[0, 1:] = torch.tensor([10., 10., 10., 10.])            # zero-size
        r[1, 1:] = torch.tensor([-60., -60., frame + 60., frame + 60.])  # larger than the frame
        r[2, 1:] = torch.tensor([frame - 8., frame - 8., frame + 30., frame + 30.])  # hangs off the corner
        r[3, 1:] = torch.tensor([4., 4., 12., 12.])               # exactly on pixel centres at stride 4
        r[4, 1:] = torch.tensor([0., 0., float(frame), float(frame)])  # frame-filling (slow path in the fast kernel)
        r[5, 1:] = torch.tensor([-500., -500., -400., -400.])      # fully outside
        r[6, 1:] = torch.tensor([30., 40., 33., 200.])             # thin and tall
    return r


@pytest.mark.parametrize("C", [64, 256, 24])
@pytest.mark.parametrize("P,sr", [(7, 0), (7, 2), (14, 0), (14, 2)])
def test_single_level_matches_oracle(oracle, C, P, sr):
    import nuhtc_b200 as nb
    torch.manual_seed(1)
    B, H = 2, 32
    x = torch.randn(B, C, H, H)
    rois = _rois(96, B, H * 4, seed=P * 10 + sr)
    ref = oracle.roi_align(x, rois, P, 0.25, sr)
    out = nb.roi_align(x.cuda(), rois.cuda(), P, 0.25, sr, "avg", True).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= TOL
    # the literal kernel follows the reference accumulation order: bit-exact
    lit = nb.roi_align_levels([x.cuda()], rois.cuda(), P, [0.25], sr, True, impl="direct").cpu()
    assert torch.equal(lit, ref)


def test_module_surface_and_empty():
    import nuhtc_b200 as nb
    layer = nb.RoIAlign(output_size=7, spatial_scale=1 / 4, sampling_ratio=2)
    assert layer.output_size == (7, 7) and layer.sampling_ratio == 2
    x = torch.randn(1, 64, 16, 16, device="cuda")
    out = layer(x, torch.zeros(0, 5, device="cuda"))
    assert out.shape == (0, 64, 7, 7)
    with pytest.raises(nb.NuhtcError):
        nb.roi_align(x.cpu(), torch.zeros(1, 5), 7)


def test_channels_last_input_and_non_aligned(oracle):
    import nuhtc_b200 as nb
    torch.manual_seed(2)
    x = torch.randn(2, 64, 24, 40)
    rois = _rois(64, 2, 96, seed=5, edge=False)
    ref = oracle.roi_align(x, rois, 7, 0.25, 2)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    out = nb.roi_align(xc, rois.cuda(), 7, 0.25, 2).cpu()
    assert (out - ref).abs().max().item() <= TOL
    ref0 = oracle.roi_align(x, rois, (5, 3), 0.25, 0, aligned=False)
    out0 = nb.roi_align(x.cuda(), rois.cuda(), (5, 3), 0.25, 0, "avg", False).cpu()
    assert torch.equal(out0, ref0)


@pytest.mark.parametrize("C", [64, 256])
@pytest.mark.parametrize("P,sr", [(7, 0), (7, 2), (14, 0)])
def test_routed_levels_match_single_roi_extractor(oracle, C, P, sr):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    B = 2
    feats = synth.fpn_levels(B, C, frame=256, seed=3)
    rois = synth.proposals(B, 200, "routed", frame=256, seed=4)
    rois[:, 1:] *= 1.0  # 256 frame: sides up to 512 are clipped -> all four levels are exercised
    lv = oracle.map_roi_levels(rois, 4)
    assert len(torch.unique(lv)) >= 3
    ref = oracle.single_roi_extract(feats, rois, synth.FPN_STRIDES, P, sr)
    out = nb.roi_align_levels([f.cuda() for f in feats], rois.cuda(), P, [1 / s for s in synth.FPN_STRIDES], sr,
                              mode="route", finest_scale=56).cpu()
    assert (out - ref).abs().max().item() <= TOL
    lit = nb.roi_align_levels([f.cuda() for f in feats], rois.cuda(), P, [1 / s for s in synth.FPN_STRIDES], sr,
                              mode="route", finest_scale=56, impl="direct").cpu()
    assert torch.equal(lit, ref)


@pytest.mark.parametrize("P,sr", [(7, 2), (14, 0)])
def test_level_sum_matches_attention_extractor_roialign_branch(oracle, P, sr):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    B, C = 2, 64
    feats = synth.fpn_levels(B, C, frame=256, seed=6)[:2]
    rois = synth.proposals(B, 150, "nuclei", frame=256, seed=7)
    ref = oracle.sum_roi_extract(feats, rois, synth.FPN_STRIDES[:2], P, sr)
    out = nb.roi_align_levels([f.cuda() for f in feats], rois.cuda(), P, [1 / 4, 1 / 8], sr, mode="sum").cpu()
    assert (out - ref).abs().max().item() <= 2 * TOL  # two pooled terms are added


def _kernel_levels(rois, impl="auto", C=64, frame=512):
    """The level the KERNEL routes every RoI to: level l is a constant map of value l+1, so any bin of the routed output
    reads it back (the RoIs are kept inside the frame so that every sample lands in the map)."""
    import nuhtc_b200 as nb
    feats = [torch.full((1, C, frame // s, frame // s), float(l + 1), device="cuda") for l, s in enumerate((4, 8, 16, 32))]
    out = nb.roi_align_levels(feats, rois.cuda(), 7, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 2, mode="route", finest_scale=56, impl=impl)
    lv = out[:, 0, 3, 3].round().long() - 1
    # a bin of a constant map is that constant times a sum of bilinear weights (1 up to fp32 rounding)
    assert (out - (lv + 1).float()[:, None, None, None]).abs().max().item() < 1e-4
    return lv.cpu()


@pytest.mark.parametrize("impl", ["auto", "direct"])
def test_routing_matches_golden_levels(impl):
    """tests/golden/roi_levels.npz: levels produced by executing the reference's own map_roi_levels source
    (single_level_roi_extractor.py:36-55, make_golden.py) -- here against the routing INSIDE the CUDA kernels."""
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_levels.npz"))
    rois = torch.from_numpy(z["rois"]).clone()      # boundary boxes (0,0,112,112), (0,0,111.99,112), ... + random ones
    assert float(rois[:, 1:].min()) >= 0 and float(rois[:, 1:].max()) <= 512
    rois[:, 0] = 0
    got = _kernel_levels(rois, impl, frame=512)
    assert torch.equal(got, torch.from_numpy(z["levels"]))
    assert set(got.tolist()) == {0, 1, 2, 3}


@pytest.mark.parametrize("impl", ["auto", "direct"])
def test_routing_at_level_boundaries_every_float(impl):
    """map_roi_levels is floor(log2(v)) evaluated in fp32, so just below a power of two the logarithm can round up to the
    integer.  Sweep +-3000 consecutive floats of the box side around each boundary (v = 2, 4, 8) and compare the kernel's
    level with the reference expression evaluated by torch on the CPU (the contract: the reference's CPU path)."""
    rows = []
    for k in (1, 2, 3):
        side0 = torch.tensor(56.0 * (2.0 ** k - 1e-6), dtype=torch.float32)
        bits = side0.view(torch.int32) + torch.arange(-3000, 3001, dtype=torch.int32)
        side = bits.view(torch.float32)
        x1 = torch.zeros_like(side)      # x2 - x1 == side exactly
        rows.append(torch.stack([torch.zeros_like(side), x1, x1, x1 + side, x1 + side], 1))
        # rectangles: sqrt(w*h) falls between representable sides
        rows.append(torch.stack([torch.zeros_like(side), x1, x1, x1 + side, x1 + side0.expand_as(side).clone()], 1))
    rois = torch.cat(rows).contiguous()

    def levels(r):   # single_level_roi_extractor.py:51-55
        scale = torch.sqrt((r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2]))
        return torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(min=0, max=3).long()
    ref = levels(rois)
    got = _kernel_levels(rois, impl)
    assert torch.equal(got, ref)
    assert set(ref.tolist()) == {0, 1, 2, 3}
    v = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2])) / 56 + 1e-6
    exact = torch.floor(torch.log2(v.double())).clamp(min=0, max=3).long()
    dev = levels(rois.cuda()).cpu()
    print(f"routing sweep ({len(rois)} boxes): fp32-log2 level != exact-log2 level on {int((exact != ref).sum())}; "
          f"the reference's own CUDA evaluation differs from its CPU evaluation on {int((dev != ref).sum())}")


def test_full_size_launch_vs_oracle_and_literal(oracle):
    """The benched kernel instance itself (B=16, C=256, K=16000, 7x7): every RoI of the launch against the literal
    kernel (bit-exact with the oracle, see above), and a stride-40 sample of the SAME launch against the CPU oracle."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    B, C = 16, 256
    g = torch.Generator().manual_seed(0)
    a = torch.randn(B, C, 128, 128, generator=g)
    for dist in ("nuclei", "routed"):
        rois = synth.proposals(B, 1000, dist)
        ac = a.cuda()
        fast = nb.roi_align(ac, rois.cuda(), 7, 0.25, 0)
        lit = nb.roi_align_levels([ac], rois.cuda(), 7, [0.25], 0, impl="direct")
        assert fast.shape == (16000, 256, 7, 7)
        assert (fast - lit).abs().max().item() <= TOL
        idx = torch.arange(0, 16000, 40)
        ref = oracle.roi_align(a, rois[idx], 7, 0.25, 0, nthreads=8)
        assert (fast[idx.cuda()].cpu() - ref).abs().max().item() <= TOL
        assert torch.equal(lit[idx.cuda()].cpu(), ref)
    # the 14x14 mask-branch instance (K = 8000 detection slots)
    rois = synth.proposals(B, 500, "nuclei", seed=3)
    fast = nb.roi_align(ac, rois.cuda(), 14, 0.25, 0)
    idx = torch.arange(0, 8000, 40)
    ref = oracle.roi_align(a, rois[idx], 14, 0.25, 0, nthreads=8)
    assert (fast[idx.cuda()].cpu() - ref).abs().max().item() <= TOL


def test_full_size_linearity_property():
    """BASELINE cfg-2 size (B=16, C=256, K=16000): RoIAlign is linear in the feature map."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    B, C = 16, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(B, C, 128, 128, device="cuda", generator=g)
    b = torch.randn(B, C, 128, 128, device="cuda", generator=g)
    rois = synth.proposals(B, 1000, "nuclei").cuda()
    fa = nb.roi_align(a, rois, 7, 0.25, 0)
    fb = nb.roi_align(b, rois, 7, 0.25, 0)
    fab = nb.roi_align(a + 2 * b, rois, 7, 0.25, 0)
    assert fa.shape == (16000, 256, 7, 7)
    assert (fab - (fa + 2 * fb)).abs().max().item() <= 5e-5
    # and the fast kernel agrees with the literal kernel on a slice of the same launch shape
    sub = rois[::40].contiguous()
    assert (nb.roi_align(a, sub, 7, 0.25, 0) - nb.roi_align_levels([a], sub, 7, [0.25], 0, impl="direct")).abs().max().item() <= TOL


@pytest.mark.parametrize("P", [7, 14])
def test_strip_path_stress_sparse_repeated_concurrent(P):
    """The strip kernels are persistent producer / consumer pipelines (TMA ring, mbarriers): exercise the schedules that
    differ from the dense bench shape -- a handful of RoIs per image (consumers skip far more rows than the ring holds),
    all four levels with large windows (leftover list), back-to-back launches on a warm L2, and three streams at once --
    and hold every result to the literal kernel."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    B, C = 8, 64
    feats = [f.cuda() for f in synth.fpn_levels(B, C, frame=512, seed=5)]
    staged = nb.stage_levels(feats)
    scales = [1 / s for s in synth.FPN_STRIDES]
    cases = {"sparse": synth.proposals(B, 3, "nuclei", frame=512, seed=6), "routed": synth.proposals(B, 400, "routed", frame=512, seed=7),
             "nuclei": synth.proposals(B, 600, "nuclei", frame=512, seed=8)}
    refs = {}
    for name, r in cases.items():
        refs[name] = nb.roi_align_levels(feats, r.cuda(), P, scales, 0, mode="route", impl="direct")
    for name, r in cases.items():
        r = r.cuda()
        outs = [nb.roi_align_levels(staged, r, P, scales, 0, mode="route") for _ in range(6)]   # no sync in between
        for o in outs:
            assert (o - refs[name]).abs().max().item() <= TOL, name
    streams = [torch.cuda.Stream() for _ in range(3)]
    torch.cuda.synchronize()
    res = []
    for it in range(4):
        for s, (name, r) in zip(streams, cases.items()):
            with torch.cuda.stream(s):
                res.append((name, nb.roi_align_levels(staged, r.cuda(), P, scales, 0, mode="route")))
    torch.cuda.synchronize()
    for name, o in res:
        assert (o - refs[name]).abs().max().item() <= TOL, name


@pytest.mark.parametrize("seed", list(range(12)))
def test_shape_fuzz_default_path_vs_literal_kernel(seed):
    """Random batch sizes, channel counts, frame sizes (odd map sizes, several x-strips or one), output sizes, sampling ratios
    and box populations (nucleus-sized, map-sized, degenerate, partly or wholly outside the frame): the default path (strip
    kernels + pipelined leftovers on the channel-group layout, or the NHWC / direct kernels where it does not apply) against the
    literal per-sample kernel, which is bit-exact with the oracle (test_single_level_matches_oracle)."""
    import nuhtc_b200 as nb
    g = torch.Generator().manual_seed(1000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g).item())
    B = ri(1, 5)
    C = [32, 64, 96, 128, 256][ri(0, 4)]
    frame = 32 * ri(3, 20)
    strides = [4, 8, 16, 32]
    feats = [torch.randn(B, C, max(1, frame // s), max(1, frame // s), generator=g).cuda() for s in strides]
    P = [7, 14][ri(0, 1)]
    sr = [0, 2][ri(0, 1)]
    K = ri(1, 600)
    ctr = torch.rand(K, 2, generator=g) * frame * 1.2 - frame * 0.1
    kind = torch.rand(K, generator=g)
    side = torch.where(kind < 0.6, 8 + torch.rand(K, generator=g) * 72,                       # nuclei
                       torch.where(kind < 0.85, torch.exp(torch.rand(K, generator=g) * 4.2 + 2.3),   # 10 .. 660 px
                                   torch.rand(K, generator=g) * 2.0))                         # degenerate
    wh = side[:, None] * (0.5 + torch.rand(K, 2, generator=g))
    rois = torch.cat([torch.randint(0, B, (K, 1), generator=g).float(), ctr - wh / 2, ctr + wh / 2], 1).cuda()
    scales = [1.0 / s for s in strides]
    for mode in ("route", "sum"):
        lv = feats if mode == "route" else feats[:2]
        sc = scales if mode == "route" else scales[:2]
        a = nb.roi_align_levels(lv, rois, P, sc, sr, mode=mode)
        b = nb.roi_align_levels(lv, rois, P, sc, sr, mode=mode, impl="direct")
        tol = 1e-5 if mode == "route" else 2e-5
        err = (a - b).abs().max().item()
        assert err <= tol, (seed, mode, B, C, frame, P, sr, K, err)
