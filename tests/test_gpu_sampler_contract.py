"""Contracts of the sampler around the kernels (through the C ABI; needs a B200: `pytest -m gpu`):
the built-in Gaussian generator (distribution, identity with the injected-noise path, independence from
batch composition / rank count), the per-shape CUDA-graph cache, weight validation at the public boundary,
and the fp16 overflow guard.  Reference behaviour being matched: diffusion.py:157-221 (p_sample_loop draws
torch.randn per step; any N(0,1) stream is a valid sample) and model/model.py:148-160 (strict load)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(oracle, schedule):
    from fastdiffsr_b200 import Engine
    cfg = dict(oracle.DEFAULT_UNET)
    e = Engine(cfg, "cuda:0", "fp16")
    e.load_state_dict(oracle.make_state_dict(cfg, seed=0))
    e.set_schedule(schedule["betas"])
    yield e
    e.close()


def test_builtin_noise_is_standard_normal(eng):
    """Moments, a Kolmogorov-Smirnov test against N(0,1) and independence across streams / images / seeds of the
    Philox generator that feeds x_T and every z_t when no noise is injected (only determinism was tested)."""
    from scipy import stats
    B, H, W = 4, 256, 256
    a = eng.debug_noise(B, H, W, seed=1234, stream_id=20).cpu().double().numpy()       # x_T stream
    b = eng.debug_noise(B, H, W, seed=1234, stream_id=19).cpu().double().numpy()       # z of step 19
    c = eng.debug_noise(B, H, W, seed=1235, stream_id=20).cpu().double().numpy()       # another seed
    n = a.size
    assert n == 4 * 3 * 256 * 256
    for v in (a, b, c):
        f = v.ravel()
        assert abs(f.mean()) < 5.0 / np.sqrt(n)                       # 5 sigma of the sample mean
        assert abs(f.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
        assert abs(stats.skew(f)) < 5.0 * np.sqrt(6.0 / n)
        assert abs(stats.kurtosis(f)) < 5.0 * np.sqrt(24.0 / n)
        ks = stats.kstest(f[:200000], "norm")
        assert ks.pvalue > 1e-3, ks
        assert np.abs(f).max() < 6.5 and np.abs(f).max() > 4.0        # tails exist, nothing absurd
    # independence: streams, seeds, neighbouring images and neighbouring elements are uncorrelated
    bound = 5.0 / np.sqrt(n / 4)
    cc = lambda x, y: float(np.corrcoef(x.ravel(), y.ravel())[0, 1])
    assert abs(cc(a, b)) < bound and abs(cc(a, c)) < bound
    assert abs(cc(a[0], a[1])) < 2 * bound and abs(cc(a[..., :-1], a[..., 1:])) < bound
    # the four values of one Philox block (Box-Muller pairs) are uncorrelated too
    q = a.reshape(-1, 4)
    for i in range(4):
        for j in range(i + 1, 4):
            assert abs(cc(q[:, i], q[:, j])) < 2 * bound


def test_builtin_noise_equals_injected_noise_path(eng):
    """The z the posterior kernel generates in registers IS the debug hook's stream: sampling with the built-in
    generator equals, bit for bit, sampling with that noise injected as a tensor (draw order: x_T, z_19 .. z_1)."""
    g = torch.Generator().manual_seed(3)
    B, H, W, T = 2, 64, 96, 20
    cond = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).cuda()
    seed, off = 77, 5
    noises = torch.stack([eng.debug_noise(B, H, W, seed, T, image_offset=off)] +
                         [eng.debug_noise(B, H, W, seed, t, image_offset=off) for t in range(T - 1, 0, -1)]).contiguous()
    assert tuple(noises.shape) == (T, B, 3, H, W)                     # x_T, z_19 .. z_1 (t = 0 adds no noise)
    a = eng.sample(cond, seed=seed, image_offset=off)
    b = eng.sample(cond, noise=noises)
    assert torch.equal(a, b)


def test_builtin_noise_is_independent_of_batch_composition(eng):
    """Image k of a job gets the same noise whether it is sampled in a batch of 4, alone, or as part of another
    shard (N-rank == 1-rank bitwise): the Philox counter is the GLOBAL image index."""
    g = torch.Generator().manual_seed(4)
    cond = (torch.rand(4, 3, 64, 64, generator=g) * 2 - 1).cuda()
    full = eng.sample(cond, seed=9)                                    # one "rank"
    lo = eng.sample(cond[:2].contiguous(), seed=9, image_offset=0)     # two "ranks"
    hi = eng.sample(cond[2:].contiguous(), seed=9, image_offset=2)
    assert torch.equal(torch.cat([lo, hi]), full)
    one = eng.sample(cond[3:].contiguous(), seed=9, image_offset=3)    # four "ranks", last one
    assert torch.equal(one[0], full[3])
    assert not torch.equal(eng.sample(cond[3:].contiguous(), seed=9, image_offset=0)[0], full[3])
    # the host-buffer path honours the offset as well
    lr = np.random.default_rng(1).integers(0, 256, size=(4, 16, 16, 3), dtype=np.uint8)
    whole = eng.super_resolve_u8_host(lr, 64, 64, seed=5)
    part = eng.super_resolve_u8_host(lr[2:], 64, 64, seed=5, image_offset=2)
    assert np.array_equal(whole[2:], part)


def test_sharded_super_resolution_is_world_size_invariant(oracle, schedule):
    """fastdiffsr_b200.parallel.sharded_super_resolution with a fake 2-rank world (no process group: the shard of
    each rank is computed in turn on the one GPU) gathers to exactly the single-rank result."""
    import fastdiffsr_b200 as F
    from fastdiffsr_b200 import parallel as P
    opt = F.config.default_config()
    netG = F.define_G(opt)
    netG.load_state_dict(oracle.make_state_dict(oracle.DEFAULT_UNET, seed=0), strict=False)
    netG.to("cuda")
    netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
    g = torch.Generator().manual_seed(6)
    cond = (torch.rand(5, 3, 64, 64, generator=g) * 2 - 1).cuda()       # 5 images over 2 ranks: 3 + 2 (padded)
    single = P.sharded_super_resolution(netG, cond, seed=11)
    parts = []
    for rank in range(2):
        local, n_valid = P.shard_batch(cond, rank, 2)
        start, _, _ = P.shard_bounds(5, rank, 2)
        parts.append(netG.super_resolution(local, False, seed=11, image_offset=start)[:n_valid])
    assert torch.equal(torch.cat(parts), single)


def test_graph_cache_survives_fresh_tensors_and_alternating_shapes(eng):
    """One capture per (B, H, W, injected?, trace?): new seeds, new noise / trace tensors, a ragged last batch followed
    by a full one and alternating image shapes all replay cached graphs."""
    g = torch.Generator().manual_seed(8)
    c64 = (torch.rand(3, 3, 64, 64, generator=g) * 2 - 1).cuda()
    c96 = (torch.rand(2, 3, 64, 96, generator=g) * 2 - 1).cuda()
    eng.set_use_graph(True)
    ref64 = eng.sample(c64, seed=1)
    ref96 = eng.sample(c96, seed=1)
    ref2 = eng.sample(c64[:2].contiguous(), seed=1)                     # "ragged last batch"
    n0 = eng.graph_captures()
    for i in range(3):
        assert torch.equal(eng.sample(c64, seed=1), ref64)
        assert torch.equal(eng.sample(c96, seed=1), ref96)
        assert torch.equal(eng.sample(c64[:2].contiguous(), seed=1), ref2)
        assert not torch.equal(eng.sample(c64, seed=2 + i), ref64)
    assert eng.graph_captures() == n0
    # injected noise: a fresh tensor (new pointer) per call re-uses one graph
    outs = []
    for i in range(3):
        nz = torch.randn(20, 3, 3, 64, 64, generator=torch.Generator().manual_seed(5)).cuda()
        junk = torch.empty(1000 * (i + 1), device="cuda")               # perturb the allocator
        outs.append(eng.sample(c64, noise=nz))
        del junk
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert eng.graph_captures() == n0 + 1
    sr, tr1 = eng.sample(c64, noise=nz, trace=True)
    sr2, tr2 = eng.sample(c64, noise=nz.clone(), trace=True)
    assert torch.equal(tr1, tr2) and torch.equal(sr, outs[0]) and eng.graph_captures() == n0 + 2
    # graph replay == plain stream launches
    eng.set_use_graph(False)
    assert (eng.sample(c64, seed=1) - ref64).abs().max().item() <= 1e-5
    eng.set_use_graph(True)


def test_load_weights_rejects_mismatched_state_dicts(oracle, schedule):
    """the C ABI must refuse, by tensor name, a state_dict that does not fit the configured network
    (different inner_channel / channel_mults, truncated Linear / FiLM / CLAM / bias tensors) instead of reading past the
    host arrays."""
    from fastdiffsr_b200 import Engine, FdsrError
    cfg = dict(oracle.DEFAULT_UNET)
    good = oracle.make_state_dict(cfg, seed=0)
    e = Engine(cfg, "cuda:0", "fp16")
    wide = dict(cfg, inner_channel=128)
    with pytest.raises(FdsrError, match="needs"):
        e.load_state_dict(oracle.make_state_dict(wide, seed=0))
    other = dict(cfg, channel_multiplier=[1, 2, 2, 4])
    with pytest.raises(FdsrError, match="needs"):
        e.load_state_dict(oracle.make_state_dict(other, seed=0))
    for key, new in (("denoise_fn.noise_level_mlp.1.weight", torch.zeros(64, 64)),
                     ("denoise_fn.downs.1.res_block.noise_func.noise_func.0.weight", torch.zeros(64, 32)),
                     ("denoise_fn.mid.0.ca.fc1.weight", torch.zeros(8, 256, 1, 1)),
                     ("denoise_fn.downs.4.res_block.res_conv.weight", torch.zeros(128, 32, 1, 1)),
                     ("denoise_fn.final_conv.block.3.bias", torch.zeros(4)),
                     ("denoise_fn.downs.2.res_block.block1.block.0.bias", torch.zeros(32))):
        bad = dict(good)
        bad[key] = new
        with pytest.raises(FdsrError, match=key.replace(".", r"\.")):
            e.load_state_dict(bad)
    missing = {k: v for k, v in good.items() if k != "denoise_fn.ups.3.conv.bias"}
    with pytest.raises(FdsrError, match="missing"):
        e.load_state_dict(missing)
    # a failed load leaves the context unusable-but-safe (no stale half-packed weights)...
    with pytest.raises(FdsrError):
        e.unet_forward(torch.zeros(1, 3, 64, 64, device="cuda"), torch.zeros(1, 3, 64, 64, device="cuda"), 0)
    # ...and a good state_dict afterwards works
    e.load_state_dict(good)
    e.set_schedule(schedule["betas"])
    assert torch.isfinite(e.unet_forward(torch.zeros(1, 3, 64, 64, device="cuda"),
                                         torch.zeros(1, 3, 64, 64, device="cuda"), 3)).all()
    e.close()


def _scaled_level0(sd, k):
    """Every conv that writes into the full-resolution residual stream (stem, downs.1 / downs.2 block2) scaled by k:
    that stream — downs.0, downs.1, downs.2 and the three skip tensors — grows exactly k-fold while GroupNorm keeps
    every conv INPUT normalised: the situation of a trained network whose un-normalised residual stream is large
    (ADVICE r1 medium).  Up to GroupNorm's eps the first level is scale-equivariant, so its relative accuracy must not
    depend on k as long as the storage format holds the values."""
    out = dict(sd)
    for name in ("denoise_fn.downs.0", "denoise_fn.downs.1.res_block.block2.block.3",
                 "denoise_fn.downs.2.res_block.block2.block.3"):
        out[name + ".weight"] = sd[name + ".weight"] * k
        out[name + ".bias"] = sd[name + ".bias"] * k
    return out


def test_fp16_overflow_is_detected_and_bf16_handles_it(oracle, schedule):
    from fastdiffsr_b200 import Engine, FdsrOverflowError
    cfg = dict(oracle.DEFAULT_UNET)
    base = oracle.make_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(12)
    cond = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    x = torch.randn(1, 3, 64, 64, generator=g)
    t = 6
    nl = torch.full((1, 1), float(np.float32(schedule["sqrt_alphas_cumprod_prev"][t + 1])))
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()

    def run(dtype, sd):
        taps = {}
        oracle.unet_forward(sd, cfg, torch.cat([cond, x], 1), nl, taps=taps)
        e = Engine(cfg, "cuda:0", dtype)
        e.load_state_dict(sd)
        e.set_schedule(schedule["betas"])
        eps = e.unet_forward(cond.cuda(), x.cuda(), t).cpu()
        got = e.read_tensor("downs.2", 1, taps["downs.2"].numel()).cpu()
        try:
            e.check_overflow()
            ovf = False
        except FdsrOverflowError:
            ovf = True
        e.close()
        return eps, got, taps["downs.2"], ovf

    # x100: a residual stream of a few hundred is inside the fp16 range: no flag, same relative accuracy as unscaled
    eps, got, want, ovf = run("fp16", _scaled_level0(base, 100.0))
    print(f"fp16 x100: max |downs.2| = {got.abs().max():.4g}, rel-L2 {rel(got, want):.3e}")
    assert not ovf and got.abs().max() > 50.0 and torch.isfinite(eps).all() and rel(got, want) <= 5e-3
    # x3e4: the stream (max ~1e5) leaves the fp16 range -> stored saturated (nothing turns inf / NaN) AND flagged loudly
    big = _scaled_level0(base, 3e4)
    eps, got, want, ovf = run("fp16", big)
    assert want.abs().max() > 65504.0
    assert ovf and torch.isfinite(eps).all() and torch.isfinite(got).all() and got.abs().max() <= 65504.0
    sat_err = rel(got, want)
    # the bf16 mode (fp32 exponent range in storage) computes the same network correctly
    eps, got, want, ovf = run("bf16", big)
    print(f"bf16 x3e4: max |downs.2| = {got.abs().max():.4g}, rel-L2 {rel(got, want):.3e} (fp16 saturated: {sat_err:.3e})")
    assert not ovf and got.abs().max() > 65504.0 and torch.isfinite(eps).all() and rel(got, want) <= 1e-2 < sat_err


def test_auto_dtype_falls_back_to_bf16(oracle, schedule):
    import warnings
    import fastdiffsr_b200 as F
    opt = F.config.default_config()
    opt["model"]["compute_dtype"] = "auto"
    netG = F.define_G(opt)
    netG.load_state_dict(_scaled_level0(oracle.make_state_dict(oracle.DEFAULT_UNET, seed=0), 3e4), strict=False)
    netG.to("cuda")
    netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
    g = torch.Generator().manual_seed(2)
    cond = (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).cuda()
    nz = torch.randn(20, 1, 3, 64, 64, generator=g).cuda()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sr = netG.super_resolution(cond, False, noise=nz)
    assert netG.compute_dtype == "bf16" and any("bf16" in str(m.message) for m in w)
    assert torch.isfinite(sr).all()
    # explicit fp16 raises instead
    opt["model"]["compute_dtype"] = "fp16"
    net16 = F.define_G(opt)
    net16.load_state_dict(netG.state_dict(), strict=False)
    net16.to("cuda")
    net16.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
    with pytest.raises(F.FdsrOverflowError):
        net16.super_resolution(cond, False, noise=nz)


def test_pipelined_host_path_equals_synchronous(eng):
    """fdsr_super_resolve_u8_submit / _wait (two slots, D2H on a copy stream) returns exactly what the synchronous call
    returns, batch by batch, and refuses a slot that still holds a result."""
    from fastdiffsr_b200 import FdsrError
    rng = np.random.default_rng(11)
    batches = [rng.integers(0, 256, size=(2, 16, 16, 3), dtype=np.uint8) for _ in range(4)]
    want = [eng.super_resolve_u8_host(b, 64, 64, seed=40 + i, image_offset=2 * i).copy() for i, b in enumerate(batches)]
    got = [np.empty((2, 3, 64, 64), dtype=np.float32) for _ in batches]
    eng.super_resolve_u8_submit(0, batches[0], 64, 64, seed=40, image_offset=0)
    with pytest.raises(FdsrError):
        eng.super_resolve_u8_submit(0, batches[1], 64, 64, seed=41)
    for i in range(1, 4):
        eng.super_resolve_u8_submit(i % 2, batches[i], 64, 64, seed=40 + i, image_offset=2 * i)
        eng.super_resolve_u8_wait((i - 1) % 2, got[i - 1])
    eng.super_resolve_u8_wait(1, got[3])
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    with pytest.raises(FdsrError):
        eng.super_resolve_u8_wait(0, got[0])
