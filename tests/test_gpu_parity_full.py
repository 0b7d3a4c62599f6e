"""Full-size parity of the CUDA path against the oracle (computed live on the GPU box's host cores): the shapes
BASELINE.json names, several weight seeds, both 16-bit modes.  Needs a B200: `pytest -m gpu`.

VERDICT r1 asked for: >= 3 weight seeds at 256^2 (SURVEY hard part 6 iii), one FULL 512^2 UNet step (only a strip
was compared), and a 256^2 20-step sampler assertion (it existed only inside bench.py's CPU leg).
Tolerances: per-step eps relative L2 <= 1e-2 (north_star), final SR within 0.05 dB PSNR of the reference result."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

EPS_TOL = 1e-2


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _engine(oracle, schedule, sd, dtype):
    from fastdiffsr_b200 import Engine
    e = Engine(dict(oracle.DEFAULT_UNET), "cuda:0", dtype)
    e.load_state_dict(sd)
    e.set_schedule(schedule["betas"])
    return e


@pytest.fixture(scope="module")
def step256(oracle, schedule):
    """Oracle eps of one 256^2 step for three weight seeds (seed, GroupNorm jitter), computed once for both dtypes."""
    cfg = dict(oracle.DEFAULT_UNET)
    out = []
    for seed, jitter, t in ((1, 0.0, 15), (2, 0.2, 8), (3, 0.1, 1)):
        sd = oracle.make_state_dict(cfg, seed=seed, gn_jitter=jitter)
        g = torch.Generator().manual_seed(100 + seed)
        cond = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
        x = torch.randn(1, 3, 256, 256, generator=g)
        nl = torch.full((1, 1), float(np.float32(schedule["sqrt_alphas_cumprod_prev"][t + 1])))
        ref = oracle.unet_forward(sd, cfg, torch.cat([cond, x], 1), nl)
        out.append((seed, sd, cond, x, t, ref))
    return out


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_three_weight_seeds_at_256(oracle, schedule, step256, dtype):
    for seed, sd, cond, x, t, ref in step256:
        e = _engine(oracle, schedule, sd, dtype)
        eps = e.unet_forward(cond.cuda(), x.cuda(), t).cpu()
        e.check_overflow()
        e.close()
        r = rel_l2(eps, ref)
        print(f"[{dtype}] weight seed {seed}, t={t}, 256x256: eps rel-L2 {r:.3e}")
        assert r <= EPS_TOL, (seed, r)


@pytest.fixture(scope="module")
def step512(oracle, schedule):
    cfg = dict(oracle.DEFAULT_UNET)
    sd = oracle.make_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(512)
    cond = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    x = torch.randn(1, 3, 512, 512, generator=g)
    t = 12
    nl = torch.full((1, 1), float(np.float32(schedule["sqrt_alphas_cumprod_prev"][t + 1])))
    return sd, cond, x, t, oracle.unet_forward(sd, cfg, torch.cat([cond, x], 1), nl)


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_full_512_step_vs_oracle(oracle, schedule, step512, dtype):
    """BASELINE configs[3] shape (x4 128 -> 512): one complete 512^2 UNet step against the oracle."""
    sd, cond, x, t, ref = step512
    e = _engine(oracle, schedule, sd, dtype)
    eps = e.unet_forward(cond.cuda(), x.cuda(), t).cpu()
    assert abs(e.unet_flops() / 1e9 - 1073.24) < 0.2            # SURVEY: 1073.24 GFLOP per image-step at 512^2
    e.close()
    r = rel_l2(eps, ref)
    print(f"[{dtype}] 512x512 eps rel-L2 {r:.3e}")
    assert r <= EPS_TOL


@pytest.fixture(scope="module")
def sampler256(oracle, schedule):
    """Oracle 20-step sampling of B = 2 images at 256^2 with injected noise (B > 1 = the reference per image)."""
    cfg = dict(oracle.DEFAULT_UNET)
    sd = oracle.make_state_dict(cfg, seed=0, gn_jitter=0.1)
    g = torch.Generator().manual_seed(256)
    lr = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8)
    cond = torch.stack([oracle.u8_to_cond(oracle.pil_bicubic_u8(lr[b].numpy(), 256, 256)[None])[0] for b in range(2)])
    noises = torch.randn(20, 2, 3, 256, 256, generator=g)
    ref = torch.cat([oracle.sample_loop(sd, cfg, schedule, cond[b:b + 1], noises[:, b:b + 1]) for b in range(2)])
    return sd, lr, cond, noises, ref


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_sampler_256_batch2_vs_oracle(oracle, schedule, sampler256, dtype):
    """BASELINE configs[0/1] shape: the whole path — bit-exact bicubic conditioning, 20 UNet steps, posterior updates,
    res2img — for a batch of two 64 -> 256 images against the oracle's result with the same injected noise."""
    sd, lr, cond, noises, ref = sampler256
    e = _engine(oracle, schedule, sd, dtype)
    _, cond_gpu = e.bicubic_u8(lr.cuda(), 256, 256, want_u8=False)
    assert torch.equal(cond_gpu.cpu(), cond)
    sr = e.sample(cond_gpu, noise=noises.cuda()).cpu()
    e.check_overflow()
    e.close()
    r = rel_l2(sr, ref)
    gen = torch.Generator().manual_seed(9)
    hr = (cond + 0.1 * torch.nn.functional.avg_pool2d(torch.randn(2, 3, 256, 256, generator=gen), 3, 1, 1)).clamp(-1, 1)
    for b in range(2):
        p_ours = oracle.psnr_u8(oracle.to_u8(sr[b]), oracle.to_u8(hr[b]))
        p_ref = oracle.psnr_u8(oracle.to_u8(ref[b]), oracle.to_u8(hr[b]))
        print(f"[{dtype}] image {b}: PSNR vs HR ours {p_ours:.4f} dB, oracle {p_ref:.4f} dB; "
              f"PSNR(ours, oracle) {oracle.psnr_u8(oracle.to_u8(sr[b]), oracle.to_u8(ref[b])):.1f} dB")
        assert abs(p_ours - p_ref) <= 0.05
    print(f"[{dtype}] 256x256 B=2 T=20 SR rel-L2 vs oracle {r:.3e}")
    assert r <= 1e-2
