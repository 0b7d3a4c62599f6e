"""Pin the oracle against the real reference and write tests/golden/*.npz.

Runs ONLY in the build container (needs /root/reference, which does not exist on the GPU box).
Usage:  python oracle/make_golden.py            (writes fixtures, exits non-zero on any mismatch)

What is checked (restatement == reference, exact unless stated):
  * state_dict key set / shapes / order of define_G(sr_fastdiffsr_test_64_256.json)
  * every schedule buffer of set_new_noise_schedule + sqrt_alphas_cumprod_prev
  * UNet forward on the oracle's deterministic weights (max abs diff reported, must be <= 1e-5)
  * the 20-step super_resolution with injected noise, B=1, continous False and True (exact layout)
  * PIL bicubic on the reference's own UC-Merced fixtures lr_128 -> sr_128_512 (bit exact, all 7)
    and against Pillow for 64->256 and 32->256 on synthetic LR.
Fixtures written are the REFERENCE's outputs, so tests elsewhere compare against the reference,
not against this file's own arithmetic.
"""
import glob
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/FastDiffSR"
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import fdsr_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load_ref_cfg(name):
    txt = "".join(line.split("//")[0] + "\n" for line in open(os.path.join(REF, "config", name)))
    opt = json.loads(txt)
    opt["distributed"] = False
    return opt


def ref_netG(opt, sd, device="cpu"):
    import model.networks as networks
    netG = networks.define_G(opt)
    netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], device)
    missing = netG.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.split(".")[0] != "denoise_fn" for k in missing.missing_keys), missing
    netG.eval()
    return netG


def run_ref_sampler(netG, cond, noises, continous):
    """Drive the reference's own super_resolution with injected noise by patching the two RNG
    call sites of its diffusion module (diffusion.py:207 and :189)."""
    import model.fastdiffsr_modules.diffusion as D
    queue = [noises[i] for i in range(noises.shape[0])]

    class _T:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def randn(shape, device=None):
            return queue.pop(0).clone()

        @staticmethod
        def randn_like(x):
            return queue.pop(0).clone()

    real = D.torch
    D.torch = _T()
    try:
        out = netG.super_resolution(cond, continous)
    finally:
        D.torch = real
    assert len(queue) == 0, "reference consumed %d fewer noise tensors than provided" % len(queue)
    return out


def run_ref_sampler_sr3(netG, cond, noises, continous):
    """The SR3 baseline's sampler (ddpm_modules/diffusion.py:201-231) draws torch.randn for x_T and then,
    through noise_like (:70-76), one torch.randn per step including t = 0 (where it is masked out)."""
    import model.ddpm_modules.diffusion as D
    queue = [noises[i] for i in range(noises.shape[0])] + [torch.zeros_like(noises[0])]

    class _T:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def randn(shape, device=None):
            return queue.pop(0).clone()

    real = D.torch
    D.torch = _T()
    try:
        out = netG.super_resolution(cond, continous)
    finally:
        D.torch = real
    assert len(queue) == 0, "reference consumed %d fewer noise tensors than provided" % len(queue)
    return out


def sr3_noise(T, B, H, seed):
    """Injected noise of the SR3 fixtures: regenerated from the seed by the tests (kept out of the fixtures
    for size); the fixture stores a checksum."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(T, B, 3, H, H, generator=g)


def main_sr3():
    """SR3 baseline (which_model_G = 'ddpm', config/sr_ddpm_test_64_256.json): state_dict surface, UNet forward
    (two configurations: the shipped one — attention at 16x16 of a 256 image — and image_size 64 so that the
    fixture's 64x64 input runs attention on 16x16 = 256 tokens with 128 channels), and the sampler with a
    short T = 12 linear schedule, all against the real reference."""
    import model.networks as networks
    ok = True
    opt = load_ref_cfg("sr_ddpm_test_64_256.json")
    ucfg = dict(opt["model"]["unet"])
    ucfg["norm_groups"] = 32
    for k in O.SR3_UNET:
        assert O.SR3_UNET[k] == ucfg[k], k
    assert opt["model"]["beta_schedule"]["val"] == O.SR3_SCHEDULE, opt["model"]["beta_schedule"]["val"]
    for image_size in (256, 64):
        opt["model"]["diffusion"]["image_size"] = image_size
        torch.manual_seed(0)
        netG0 = networks.define_G(opt)
        ref_keys = [(k, tuple(v.shape)) for k, v in netG0.state_dict().items()]
        my_keys = [(k, s) for (k, s, _, _) in O.sr3_state_dict_spec(ucfg, image_size)]
        assert ref_keys == my_keys, "sr3 state_dict spec mismatch"
        print(f"sr3[image_size={image_size}] state_dict spec: {len(my_keys)} tensors identical")
        sd = O.make_state_dict(ucfg, seed=3, gn_jitter=0.2, spec=O.sr3_state_dict_spec(ucfg, image_size))
        assert torch.equal(sd["denoise_fn.time_mlp.0.inv_freq"], netG0.state_dict()["denoise_fn.time_mlp.0.inv_freq"])
        sched = dict(schedule="linear", n_timestep=12, linear_start=1e-4, linear_end=0.35)
        netG = networks.define_G(opt)
        netG.set_new_noise_schedule(sched, "cpu")
        netG.load_state_dict(sd, strict=False)
        netG.eval()
        tab = O.schedule_tables(O.make_beta_schedule(**sched))
        H = 128 if image_size == 256 else 64
        g = torch.Generator().manual_seed(21)
        x6 = torch.randn(2, 6, H, H, generator=g).half().float()   # stored as fp16 (exactly representable)
        steps = [11, 640]   # time is embedded as a float: any integer step of a T=1000 schedule is a valid input
        eps_ref = []
        for t in steps:
            tt = torch.full((2,), t, dtype=torch.long)
            with torch.no_grad():
                r = netG.denoise_fn(x6, tt)
            m = O.sr3_unet_forward(sd, ucfg, x6, tt, image_size)
            d = (r - m).abs().max().item()
            print(f"sr3 unet[image_size={image_size}] t={t}: max|ref-oracle| = {d:.3e}, |ref| max {r.abs().max():.3f}")
            ok &= d <= 1e-5
            eps_ref.append(r.numpy())
        cond = torch.rand(1, 3, H, H, generator=g) * 2 - 1
        noises = sr3_noise(12, 1, H, 22)
        sr_ref = run_ref_sampler_sr3(netG, cond, noises, False)
        sr_ref_c = run_ref_sampler_sr3(netG, cond, noises, True)
        sr_m = O.sr3_sample_loop(sd, ucfg, tab, cond, noises, image_size, False)
        sr_m_c = O.sr3_sample_loop(sd, ucfg, tab, cond, noises, image_size, True)
        d1 = (sr_ref - sr_m[0]).abs().max().item()
        d2 = (sr_ref_c - sr_m_c).abs().max().item()
        print(f"sr3 sampler[image_size={image_size}] T=12 {H}x{H}: max|ref-oracle| = {d1:.3e} "
              f"(continous: {d2:.3e}, shape {tuple(sr_ref_c.shape)}; plain returns {tuple(sr_ref.shape)})")
        ok &= d1 <= 1e-4 and d2 <= 1e-4 and sr_ref_c.shape == sr_m_c.shape
        np.savez_compressed(os.path.join(GOLD, f"sr3_{image_size}.npz"), x6=x6.numpy().astype(np.float16),
                            steps=np.array(steps), eps=np.stack(eps_ref), cond=cond.numpy(),
                            noise_seed=22, noise_sum=float(noises.double().sum()), sr=sr_ref.numpy(),
                            sr_continous=sr_ref_c[::4].numpy(), image_size=image_size,   # frames 0, 4, 8, 12
                           
                            sched=json.dumps(sched), gn_jitter=0.2, seed=3)
    return ok


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    opt = load_ref_cfg("sr_fastdiffsr_test_64_256.json")
    ucfg = dict(opt["model"]["unet"])
    ucfg["norm_groups"] = 32
    ok = True

    # ---- state_dict surface
    import model.networks as networks
    torch.manual_seed(0)
    netG0 = networks.define_G(opt)
    ref_keys = [(k, tuple(v.shape)) for k, v in netG0.state_dict().items()]
    my_keys = [(k, s) for (k, s, _, _) in O.state_dict_spec(ucfg)]
    assert ref_keys == my_keys, "state_dict spec mismatch"
    print("state_dict spec: %d tensors identical (order, names, shapes)" % len(my_keys))

    # ---- schedule
    sched_opt = opt["model"]["beta_schedule"]["val"]
    betas = O.make_beta_schedule(**sched_opt)
    tab = O.schedule_tables(betas)
    netG0.set_new_noise_schedule(sched_opt, "cpu")
    gold_sched = {}
    for k in ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
              "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"]:
        r = getattr(netG0, k).numpy()
        gold_sched[k] = r
        assert np.array_equal(r, tab[k].astype(np.float32)), k
    gold_sched["sqrt_alphas_cumprod_prev"] = netG0.sqrt_alphas_cumprod_prev
    assert np.array_equal(netG0.sqrt_alphas_cumprod_prev, tab["sqrt_alphas_cumprod_prev"])
    for name in ["linear", "quad", "const", "jsd", "warmup10", "warmup50"]:
        import model.fastdiffsr_modules.diffusion as D
        a = D.make_beta_schedule(name, 20, 1e-4, 2e-2)
        assert np.array_equal(a, O.make_beta_schedule(name, 20, 1e-4, 2e-2)), name
    np.savez(os.path.join(GOLD, "schedule_T20.npz"), **gold_sched)
    print("schedule: 13 tables bit-identical")

    # ---- UNet forward + sampler at 64x64 (B=2 for eps, B=1 for the loop), jittered GN affine
    for tag, jitter in (("default", 0.0), ("jitter", 0.2)):
        sd = O.make_state_dict(ucfg, seed=0, gn_jitter=jitter)
        netG = ref_netG(opt, sd)
        g = torch.Generator().manual_seed(11)
        H = 64
        x6 = torch.randn(2, 6, H, H, generator=g)
        nls = [tab["sqrt_alphas_cumprod_prev"][20], tab["sqrt_alphas_cumprod_prev"][8]]
        eps_ref = []
        for nl in nls:
            nlt = torch.full((2, 1), float(np.float32(nl)))
            with torch.no_grad():
                r = netG.denoise_fn(x6, nlt)
            m = O.unet_forward(sd, ucfg, x6, nlt)
            d = (r - m).abs().max().item()
            print(f"unet[{tag}] nl={nl:.4g}: max|ref-oracle| = {d:.3e}, |ref| max {r.abs().max():.3f}")
            ok &= d <= 1e-5
            eps_ref.append(r.numpy())
        cond = torch.rand(1, 3, H, H, generator=g) * 2 - 1
        noises = torch.randn(20, 1, 3, H, H, generator=g)
        sr_ref = run_ref_sampler(netG, cond, noises, False)
        sr_ref_c = run_ref_sampler(netG, cond, noises, True)
        trace = []
        sr_m = O.sample_loop(sd, ucfg, tab, cond, noises, False, trace=trace)
        sr_m_c = O.sample_loop(sd, ucfg, tab, cond, noises, True)
        d1 = (sr_ref - sr_m).abs().max().item()
        d2 = (sr_ref_c - sr_m_c).abs().max().item()
        print(f"sampler[{tag}] T=20 64x64: max|ref-oracle| = {d1:.3e} (continous: {d2:.3e}, shape {tuple(sr_ref_c.shape)})")
        ok &= d1 <= 1e-4 and d2 <= 1e-4 and sr_ref_c.shape == sr_m_c.shape
        np.savez(os.path.join(GOLD, f"unet64_{tag}.npz"), x6=x6.numpy(), noise_levels=np.array(nls, dtype=np.float32),
                 eps=np.stack(eps_ref), cond=cond.numpy(), noises=noises.numpy(),
                 sr=sr_ref.numpy(), sr_continous=sr_ref_c.numpy(), gn_jitter=jitter, seed=0,
                 eps_first=trace[0]["eps"].numpy(), eps_last=trace[-1]["eps"].numpy())

    # ---- bicubic: the reference's own fixtures
    from PIL import Image
    lr_files = sorted(glob.glob(os.path.join(REF, "dataset/UCM_128_512/lr_128/*.tif")))
    assert len(lr_files) == 7
    keep = {}
    for i, f in enumerate(lr_files):
        lr = np.array(Image.open(f).convert("RGB"))
        sr = np.array(Image.open(f.replace("lr_128", "sr_128_512")).convert("RGB"))
        mine = O.pil_bicubic_u8(lr, 512, 512)
        nbad = int((mine != sr).sum())
        print(f"bicubic UCM {os.path.basename(f)}: {nbad} differing bytes")
        ok &= nbad == 0
        if i < 2:
            keep[f"lr{i}"] = lr
            keep[f"sr{i}"] = sr
    rng = np.random.default_rng(5)
    for (h, H) in ((64, 256), (32, 256)):
        lr = rng.integers(0, 256, size=(h, h, 3), dtype=np.uint8)
        pil = np.array(Image.fromarray(lr).resize((H, H), Image.BICUBIC))
        mine = O.pil_bicubic_u8(lr, H, H)
        nbad = int((mine != pil).sum())
        print(f"bicubic synthetic {h}->{H}: {nbad} differing bytes vs Pillow")
        ok &= nbad == 0
        keep[f"syn_lr_{h}"] = lr
        keep[f"syn_sr_{h}"] = pil
    np.savez_compressed(os.path.join(GOLD, "bicubic.npz"), **keep)

    ok &= main_sr3()
    print("GOLDEN", "OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
