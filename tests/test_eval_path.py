"""Evaluation path around the sampler: dataset / batched loader (CPU), the oracle's
restatement of the reference's metrics (CPU), device metrics and the batched evaluation loop (GPU)."""
import os

import numpy as np
import pytest
import torch

import fastdiffsr_b200 as F
from fastdiffsr_b200 import data as D


def _make_dataset(root, n=5, l=16, r=64, with_sr=False, seed=0):
    """hr_{r}/, lr_{l}/ (PIL BICUBIC downsample like data/prepare_data_mfe_dm.py:17-21) and optionally sr_{l}_{r}/."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    for d in (f"hr_{r}", f"lr_{l}") + ((f"sr_{l}_{r}",) if with_sr else ()):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    for i in range(n):
        base = rng.integers(0, 256, (r // 4, r // 4, 3), dtype=np.uint8)
        hr = Image.fromarray(base).resize((r, r), Image.BICUBIC)
        lr = hr.resize((l, l), Image.BICUBIC)
        ext = "tif" if i % 2 else "png"
        hr.save(os.path.join(root, f"hr_{r}", f"img{i:03d}.{ext}"))
        lr.save(os.path.join(root, f"lr_{l}", f"img{i:03d}.{ext}"))
        if with_sr:
            lr.resize((r, r), Image.BICUBIC).save(os.path.join(root, f"sr_{l}_{r}", f"img{i:03d}.{ext}"))
    return root


def _ssim_brute(a, b):
    """Independent per-window evaluation of the SSIM definition (7x7, sample covariance, 3-pixel crop)."""
    H, W, _ = a.shape
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    tot = []
    for ch in range(3):
        X, Y = a[..., ch].astype(np.float64), b[..., ch].astype(np.float64)
        acc, n = 0.0, 0
        for y in range(3, H - 3):
            for x in range(3, W - 3):
                wx, wy = X[y - 3:y + 4, x - 3:x + 4], Y[y - 3:y + 4, x - 3:x + 4]
                ux, uy = wx.mean(), wy.mean()
                vx, vy = wx.var(ddof=1), wy.var(ddof=1)
                vxy = ((wx - ux) * (wy - uy)).sum() / 48.0
                acc += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
                n += 1
        tot.append(acc / n)
    return float(np.mean(tot))


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_metrics_against_definitions(oracle):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (24, 31, 3), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-25, 26, a.shape), 0, 255).astype(np.uint8)
    assert abs(oracle.ssim_u8(a, b) - _ssim_brute(a, b)) < 1e-12
    assert oracle.ssim_u8(a, a) == pytest.approx(1.0, abs=1e-15)
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    assert oracle.mse_u8(a, b) == mse
    assert oracle.psnr_u8(a, b) == pytest.approx(10 * np.log10(255.0 ** 2 / mse), rel=1e-14)
    assert oracle.ergas_u8(a, b, 4) == pytest.approx(100 * np.sqrt(mse / np.mean(a, dtype=np.float64) ** 2 / 3) / 4)


def test_dataset_and_batch_loader(tmp_path):
    root = _make_dataset(str(tmp_path / "ds"), n=5, l=16, r=64, with_sr=True)
    ds = D.create_dataset(F.dict_to_nonedict({"dataroot": root, "datatype": "img", "l_resolution": 16,
                                              "r_resolution": 64, "mode": "LRHR", "data_len": -1}), "val")
    assert len(ds) == 5
    it = ds[2]
    assert it["HR"].shape == (3, 64, 64) and it["LR"].shape == (3, 16, 16) and it["SR"].shape == (3, 64, 64)
    assert it["HR"].dtype == torch.float32 and -1.0 <= it["HR"].min() and it["HR"].max() <= 1.0
    u8 = ds.get_u8(2)
    # ToTensor then *2-1 (data/util.py:66-75)
    assert torch.equal(it["HR"], (torch.from_numpy(u8["HR"]).permute(2, 0, 1).float() / 255.0) * 2 - 1)
    seen = []
    for b in D.BatchLoader(ds, 2, indices=range(1, 5), pin=False):
        assert b["HR"].dtype == torch.uint8 and b["HR"].shape[1:] == (64, 64, 3) and b["LR"].shape[1:] == (16, 16, 3)
        seen += b["Index"]
    assert seen == [1, 2, 3, 4]
    assert len(D.LRHRDataset(root, "img", 16, 64, data_len=3)) == 3
    with pytest.raises(NotImplementedError):
        D.LRHRDataset(root, "hdf5", 16, 64)
    # sr_ folder is optional (the library rebuilds it from lr_), but one of the two must exist
    root2 = _make_dataset(str(tmp_path / "ds2"), n=2, with_sr=False)
    assert "SR" not in D.LRHRDataset(root2, "img", 16, 64).get_u8(0)
    with pytest.raises(AssertionError):
        D.LRHRDataset(str(tmp_path), "img", 16, 64)


def test_lmdb_dataset_layout(tmp_path):
    """datatype 'lmdb' (data/LRHR_dataset.py:17-27, 60-93): the key scheme written by prepare_data_mfe_dm.py —
    `hr_{r}_{idx:05d}`, `sr_{l}_{r}_{idx:05d}`, `lr_{l}_{idx:05d}`, `length` — read through a key-value store.  The
    `lmdb` package is not in this image, so the store is a dict here; with the package the same code reads a real
    environment (and the real round trip is tested when it is importable)."""
    from io import BytesIO
    from PIL import Image
    rng = np.random.default_rng(0)

    def png(a):
        buf = BytesIO()
        Image.fromarray(a).save(buf, format="PNG")
        return buf.getvalue()

    imgs = [{"hr": rng.integers(0, 256, (64, 64, 3), dtype=np.uint8), "lr": rng.integers(0, 256, (16, 16, 3), dtype=np.uint8),
             "sr": rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)} for _ in range(3)]
    kv = {b"length": b"3"}
    for i, im in enumerate(imgs):
        kv['hr_64_{}'.format(str(i).zfill(5)).encode()] = png(im["hr"])
        kv['sr_16_64_{}'.format(str(i).zfill(5)).encode()] = png(im["sr"])
        kv['lr_16_{}'.format(str(i).zfill(5)).encode()] = png(im["lr"])
    ds = D.LRHRDataset("unused", "lmdb", 16, 64, kv=kv)
    assert len(ds) == 3
    u8 = ds.get_u8(1)
    assert np.array_equal(u8["HR"], imgs[1]["hr"]) and np.array_equal(u8["SR"], imgs[1]["sr"]) and np.array_equal(u8["LR"], imgs[1]["lr"])
    assert ds[2]["HR"].shape == (3, 64, 64) and ds[2]["LR"].shape == (3, 16, 16)
    batches = list(D.BatchLoader(ds, 2, pin=False))
    assert [b["Index"] for b in batches] == [[0, 1], [2]] and batches[0]["HR"].shape == (2, 64, 64, 3)
    del kv[b"hr_64_00001"]
    with pytest.raises(KeyError):
        ds.get_u8(1)
    assert len(D.LRHRDataset("unused", "lmdb", 16, 64, data_len=2, kv=kv)) == 2
    try:
        import lmdb
    except ImportError:
        with pytest.raises(ImportError, match="lmdb"):
            D.LRHRDataset(str(tmp_path), "lmdb", 16, 64)
    else:
        env = lmdb.open(str(tmp_path / "db"), map_size=1 << 26)
        with env.begin(write=True) as txn:
            for k, v in kv.items():
                txn.put(k, v)
        env.close()
        real = D.LRHRDataset(str(tmp_path / "db"), "lmdb", 16, 64)
        assert np.array_equal(real.get_u8(0)["HR"], imgs[0]["hr"])


def test_entry_scripts_refuse_training():
    from fastdiffsr_b200.evaluate import main
    with pytest.raises(NotImplementedError):
        main(["-p", "train"])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def net():
    opt = F.config.default_config()
    torch.manual_seed(0)
    netG = F.define_G(opt).to("cuda")
    netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
    return netG.eval()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 64, 64), (2, 37, 53), (1, 7, 9), (1, 256, 256)])
def test_device_metrics_vs_oracle(net, oracle, shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(H * W)
    hr = torch.nn.functional.avg_pool2d(torch.rand(B, 3, H + 2, W + 2, generator=g), 3, 1) * 2.4 - 1.2   # exercises the clamp
    sr = (hr + 0.15 * torch.randn(B, 3, H, W, generator=g))
    m = net.engine().metrics_u8(sr.cuda(), hr.cuda(), scale=4).cpu().numpy()
    for b in range(B):
        a8, b8 = oracle.to_u8(sr[b]), oracle.to_u8(hr[b])
        assert m[b, 0] == oracle.mse_u8(a8, b8)                                   # integer sums: exact
        assert m[b, 1] == pytest.approx(oracle.psnr_u8(a8, b8), rel=1e-13)
        assert m[b, 2] == pytest.approx(oracle.ssim_u8(a8, b8), abs=1e-9)
        assert m[b, 3] == pytest.approx(oracle.ergas_u8(a8, b8, 4), rel=1e-13)
    same = net.engine().metrics_u8(hr.cuda(), hr.cuda()).cpu().numpy()
    assert np.all(same[:, 0] == 0) and np.all(np.isinf(same[:, 1])) and np.allclose(same[:, 2], 1.0, atol=1e-12)


@pytest.mark.gpu
def test_batched_evaluation_loop(net, oracle, tmp_path):
    from fastdiffsr_b200.evaluate import evaluate
    root = _make_dataset(str(tmp_path / "ds"), n=5, l=16, r=64, with_sr=False)
    ds = D.LRHRDataset(root, "img", 16, 64)
    out_dir = str(tmp_path / "results")
    res = evaluate(net, ds, batch_size=2, scale=4, result_path=out_dir, seed=3)
    assert res["n"] == 5 and all(np.isfinite(res[k]) for k in res if k != "n")
    # the bicubic baseline is deterministic: recompute it on the host from the files with the oracle
    bic = {k: 0.0 for k in ("mse", "psnr", "ssim", "ergas")}
    for i in range(5):
        it = ds.get_u8(i)
        up = oracle.pil_bicubic_u8(it["LR"], 64, 64)
        bic["mse"] += oracle.mse_u8(up, it["HR"])
        bic["psnr"] += oracle.psnr_u8(up, it["HR"])
        bic["ssim"] += oracle.ssim_u8(up, it["HR"])
        bic["ergas"] += oracle.ergas_u8(up, it["HR"], 4)
    for k, v in bic.items():
        assert res["bic_" + k] == pytest.approx(v / 5, rel=1e-9), k
    from PIL import Image
    files = sorted(os.listdir(out_dir))
    assert files == [f"0_{i}_sr.tif" for i in range(1, 6)]
    assert np.asarray(Image.open(os.path.join(out_dir, files[0]))).shape == (64, 64, 3)
    # a precomputed sr_ folder (the reference's layout) gives the same conditioning, hence the same baseline
    root2 = _make_dataset(str(tmp_path / "ds2"), n=5, l=16, r=64, with_sr=True)
    res2 = evaluate(net, D.LRHRDataset(root2, "img", 16, 64), batch_size=5, scale=4, seed=3)
    for k in ("mse", "psnr", "ssim", "ergas"):
        assert res2["bic_" + k] == pytest.approx(res["bic_" + k], rel=1e-12)


@pytest.mark.gpu
def test_batched_evaluation_loop_sr3_baseline(oracle, tmp_path):
    """The same validation loop with which_model_G = 'ddpm' (SR3 baseline; short schedule, 5 images in batches
    of 2 so that the last call has a single image and returns the reference's batch-less (3,H,W) tensor)."""
    from fastdiffsr_b200.evaluate import evaluate
    opt = F.config.default_config("sr_ddpm_test_64_256")
    torch.manual_seed(0)
    netG = F.define_G(opt).to("cuda")
    netG.set_new_noise_schedule(dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=0.4), "cuda")
    root = _make_dataset(str(tmp_path / "ds"), n=5, l=16, r=64, with_sr=False)
    ds = D.LRHRDataset(root, "img", 16, 64)
    res = evaluate(netG.eval(), ds, batch_size=2, scale=4, seed=3)
    assert res["n"] == 5 and all(np.isfinite(res[k]) for k in res if k != "n")
    res_b = evaluate(netG, ds, batch_size=5, scale=4, seed=3)
    for k in ("mse", "psnr", "ssim", "ergas"):
        assert res_b["bic_" + k] == pytest.approx(res["bic_" + k], rel=1e-12)
