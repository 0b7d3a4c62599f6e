"""Size-independent properties of the host-side logic and of the oracle's integer paths (CPU, hypothesis)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from fastdiffsr_b200 import make_beta_schedule
from fastdiffsr_b200.parallel import psnr_from_sse, shard_batch, shard_bounds


@given(n=st.integers(0, 300), world=st.integers(1, 16))
def test_shard_bounds_partition_the_batch(n, world):
    covered = []
    per0 = None
    for r in range(world):
        a, b, per = shard_bounds(n, r, world)
        per0 = per if per0 is None else per0
        assert per == per0 and 0 <= a <= b <= n and b - a <= per
        covered += list(range(a, b))
    assert covered == list(range(n))               # every image exactly once, in order
    assert per0 * world >= n


@given(n=st.integers(1, 40), world=st.integers(1, 8))
def test_shard_batch_pads_to_equal_shapes(n, world):
    x = torch.arange(n * 2, dtype=torch.float32).view(n, 2)
    parts, valid = zip(*[shard_batch(x, r, world) for r in range(world)])
    assert len({tuple(p.shape) for p in parts}) == 1
    assert sum(valid) == n
    back = torch.cat([p[:v] for p, v in zip(parts, valid)], 0)
    assert torch.equal(back, x)


@given(T=st.integers(2, 200), name=st.sampled_from(["linear", "quad", "const", "jsd", "warmup10", "warmup50",
                                                    "linear_cosine", "cosine"]))
@settings(max_examples=60, deadline=None)
def test_beta_schedules_are_valid_probabilities(T, name):
    b = np.asarray(make_beta_schedule(name, T, 1e-4, 2e-2), dtype=np.float64)
    assert b.shape == (T,) and np.all(b > 0) and np.all(b <= 1.0)
    ac = np.cumprod(1.0 - b)
    assert np.all(np.diff(ac) <= 0)                 # signal level never increases along the chain


@given(h=st.integers(1, 24), scale=st.sampled_from([2, 3, 4, 8]))
@settings(max_examples=40, deadline=None)
def test_bicubic_taps_are_normalised_and_constant_images_stay_constant(oracle, h, scale):
    bounds, taps = oracle.bicubic_coeffs(h, h * scale)
    for i, (xmin, n) in enumerate(bounds):
        assert 0 <= xmin and xmin + n <= h and n >= 1
        assert abs(int(taps[i, :n].sum()) - (1 << 22)) <= n       # 22-bit fixed point, rounding of each tap
    v = int(np.random.default_rng(h).integers(0, 256))
    img = np.full((h, h, 3), v, dtype=np.uint8)
    out = oracle.pil_bicubic_u8(img, h * scale, h * scale)
    assert out.shape == (h * scale, h * scale, 3) and np.all(out == v)


@given(sse=st.lists(st.floats(1.0, 1e12), min_size=2, max_size=8))
def test_psnr_decreases_with_error(sse):
    s = torch.tensor(sorted(sse), dtype=torch.float64)
    p = psnr_from_sse(s, 3 * 64 * 64)
    assert torch.all(p[:-1] >= p[1:])
    assert torch.isinf(psnr_from_sse(torch.zeros(1, dtype=torch.float64), 10)).all()
