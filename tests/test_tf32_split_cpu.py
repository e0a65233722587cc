"""The operand split behind the sweep's tensor-core kernels (csrc/sweep.cu, "3xTF32"), restated in numpy: what the kernels
store in shared memory and what the tf32 datapath makes of it.  No GPU: this pins the arithmetic claim - hi + lo == x exactly,
and hi*hi + hi*lo + lo*hi with tf32-truncated terms and fp32 accumulation stays at fp32 level - that the GPU tests then
observe end to end (tests/test_sweep_tc_gpu.py)."""
import numpy as np

MASK = np.uint32(0xFFFFE000)  # sign, exponent and the 10 mantissa bits tf32 keeps


def tf32_trunc(x):
    return (x.view(np.uint32) & MASK).view(np.float32)


def split(x):
    hi = tf32_trunc(x)
    lo = (x - hi).astype(np.float32)
    return hi, lo


def test_split_is_exact_and_lo_is_small():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1 << 16) * np.exp(rng.uniform(-20, 20, 1 << 16))).astype(np.float32)
    hi, lo = split(x)
    assert np.array_equal((hi.astype(np.float64) + lo.astype(np.float64)).astype(np.float32), x)  # hi + lo == x, exactly
    assert np.all(np.abs(lo) <= np.abs(x) * 2.0 ** -10)          # lo holds the 13 bits tf32 drops
    assert np.array_equal(tf32_trunc(hi), hi)                    # hi is what the datapath reads
    # zeros, signed zeros and tiny values survive
    z = np.array([0.0, -0.0, 1e-38, -1e-38], dtype=np.float32)
    hz, lz = split(z)
    assert np.array_equal(hz + lz, z)


def test_three_term_product_stays_at_fp32_level():
    rng = np.random.default_rng(1)
    n, k = 64, 512
    a = rng.standard_normal((n, k)).astype(np.float32)
    b = rng.standard_normal((n, k)).astype(np.float32)
    ref = (a.astype(np.float64) * b.astype(np.float64)).sum(1)
    scale = (np.abs(a.astype(np.float64)) * np.abs(b.astype(np.float64))).sum(1)
    ah, al = split(a)
    bh, bl = split(b)
    al_t, bl_t = tf32_trunc(al), tf32_trunc(bl)  # the datapath reads the lo terms through tf32 as well
    one = (ah.astype(np.float64) * bh.astype(np.float64)).sum(1)
    three = (ah.astype(np.float64) * bh + ah.astype(np.float64) * bl_t + al_t.astype(np.float64) * bh).sum(1)
    fp32 = np.zeros(n, dtype=np.float32)
    for j in range(k):  # the FFMA kernels' sequential sum
        fp32 = (fp32 + a[:, j] * b[:, j]).astype(np.float32)
    err1 = np.abs(one - ref) / scale
    err3 = np.abs(three - ref) / scale
    errf = np.abs(fp32.astype(np.float64) - ref) / scale
    assert err1.max() > 1e-5                      # one tf32 term per operand: not acceptable for the exact path
    assert err3.max() < 2.0 ** -20                # three terms: the dropped lo*lo and the lo terms' truncation
    assert err3.max() < 4 * max(errf.max(), 2.0 ** -24) or err3.max() < 1e-6  # the same league as an fp32 FFMA chain
