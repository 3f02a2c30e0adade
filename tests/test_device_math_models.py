"""CPU models of device-side arithmetic shortcuts (lumillyrender_b200/csrc/device_path.cuh), checked against the
operation they replace with the same IEEE fp32 operations in numpy.  The device code itself is exercised by the
GPU parity tests; these pin the ALGORITHMS bit for bit on inputs the renders may never hit."""
import numpy as np


def fmod_pos_model(x, m):
    """device_path.cuh: fmod_pos — q = trunc(fl(x / m)); r = fma(-q, m, x); r < 0 ? r + m : r   (x >= 0, m > 0)"""
    x = x.astype(np.float32)
    m = np.float32(m)
    q = np.trunc((x / m).astype(np.float32)).astype(np.float32)
    # fmaf: one rounding of the exact x - q*m; exact in float64 here (q*m < 2^48, the difference needs <= 25 bits)
    r = (x.astype(np.float64) - q.astype(np.float64) * np.float64(m)).astype(np.float32)
    return np.where(r < 0, (r + m).astype(np.float32), r)


def test_fast_fmod_is_exact():
    rng = np.random.RandomState(5)
    for m in (30.0, 150.0, 300.0, 7.25, 0.1):
        mf = np.float32(m)
        xs = [rng.uniform(0, 1e4, 400000), rng.uniform(0, 1e9, 400000), np.abs(rng.normal(0, 1, 100000)) * 1e-3,
              10.0 ** rng.uniform(-30, 9, 200000)]
        # adversarial: multiples of m and their fp32 neighbours (the quotient rounds to an integer from either side)
        k = rng.randint(0, 1 << 22, 300000).astype(np.float64)
        mult = (k * np.float64(mf)).astype(np.float32)
        xs += [mult, np.nextafter(mult, np.float32(0)), np.nextafter(mult, np.float32(np.inf)), np.array([0.0, m, 2 * m, 1e9 - 64])]
        x = np.concatenate(xs).astype(np.float32)
        x = x[(x >= 0) & (x < mf * np.float32(8388608.0))]              # beyond: the device calls fmodf itself
        want = np.fmod(x.astype(np.float64), np.float64(mf)).astype(np.float32)     # exact: fmod is an exact operation
        got = fmod_pos_model(x, mf)
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (m, x[bad[:5]], got[bad[:5]], want[bad[:5]])




def test_fmod_by_one_is_x_minus_trunc_exactly():
    """sky.rs:60-61 takes `% 1.0` of the equirect coordinates; the device computes x - trunc(x) (device_path.cuh: sky_ibl).
    Same IEEE operations in numpy: equal bit for bit to fmod on random, tiny, huge, negative and special inputs (up to the
    sign of a zero result, which indexes the same texel)."""
    rng = np.random.RandomState(3)
    x = np.concatenate([
        rng.uniform(-4, 4, 2_000_000), rng.uniform(-1e-3, 1e-3, 200_000), rng.uniform(-2e7, 2e7, 200_000),
        np.ldexp(rng.uniform(0.5, 1, 200_000), rng.randint(-40, 40, 200_000)) * rng.choice([-1, 1], 200_000),
        np.array([0.0, -0.0, 1.0, -1.0, 0.99999994, 1.0000001, 8388608.0, -8388608.0, 16777216.0, 3.4e38, np.inf, -np.inf, np.nan]),
    ]).astype(np.float32)
    with np.errstate(invalid="ignore"):
        want = np.fmod(x, np.float32(1.0))
        got = x - np.trunc(x)
    assert got.dtype == np.float32
    both_nan = np.isnan(want) & np.isnan(got)
    assert np.array_equal(want[~both_nan], got[~both_nan])          # -0.0 == +0.0 here: the texel index floor(W * u) is 0 for both
    assert both_nan.sum() == 3
