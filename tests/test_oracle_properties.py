"""Property tests (hypothesis) of the oracle against the independent numpy restatement: random scenes with the
awkward values the reference's arithmetic can meet (zeros, denormals, huge / tiny ranges, ties, cutout faces)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from np_restatement import np_pack, np_sort

finite = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False, width=32)
tricky = st.sampled_from([0.0, -0.0, 1e-30, -1e-30, 1e-45, -1e-45, 1.0, -1.0, 3.4e38, -3.4e38, 1e-6, -1e-6, 65535.0, 0.5, -0.5])
value = st.one_of(finite, tricky)


@settings(max_examples=150, deadline=None)
@given(n=st.integers(0, 200), seed=st.integers(0, 2 ** 31 - 1), view=st.lists(value, min_size=4, max_size=4),
       mode=st.sampled_from(["random", "ties", "narrow", "wide"]), use_cutout=st.booleans())
def test_sort_matches_numpy_property(orc, n, seed, view, mode, use_cutout):
    rng = np.random.default_rng(seed)
    m = np.zeros((n, 16), np.float32)
    if mode == "random":
        m[:, 12:15] = rng.normal(0, 5, (n, 3))
    elif mode == "ties":
        m[:, 12:15] = rng.integers(-2, 3, (n, 3))
    elif mode == "narrow":  # depth range far below f32 resolution: keys leave [0, 65535] (quirk Q5)
        m[:, 12:15] = 1000.0 + rng.normal(0, 1e-5, (n, 3))
    else:
        m[:, 12:15] = rng.normal(0, 1, (n, 3)) * 10.0 ** rng.integers(-20, 20, (n, 1))
    m[:, 15] = np.abs(rng.normal(0, 1, n)) * rng.integers(0, 2, n)
    v = np.array(view, np.float32)
    cut = None
    if use_cutout:
        cut = np.eye(4, dtype=np.float32)
        cut[:3, :3] *= rng.uniform(0.05, 2.0, 3).astype(np.float32)
        cut[3, :3] = rng.normal(0, 1, 3)  # column-major: translation in elements 12..14
        cut = cut.reshape(16)
    with np.errstate(all="ignore"):
        exp = np_sort(m, v, cut)
    got = orc.sort(m, v, cut)
    assert got.dtype == np.uint32 and np.array_equal(got, exp)


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 300), scale_exp=st.integers(-30, 3))
def test_pack_matches_numpy_property(orc, seed, n, scale_exp):
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 32), np.uint8)
    f = np.zeros((n, 6), np.float32)
    f[:, :3] = rng.normal(0, 3, (n, 3))
    f[:, 3:] = np.abs(rng.normal(0, 1, (n, 3))) * 10.0 ** scale_exp
    rows[:, :24] = f.view(np.uint8).reshape(n, 24)
    rows[:, 24:] = rng.integers(0, 256, (n, 8), dtype=np.uint8)
    cs, cc, m = orc.pack(rows)
    with np.errstate(all="ignore"):
        ncs, ncc, nsa, tiny = np_pack(rows)
    assert np.array_equal(cs.view(np.uint32), ncs.view(np.uint32))
    assert np.array_equal(m[:, 15].view(np.uint32), nsa.view(np.uint32))
    ok = ~tiny
    assert np.array_equal(cc[ok], ncc[ok])
    # where the parseInt exponent-form quirk applies the stored value is a single leading digit
    a = cc[tiny][:, :3].copy().view(np.int16)
    b = ncc[tiny][:, :3].copy().view(np.int16)
    assert np.all(np.abs(a[a != b]) <= 9)
