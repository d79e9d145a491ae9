"""The rounding bound behind the fast assignment kernel's exactness guard ('Exactness guard'),
checked on the CPU by emulating the kernel's fp32 arithmetic in numpy:

    d_hat = fl32(x - fl32(c')),   acc <- fl32(d_hat * d_hat + acc)   (one fused multiply-add per entry)
    claim:  |acc - sum (x - c')^2|  <=  E(acc) = 1.01 (m+5) u acc + 2.02 u cmax sqrt(m acc) + 2.1 u^2 m cmax^2

with u = 2^-24, c' the fp64 scaled centre, cmax = max |fl32(c')|.  A column is certified only if its best and
second-best sums are further apart than E(best) + E(second), so a violated bound would be a parity bug."""
import numpy as np
import pytest

U = 2.0 ** -24


def _fp32_sums(x32, c64):
    """acc per centre, emulating fmaf(d, d, acc) (the product of two floats is exact in double)."""
    c32 = c64.astype(np.float32)
    acc = np.zeros(c64.shape[1], dtype=np.float32)
    for t in range(x32.shape[0]):
        d = (x32[t] - c32[t]).astype(np.float32)                       # one rounding
        acc = (d.astype(np.float64) * d.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    return acc, c32


def _bound(acc, m, cmax):
    acc = acc.astype(np.float64)
    return 1.01 * (m + 5) * U * acc + 2.02 * U * cmax * np.sqrt(m * acc) + 2.1 * U * U * m * cmax ** 2


@pytest.mark.parametrize("m,scale_x,scale_c", [(3, 1.0, 1.0), (26, 20.0, 20.0), (78, 10.0, 10.0), (78, 1e-3, 1e3),
                                               (51, 1e4, 1e-2), (254, 3.0, 5.0), (78, 1.0, 0.0)])
def test_fp32_sum_error_is_inside_the_guard(m, scale_x, scale_c):
    rng = np.random.default_rng(m)
    worst = 0.0
    for trial in range(200):
        K = 6
        x32 = (scale_x * rng.standard_normal(m)).astype(np.float32)
        c64 = scale_c * rng.standard_normal((m, K))
        if trial % 4 == 0:                                             # near-cancellation: points sitting on a centre
            c64[:, 0] = x32.astype(np.float64) * (1 + 1e-7 * rng.standard_normal(m))
        acc, c32 = _fp32_sums(x32, c64)
        exact = ((x32.astype(np.float64)[:, None] - c64) ** 2).sum(axis=0)
        cmax = float(np.max(np.abs(c32))) if c32.size else 0.0
        E = _bound(acc, m, cmax)
        err = np.abs(acc.astype(np.float64) - exact)
        assert np.all(err <= E + 1e-300), (trial, err.max(), E[np.argmax(err - E)])
        worst = max(worst, float(np.max(err / np.maximum(E, 1e-300))))
    assert worst <= 1.0


def test_certified_winner_is_the_exact_winner():
    """Whenever the guard certifies (gap > E(best) + E(second)), the fp32 argmin equals the fp64 argmin."""
    rng = np.random.default_rng(1)
    certified = 0
    for trial in range(400):
        m, K = 40, 8
        x32 = (10 * rng.standard_normal(m)).astype(np.float32)
        c64 = 10 * rng.standard_normal((m, K))
        c64[:, 1] = c64[:, 0] * (1 + (10.0 ** -rng.integers(3, 9)) * rng.standard_normal(m))   # a near twin
        acc, c32 = _fp32_sums(x32, c64)
        order = np.argsort(acc, kind="stable")
        b1, b2 = acc[order[0]], acc[order[1]]
        cmax = float(np.max(np.abs(c32)))
        if float(b2) - float(b1) > _bound(np.array([b1]), m, cmax)[0] + _bound(np.array([b2]), m, cmax)[0]:
            certified += 1
            exact = ((x32.astype(np.float64)[:, None] - c64) ** 2).sum(axis=0)
            assert int(np.argmin(exact)) == int(order[0])
    assert certified > 50


def test_masked_distance_moves_by_at_most_the_centre_shift():
    """The inequality the bounded assignment rests on (csrc/bounded.cu): the masked distance is a seminorm of the
    centre, so |d_j(c_new) - d_j(c_old)| <= ||c_new - c_old||_2 for every support; hence a lower bound on the
    distance to a centre stays valid after subtracting that centre's movement."""
    rng = np.random.default_rng(2)
    p = 200
    for _ in range(300):
        m = int(rng.integers(1, 60))
        rows = rng.choice(p, m, replace=False)
        x = rng.standard_normal(m) * 10
        c_old = rng.standard_normal(p) * 10
        c_new = c_old + rng.standard_normal(p) * 10.0 ** rng.integers(-6, 1)
        d_old = np.linalg.norm(x - c_old[rows])
        d_new = np.linalg.norm(x - c_new[rows])
        shift = np.linalg.norm(c_new - c_old) * (1 + 1e-9)
        assert abs(d_new - d_old) <= shift
        lb = d_old * (1 - 1e-9)                      # a valid lower bound before the move ...
        assert d_new >= (lb - shift) * (1 - 4.77e-7)  # ... lowered the way k_assign_bounded lowers it


# ---- the half-precision prefix of the pruned assignment pass (csrc/prefix16.cu) ----
def _prefix16_scale(cmax32):
    """s = 2^e with s * cmax in [2^13, 2^14), e clamped to [-100, 100] (prefix16_scale in prefix16.cu)"""
    if not (cmax32 > 0) or not np.isfinite(cmax32):
        return np.float32(1.0)
    _, ex = np.frexp(np.float32(cmax32))
    return np.float32(2.0 ** int(np.clip(14 - ex, -100, 100)))


def _prefix16_lower_bounds(x32, c64, q):
    """Emulates k_prefix16 on the first q entries: t = fp16(fl32(s c')), A = fp32 sum of fl32(s x - t)^2 by FMAs, and the
    kernel's bound expression evaluated in fp32.  Returns (candidate, lb, partial sums)."""
    c32 = c64.astype(np.float32)
    cmax = np.float32(np.max(np.abs(c32))) if c32.size else np.float32(0)
    s = _prefix16_scale(cmax)
    t = (c64[:q] * np.float64(s)).astype(np.float32).astype(np.float16)               # fp64 -> fp32 -> fp16, as the table kernel
    delta = np.float32(np.float32(4.888e-4) * np.float32(s * cmax) + np.float32(6.0e-8))
    acc = np.zeros(c64.shape[1], dtype=np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        for e in range(q):
            xs = np.float32(x32[e] * s)                                                # exact: s is a power of two
            d = (t[e].astype(np.float32) - xs).astype(np.float32)                      # sub.rn.f32.f16: one rounding
            acc = (d.astype(np.float64) * d.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    keys = (acc.view(np.uint32) & np.uint32(0xFFFFFFC0)) | np.arange(acc.size, dtype=np.uint32)   # index in the low mantissa bits
    order = np.argsort(keys, kind="stable")
    cand = int(keys[order[0]] & 63)
    v2 = np.uint32(keys[order[1]] & np.uint32(0xFFFFFFC0))
    lb = np.float32(0)
    if v2 < np.uint32(0x7F800000):
        b2 = np.array([v2], dtype=np.uint32).view(np.float32)[0]
        ga = np.float32(1.01 * (q + 5) * U)
        sq = np.float32(np.sqrt(q) * (1 + 1e-6))
        lbs = np.float32(np.float32(np.sqrt(np.float32(b2 * np.float32(1 - ga)))) * np.float32(1 - 2.4e-7)) - np.float32(sq * delta)
        if lbs > 0:
            lb = np.float32(np.float32(lbs * np.float32(1.0 / s)) * np.float32(1 - 4.8e-7))
    return cand, float(lb), acc


@pytest.mark.parametrize("scale_x,scale_c", [(1.0, 1.0), (20.0, 20.0), (1e-3, 1e3), (1e4, 1e-2), (3e-20, 5e-20), (1e12, 1e12),
                                             (1.0, 0.0), (1e-30, 1e-30)])
def test_half_table_prefix_bound_is_a_lower_bound(scale_x, scale_c):
    """lb from the fp16-table prefix never exceeds the reference's distance from the column to ANY centre but the candidate
    (the full distance is at least the prefix distance: every term is a square)."""
    rng = np.random.default_rng(int(abs(np.log10(scale_x + 1e-300)) * 7 + abs(np.log10(scale_c + 1e-300))))
    tightest = 0.0
    for trial in range(300):
        m, K, q = 52, 40, 8
        x32 = (scale_x * rng.standard_normal(m)).astype(np.float32)
        c64 = scale_c * rng.standard_normal((m, K))
        if trial % 3 == 0:                                             # clustered: the column sits near centre 0, a twin of it nearby
            c64[:, 0] = x32.astype(np.float64) * (1 + 1e-3 * rng.standard_normal(m))
            c64[:, 1] = c64[:, 0] * (1 + 10.0 ** -rng.integers(2, 6) * rng.standard_normal(m))
        cand, lb, _ = _prefix16_lower_bounds(x32, c64, q)
        full = np.sqrt(((x32.astype(np.float64)[:, None] - c64) ** 2).sum(axis=0))        # the reference's distances (fp64)
        pref = np.sqrt(((x32.astype(np.float64)[:q, None] - c64[:q]) ** 2).sum(axis=0))
        others = np.delete(np.arange(K), cand)
        assert lb <= pref[others].min() * (1 + 1e-12) + 1e-300, (trial, lb, pref[others].min())
        assert lb <= full[others].min() * (1 + 1e-12) + 1e-300
        if pref[others].min() > 0:
            tightest = max(tightest, lb / pref[others].min())
    if scale_c > 0 and 1e-25 < scale_x < 1e13:
        assert tightest > 0.97, f"the bound should also be useful, best ratio {tightest}"   # fp16 costs ~1e-3 of the distance
