"""Host model of block_hyst_min_scan (csrc/select.cuh): the sequential "running best with hysteresis" scan of
PrimalSimplexSolver2.cs:102-141 / DualSimplex.cs:27-70 (accept k iff val_k < best - eps) against its prefix-minimum
formulation -- sure acceptances, sure rejections, and a literal replay of the span of "unsure" candidates behind the
last sure acceptance -- on value patterns chosen to hit the unsure case (near-ties of the minimum, slowly falling
staircases, exact ties).  The device code is checked against the oracle in the GPU tests; this pins the argument."""
import numpy as np
import pytest


def sequential(vals, ok, b0, eps, lo=0, hi=None, best=None, idx=-1):
    best = b0 if best is None else best
    hi = len(vals) - 1 if hi is None else hi
    for k in range(lo, hi + 1):
        if ok[k] and vals[k] < best - eps:
            best, idx = vals[k], k
    return idx


def prefix_min_form(vals, ok, b0, eps, scan_threads=8):
    n = len(vals)
    per = ((n + scan_threads - 1) // scan_threads) | 1
    q0 = b0 - eps
    local_min = []
    for t in range(scan_threads):
        lm = np.inf
        for k in range(t * per, min(n, t * per + per)):
            if ok[k]:
                lm = min(lm, vals[k])
        local_min.append(lm)
    kstar, ulo, uhi = -1, 10 ** 9, -1
    for t in range(scan_threads):
        m = min([np.inf] + local_min[:t])  # minimum of the valid values before this thread's run
        for k in range(t * per, min(n, t * per + per)):
            if not ok[k]:
                continue
            v = vals[k]
            if v < min(b0, m) - eps:        # accepted whatever happened before
                kstar = max(kstar, k)
            elif v < min(q0, m):            # a new prefix minimum by less than eps: depends on the history
                ulo, uhi = min(ulo, k), max(uhi, k)
            m = min(m, v)
    idx = kstar
    best = vals[idx] if idx >= 0 else b0
    lo = max(ulo, idx + 1)
    if lo <= uhi:
        idx = sequential(vals, ok, b0, eps, lo, uhi, best, idx)
    return idx


@pytest.mark.parametrize("kind", ["integers", "integers_with_noise", "noise_around_zero", "falling_staircase"])
def test_prefix_minimum_form_equals_the_sequential_scan(kind):
    rng = np.random.default_rng({"integers": 1, "integers_with_noise": 2, "noise_around_zero": 3, "falling_staircase": 4}[kind])
    for _ in range(4000):
        n = int(rng.integers(1, 150))
        if kind == "integers":
            vals = rng.integers(-5, 5, n).astype(float)
        elif kind == "integers_with_noise":
            vals = rng.integers(-5, 5, n) + rng.normal(0, 1e-9, n)
        elif kind == "noise_around_zero":
            vals = rng.normal(0, 1e-9, n) * rng.integers(0, 3, n)
        else:
            vals = np.cumsum(-np.abs(rng.normal(0, 0.7e-9, n))) + rng.integers(0, 2, n) * 1e-9
        ok = rng.random(n) < 0.8
        eps = (1e-9, 1e-10)[int(rng.integers(0, 2))]
        b0 = (0.0, np.inf)[int(rng.integers(0, 2))]
        threads = int(rng.integers(1, 12))
        assert prefix_min_form(vals, ok, b0, eps, threads) == sequential(vals, ok, b0, eps)
