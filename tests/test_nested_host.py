"""CPU: the nested-sampling driver (isochrones_b200/nested.py) on analytic problems.

The driver only needs an object with ``n_params``, ``param_names`` and ``mnest_lnpost_batch(cube[n, ndim]) -> lnpost``
(in place: cube -> parameters), which on a GPU box is BasicStarModel (one launch per batch).  Here a numpy stand-in with
known evidence exercises the host logic: volume bookkeeping, the multi-ellipsoid bound (MultiNest's decomposition) on a
bimodal and on a curved posterior, the finite-fraction correction, determinism."""
import numpy as np

from isochrones_b200.nested import _bounding_ellipsoids, _draw_from_union, nested_sample


class _Analytic(object):
    def __init__(self, fn, ndim):
        self.fn, self.n_params, self.param_names = fn, ndim, ["p%d" % i for i in range(ndim)]
        self.calls = 0

    def mnest_lnpost_batch(self, work):
        self.calls += 1
        return self.fn(work)


def _two_gaussians(u, s=0.02):
    n = u.shape[1]
    norm = -0.5 * n * np.log(2 * np.pi * s * s)
    a = norm - 0.5 * np.sum((u - 0.3) ** 2, axis=1) / s ** 2
    b = norm - 0.5 * np.sum((u - 0.75) ** 2, axis=1) / s ** 2
    return np.logaddexp(a, b)                       # integral over the cube: 2


def _banana(u):
    x = (u - 0.5) * 6.0
    return (-0.5 * (x[:, 0] / 0.5) ** 2 - 0.5 * ((x[:, 1] - x[:, 0] ** 2 + 1.0) / 0.1) ** 2
            - 0.5 * np.sum(x[:, 2:] ** 2, axis=1) / 0.3 ** 2)


_BANANA_LOGZ = np.log((2 * np.pi) ** 2 * 0.5 * 0.1 * 0.3 * 0.3 / 6.0 ** 4)


def test_two_modes_evidence_and_cost():
    single = nested_sample(_Analytic(_two_gaussians, 5), n_live=800, seed=2, multi=False)
    multi = nested_sample(_Analytic(_two_gaussians, 5), n_live=800, seed=2, multi=True)
    for r in (single, multi):
        assert r.converged and abs(r.logZ - np.log(2.0)) < 4.0 * r.logZ_err, (r.logZ, r.logZ_err)
        assert abs(r.weights.sum() - 1.0) < 1e-12
    assert multi.n_ellipsoids_max >= 2 and single.n_ellipsoids_max == 1
    assert multi.n_evals < 0.5 * single.n_evals, (multi.n_evals, single.n_evals)
    # both modes carry half of the posterior mass
    left = multi.weights[multi.samples[:, 0] < 0.5].sum()
    assert abs(left - 0.5) < 0.1, left


def test_curved_degeneracy_evidence():
    r = nested_sample(_Analytic(_banana, 4), n_live=800, seed=5)
    assert r.converged and r.n_ellipsoids_max > 3
    assert abs(r.logZ - _BANANA_LOGZ) < 4.0 * r.logZ_err, (r.logZ, _BANANA_LOGZ, r.logZ_err)
    assert r.efficiency > 0.05
    again = nested_sample(_Analytic(_banana, 4), n_live=800, seed=5)
    assert again.logZ == r.logZ and again.n_evals == r.n_evals


def test_non_finite_region_enters_through_its_measured_fraction():
    def holed(u):                                   # a Gaussian on the half of the cube where lnpost is finite at all
        lp = _two_gaussians(u)
        return np.where(u[:, 1] < 0.5, lp, np.nan)  # NaN counts as -inf (nested.py)

    r = nested_sample(_Analytic(holed, 3), n_live=600, seed=7)
    assert 0.4 < r.finite_fraction < 0.6
    assert abs(r.logZ - 0.0) < 4.0 * r.logZ_err + 0.05, r.logZ      # only the mode at 0.3 survives: integral 1


def test_union_of_ellipsoids_is_sampled_uniformly():
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.normal(0.25, 0.02, (300, 3)), rng.normal(0.7, 0.03, (300, 3))])
    ells = _bounding_ellipsoids(pts, 1.2, np.log(1e-9))
    assert len(ells) == 2
    x = _draw_from_union(ells, 20000, rng)
    share = np.mean(x[:, 0] < 0.5)
    want = np.exp(ells[0].logvol) / (np.exp(ells[0].logvol) + np.exp(ells[1].logvol))
    want = want if ells[0].mu[0] < 0.5 else 1.0 - want
    assert abs(share - want) < 0.02, (share, want)
    assert all(e.contains(pts[:300] if e.mu[0] < 0.5 else pts[300:]).all() for e in ells)


def test_live_points_only_and_iteration_cap():
    full = nested_sample(_Analytic(_two_gaussians, 3), n_live=400, seed=11)
    lean = nested_sample(_Analytic(_two_gaussians, 3), n_live=400, seed=11, return_dead=False)
    assert lean.logZ == full.logZ and lean.n_evals == full.n_evals          # same run, only the bookkeeping differs
    assert lean.samples.shape == (400, 3) and abs(lean.weights.sum() - 1.0) < 1e-12
    capped = nested_sample(_Analytic(_two_gaussians, 3), n_live=400, seed=11, max_iter=500)
    assert not capped.converged and capped.n_iter == 500 and capped.logZ < full.logZ


def test_everything_non_finite_raises():
    import pytest

    with pytest.raises(RuntimeError):
        nested_sample(_Analytic(lambda u: np.full(len(u), np.nan), 2), n_live=50, seed=0, max_init_draws=100_000)
