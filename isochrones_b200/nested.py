"""Nested sampling over the unit hypercube with every likelihood evaluation batched on the GPU — the B200 form of the
reference's DEFAULT fit path, ``pymultinest.run(mod.mnest_loglike, mod.mnest_prior, n)`` (``starmodel.py:717-802``).

MultiNest itself is third-party and un-vendored (``.ci/travis.sh:20``); what the reference defines is the problem it
hands over: live points live in the unit cube, ``mnest_prior`` maps a cube point to the parameters with the affine box
map of ``starmodel.py:1637-1640`` and the "likelihood" of a point is ``mnest_loglike`` = ``lnpost`` (``:1642-1645``: the
parameter priors are part of the integrand, the cube measure is uniform).  The quantity MultiNest reports is therefore

    Z = integral over the unit cube of exp(lnpost(theta(u))) du

and this module estimates the same Z, and the same weighted posterior sample, with the textbook algorithm (Skilling
2006) in the rejection-within-a-bounding-ellipsoid form MultiNest popularised (Feroz & Hobson 2008; a single ellipsoid
here, enlarged to hold every live point with a margin):

* the cube -> parameters -> lnpost chain of a whole batch of candidate points is ONE launch
  (``iso_mnest_lnpost_batch``); candidates are drawn uniformly inside the current bounding ellipsoid (intersected with
  the cube) thousands at a time and consumed in generation order, so the sequence of accepted points is exactly the one
  a one-at-a-time sampler drawing from the same stream would produce;
* points whose lnpost is not finite (outside the model / BC grids: most of the prior box) are excluded from the
  integral up front: the live set is initialised with finite points only and the evidence carries the measured
  fraction of the cube that is finite — a flat plateau of ``-inf`` live points would break the volume bookkeeping.

The host loop is O(log n_live) heap work per accepted point; all arithmetic that scales with the number of likelihood
evaluations runs on the device.
"""
import heapq

import numpy as np


class NestedResult(dict):
    """``logZ``, ``logZ_err``, ``information`` (nats), weighted ``samples[n, ndim]`` (parameters, not cube points) with
    ``weights`` (sum 1) and ``lnpost``, ``n_evals``, ``n_iter``, ``efficiency``, ``finite_fraction``."""

    __getattr__ = dict.__getitem__

    def equal_weighted(self, n=None, seed=0):
        """``n`` posterior draws with equal weights (systematic resampling of the weighted dead + live points)."""
        w = self["weights"]
        n = int(n or max(1, round(1.0 / np.sum(w ** 2))))          # default: the effective sample size
        rng = np.random.default_rng(seed)
        pos = (rng.random() + np.arange(n)) / n
        idx = np.searchsorted(np.cumsum(w), pos)
        return self["samples"][np.minimum(idx, len(w) - 1)]

    def mean(self):
        return self["weights"] @ self["samples"]

    def std(self):
        m = self.mean()
        return np.sqrt(self["weights"] @ (self["samples"] - m) ** 2)


def _logaddexp(a, b):
    return np.logaddexp(a, b)


class _Ellipsoid(object):
    """Bounding ellipsoid of a point set: the covariance ellipsoid scaled to hold every point, radius times ``enlarge``."""

    def __init__(self, pts, enlarge):
        self.mu = pts.mean(axis=0)
        d = pts - self.mu
        cov = d.T @ d / max(len(pts) - 1, 1)
        cov += np.eye(cov.shape[0]) * (1e-12 * max(np.trace(cov), 1e-300) + 1e-300)
        self.L = np.linalg.cholesky(cov)
        z = np.linalg.solve(self.L, d.T)                          # whitened points
        self.r = enlarge * np.sqrt(np.max(np.sum(z * z, axis=0)))

    def draw(self, n, rng):
        ndim = len(self.mu)
        z = rng.standard_normal((n, ndim))
        z *= (rng.random(n) ** (1.0 / ndim) / np.sqrt(np.sum(z * z, axis=1)))[:, None]      # uniform in the unit ball
        return self.mu + self.r * (z @ self.L.T)


def nested_sample(mod, n_live=1000, dlogz=0.5, seed=0, enlarge=1.25, batch=8192, max_iter=5_000_000, max_batches=20_000,
                  return_dead=True):
    """Nested sampling of ``mod`` (a ``BasicStarModel``): returns a :class:`NestedResult`.

    ``n_live``: live points (MultiNest's ``n_live_points``, starmodel.py:667-671 default 1000); ``dlogz``: stop when the
    live points can add at most this much to ln Z (MultiNest's ``evidence_tolerance`` 0.5); ``batch``: candidate points
    per launch; ``enlarge``: linear margin of the bounding ellipsoid."""
    rng = np.random.default_rng(seed)
    ndim = mod.n_params
    n_evals = 0

    def evaluate(u):
        """cube points ``u[n, ndim]`` -> (parameters, lnpost with NaN -> -inf); one launch"""
        nonlocal n_evals
        work = np.ascontiguousarray(u, dtype=np.float64).copy()
        lp = mod.mnest_lnpost_batch(work)
        n_evals += len(work)
        lp = np.where(np.isnan(lp), -np.inf, lp)
        return work, lp

    # ---- live set: finite points only; the finite fraction of the cube enters the evidence --------------------------
    live_u = np.empty((0, ndim))
    live_p = np.empty((0, ndim))
    live_l = np.empty(0)
    n_tried = 0
    while len(live_l) < n_live:
        u = rng.random((max(batch, 4 * n_live), ndim))
        p, lp = evaluate(u)
        ok = np.isfinite(lp)
        need = n_live - len(live_l)
        take = np.flatnonzero(ok)[:need]
        # only the draws up to the last one taken count towards the fraction (the rest of the batch was never "tried")
        n_tried += (take[-1] + 1) if len(take) == need else len(u)
        live_u = np.concatenate([live_u, u[take]])
        live_p = np.concatenate([live_p, p[take]])
        live_l = np.concatenate([live_l, lp[take]])
        if n_tried > 2_000_000_000:
            raise RuntimeError("no finite lnpost in the prior box")
    finite_fraction = n_live / float(n_tried)
    n_init = n_evals

    heap = [(live_l[i], i) for i in range(n_live)]
    heapq.heapify(heap)
    dead_p, dead_l, dead_logw = [], [], []
    logz = -np.inf
    log_x = 0.0                          # ln of the remaining prior volume (relative to the finite region)
    it, n_batches, n_acc = 0, 0, 0
    done = False
    log_shrink = -1.0 / n_live
    # ln(w_i) for X_i = exp(-i / n): w_i = X_{i-1} - X_i = X_{i-1} (1 - e^{-1/n})
    log_dw = np.log1p(-np.exp(log_shrink))
    while not done and n_batches < max_batches:
        ell = _Ellipsoid(live_u, enlarge)
        cand = ell.draw(batch, rng)
        cand = cand[np.all((cand >= 0.0) & (cand <= 1.0), axis=1)]
        n_batches += 1
        if len(cand) == 0:
            continue
        cp, cl = evaluate(cand)
        for j in range(len(cand)):
            lmin, imin = heap[0]
            if not (cl[j] > lmin):
                continue
            # the worst live point dies with weight w = X_{i-1} - X_i
            logw = log_x + log_dw
            if return_dead:
                dead_p.append(live_p[imin].copy())
            dead_l.append(lmin)
            dead_logw.append(logw)
            logz = _logaddexp(logz, lmin + logw)
            log_x += log_shrink
            live_u[imin], live_p[imin], live_l[imin] = cand[j], cp[j], cl[j]
            heapq.heapreplace(heap, (cl[j], imin))
            it += 1
            n_acc += 1
            if it % 64 == 0 or it >= max_iter:
                lmax = live_l.max()
                if it >= max_iter or _logaddexp(logz, lmax + log_x) - logz < dlogz:
                    done = True
                    break
    # ---- the live points share the remaining volume -----------------------------------------------------------------
    logw_live = log_x - np.log(n_live)
    all_l = np.concatenate([np.asarray(dead_l), live_l])
    all_logw = np.concatenate([np.asarray(dead_logw), np.full(n_live, logw_live)])
    all_p = np.concatenate([np.asarray(dead_p).reshape(-1, ndim), live_p]) if return_dead else live_p
    logp = all_l + all_logw
    logz_cond = np.logaddexp.reduce(logp)
    post = np.exp(logp - logz_cond)
    info = float(np.sum(post * all_l) - logz_cond)       # H = integral of P ln(L / Z)
    if not return_dead:
        post = post[-n_live:] / post[-n_live:].sum()
    return NestedResult(
        logZ=float(logz_cond + np.log(finite_fraction)), logZ_err=float(np.sqrt(max(info, 0.0) / n_live)), information=info,
        samples=all_p, weights=post / post.sum(), lnpost=all_l if return_dead else live_l, n_evals=int(n_evals), n_iter=int(it),
        efficiency=n_acc / float(max(n_evals - n_init, 1)), finite_fraction=finite_fraction, n_live=n_live,
        param_names=tuple(mod.param_names), n_batches=n_batches, converged=bool(done and it < max_iter))
