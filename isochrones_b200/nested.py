"""Nested sampling over the unit hypercube with every likelihood evaluation batched on the GPU — the B200 form of the
reference's DEFAULT fit path, ``pymultinest.run(mod.mnest_loglike, mod.mnest_prior, n)`` (``starmodel.py:717-802``).

MultiNest itself is third-party and un-vendored (``.ci/travis.sh:20``); what the reference defines is the problem it
hands over: live points live in the unit cube, ``mnest_prior`` maps a cube point to the parameters with the affine box
map of ``starmodel.py:1637-1640`` and the "likelihood" of a point is ``mnest_loglike`` = ``lnpost`` (``:1642-1645``: the
parameter priors are part of the integrand, the cube measure is uniform).  The quantity MultiNest reports is therefore

    Z = integral over the unit cube of exp(lnpost(theta(u))) du

and this module estimates the same Z, and the same weighted posterior sample, with the textbook algorithm (Skilling
2006) in the rejection-within-bounding-ellipsoids form MultiNest popularised (Feroz & Hobson 2008): the live points are
split recursively with 2-means wherever two ellipsoids bound them in markedly less volume than one, so curved
degeneracies and separate modes are followed by a union of small ellipsoids, each enlarged by a margin and never smaller
than the prior volume its points stand for:

* the cube -> parameters -> lnpost chain of a whole batch of candidate points is ONE launch
  (``iso_mnest_lnpost_batch``); candidates are drawn uniformly inside the current union of ellipsoids (intersected with
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


def _unit_ball_log_volume(ndim):
    from math import lgamma, log, pi

    return 0.5 * ndim * log(pi) - lgamma(0.5 * ndim + 1.0)


class _Ellipsoid(object):
    """Bounding ellipsoid of a point set: the covariance ellipsoid scaled to hold every point (radius times ``enlarge``),
    and never smaller than ``min_logvol`` — the prior volume the points are expected to stand for; a handful of points
    would otherwise draw an ellipsoid that misses part of the region they sample."""

    def __init__(self, pts, enlarge, min_logvol=-np.inf):
        ndim = pts.shape[1]
        self.mu = pts.mean(axis=0)
        d = pts - self.mu
        cov = d.T @ d / max(len(pts) - 1, 1)
        cov += np.eye(ndim) * (1e-12 * max(np.trace(cov), 1e-300) + 1e-300)
        self.L = np.linalg.cholesky(cov)
        z = np.linalg.solve(self.L, d.T)                          # whitened points
        self.r = enlarge * np.sqrt(np.max(np.sum(z * z, axis=0)))
        logdet = np.sum(np.log(np.diag(self.L)))
        self.logvol = _unit_ball_log_volume(ndim) + ndim * np.log(self.r) + logdet
        if self.logvol < min_logvol:
            self.r *= np.exp((min_logvol - self.logvol) / ndim)
            self.logvol = min_logvol

    def draw(self, n, rng):
        ndim = len(self.mu)
        z = rng.standard_normal((n, ndim))
        z *= (rng.random(n) ** (1.0 / ndim) / np.sqrt(np.sum(z * z, axis=1)))[:, None]      # uniform in the unit ball
        return self.mu + self.r * (z @ self.L.T)

    def contains(self, x):
        z = np.linalg.solve(self.L, (x - self.mu).T)
        return np.sum(z * z, axis=0) <= self.r * self.r


def _two_means(pts, n_iter=16):
    """Labels of a 2-means split of ``pts`` (cube coordinates), started from the two extreme points along the direction of
    largest variance."""
    w, v = np.linalg.eigh(np.cov(pts.T))
    proj = (pts - pts.mean(axis=0)) @ v[:, -1]
    centres = np.stack([pts[np.argmin(proj)], pts[np.argmax(proj)]])
    labels = None
    for _ in range(n_iter):
        d0 = np.sum((pts - centres[0]) ** 2, axis=1)
        d1 = np.sum((pts - centres[1]) ** 2, axis=1)
        new = d1 < d0
        if labels is not None and np.array_equal(new, labels):
            break
        labels = new
        if labels.all() or not labels.any():
            break
        centres = np.stack([pts[~labels].mean(axis=0), pts[labels].mean(axis=0)])
    return labels


def _bounding_ellipsoids(pts, enlarge, log_pointvol, ell=None, depth=0):
    """MultiNest's decomposition (Feroz & Hobson 2008, section 5.1, as restated in the public nested-sampling literature):
    split the live points in two with 2-means wherever the two bounding ellipsoids together are markedly smaller than the
    one — curved degeneracies and separate modes are then followed by a chain of small ellipsoids instead of one that
    fills the cube.  ``log_pointvol``: ln of the prior volume one live point stands for."""
    n, ndim = pts.shape
    if ell is None:
        ell = _Ellipsoid(pts, enlarge, log_pointvol + np.log(n))
    if n < 4 * (ndim + 1) or depth > 12:
        return [ell]
    labels = _two_means(pts)
    n1 = int(labels.sum())
    if min(n1, n - n1) < 2 * (ndim + 1):
        return [ell]
    a, b = pts[~labels], pts[labels]
    ea = _Ellipsoid(a, enlarge, log_pointvol + np.log(len(a)))
    eb = _Ellipsoid(b, enlarge, log_pointvol + np.log(len(b)))
    both = np.logaddexp(ea.logvol, eb.logvol)
    if both < ell.logvol + np.log(0.5):
        return (_bounding_ellipsoids(a, enlarge, log_pointvol, ea, depth + 1) +
                _bounding_ellipsoids(b, enlarge, log_pointvol, eb, depth + 1))
    if ell.logvol > np.log(2.0) + log_pointvol + np.log(n):
        # no gain at this level, but the ellipsoid is far larger than the volume its points stand for: look deeper
        out = (_bounding_ellipsoids(a, enlarge, log_pointvol, ea, depth + 1) +
               _bounding_ellipsoids(b, enlarge, log_pointvol, eb, depth + 1))
        if np.logaddexp.reduce([e.logvol for e in out]) < ell.logvol + np.log(0.5):
            return out
    return [ell]


def _draw_from_union(ells, n, rng):
    """``<= n`` points uniform in the union of the ellipsoids: pick an ellipsoid by volume, draw inside it, keep the point
    with probability 1 / (number of ellipsoids that contain it)."""
    if len(ells) == 1:
        return ells[0].draw(n, rng)
    logv = np.array([e.logvol for e in ells])
    p = np.exp(logv - np.logaddexp.reduce(logv))
    which = rng.choice(len(ells), size=n, p=p / p.sum())
    pts = np.empty((n, len(ells[0].mu)))
    for k, e in enumerate(ells):
        m = which == k
        if m.any():
            pts[m] = e.draw(int(m.sum()), rng)
    inside = np.zeros(n)
    for e in ells:
        inside += e.contains(pts)
    return pts[rng.random(n) * np.maximum(inside, 1.0) < 1.0]


def nested_sample(mod, n_live=1000, dlogz=0.5, seed=0, enlarge=1.25, batch=8192, max_iter=5_000_000, max_batches=20_000,
                  return_dead=True, multi=True, max_init_draws=2_000_000_000):
    """Nested sampling of ``mod`` (a ``BasicStarModel``): returns a :class:`NestedResult`.

    ``n_live``: live points (MultiNest's ``n_live_points``, starmodel.py:667-671 default 1000); ``dlogz``: stop when the
    live points can add at most this much to ln Z (MultiNest's ``evidence_tolerance`` 0.5); ``batch``: candidate points
    per launch; ``enlarge``: linear margin of the bounding ellipsoids; ``multi``: decompose the bound into several
    ellipsoids (MultiNest's scheme; ``False`` = one ellipsoid around all live points)."""
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
        if n_tried > max_init_draws and len(live_l) < n_live:
            raise RuntimeError("no finite lnpost in the prior box (%d of %d live points after %d draws)" % (len(live_l), n_live, n_tried))
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
    n_ell_max = 1
    while not done and n_batches < max_batches:
        # bound of the region {lnpost > worst live point}: a union of ellipsoids around clusters of the live points,
        # rebuilt for every batch; candidates = uniform draws from the union that fall inside the cube (redrawn until the
        # batch is full: a bound that sticks out of the cube costs host draws, not likelihood evaluations)
        ells = (_bounding_ellipsoids(live_u, enlarge, log_x + np.log(finite_fraction) - np.log(n_live)) if multi
                else [_Ellipsoid(live_u, enlarge)])
        n_ell_max = max(n_ell_max, len(ells))
        got, n_got = [], 0
        for _ in range(64):
            c = _draw_from_union(ells, batch, rng)
            c = c[np.all((c >= 0.0) & (c <= 1.0), axis=1)]
            got.append(c)
            n_got += len(c)
            if n_got >= batch:
                break
        cand = np.concatenate(got)[:batch]
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
        param_names=tuple(mod.param_names), n_batches=n_batches, n_ellipsoids_max=n_ell_max, converged=bool(done and it < max_iter))
