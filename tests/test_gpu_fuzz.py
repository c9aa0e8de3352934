"""GPU: randomized parity of the CUDA path against the oracle on grids of irregular SHAPE.

The other GPU suites use the MIST shapes (15 x 196 x 1710, 70 x 26 x 18 x 13).  Here every case draws its own axis
lengths (2 ... 300 nodes, including the lengths around the powers of two a bracketing search cares about), replaces
the uniform axes by irregular ones in half of the cases (so the closed-form EEP lookup AND the table search are both
exercised on every axis position), and observes a random subset of properties.  Rows mix walker-like, prior-box,
on-node, edge, out-of-bounds and NaN inputs.  Everything is seeded: a failure reproduces from the case number.

Tolerances as in test_gpu_oracle.py: identical NaN / -inf patterns, lnpost within 1e-4 absolute (rtol 1e-12),
interpolated properties within 1e-6 relative (asserted at 1e-11).
"""
import numpy as np
import pytest

from tests.helpers import assert_same_special

pytestmark = pytest.mark.gpu

ALL_BANDS = ("V", "J", "H", "K", "G", "BP", "RP", "W1", "W2")
LENGTHS = (2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 64, 65, 100)


def _irregular(axis, rng):
    """Same end points, same length, strictly increasing, no longer uniform (and not exactly representable steps)."""
    n = len(axis)
    if n < 3:
        return axis.copy()
    w = 0.25 + rng.random_sample(n - 1)
    out = axis[0] + (axis[-1] - axis[0]) * np.concatenate([[0.0], np.cumsum(w) / w.sum()])
    out[-1] = axis[-1]
    if not np.all(np.diff(out) > 0):     # a rounding collision next to an end point: keep the original axis
        return axis.copy()
    return out


def _case(case):
    from isochrones_b200 import synthetic as syn

    rng = np.random.RandomState(1000 + case)
    kind = "track" if case % 2 == 0 else "iso"
    n_stars = 1 if kind == "track" else 1 + (case // 2) % 3
    n_eep = int(rng.choice([40, 129, 257, 300]))
    if kind == "track":
        model = syn.make_track_grid(n_feh=int(rng.choice(LENGTHS[:6])), n_mass=int(rng.choice(LENGTHS[3:])), n_eep=n_eep)
    else:
        model = syn.make_iso_grid(n_age=int(rng.choice(LENGTHS[3:])), n_feh=int(rng.choice(LENGTHS[:6])), n_eep=n_eep)
    n_b = int(rng.randint(1, len(ALL_BANDS) + 1))
    bands = tuple(rng.choice(ALL_BANDS, n_b, replace=False))
    bc = syn.make_bc_grid(bands=bands, n_teff=int(rng.choice(LENGTHS[3:])), n_logg=int(rng.choice(LENGTHS[1:9])),
                          n_feh=int(rng.choice(LENGTHS[1:9])), n_av=int(rng.choice(LENGTHS[:7])))
    # the generators tabulate at most 15 [Fe/H] and 13 Av values; longer requests repeat nodes -> spread them evenly
    fix = lambda a: a if np.all(np.diff(a) > 0) else np.linspace(np.min(a), np.max(a), len(a))
    model["axes"] = tuple(fix(np.asarray(a, dtype=float)) for a in model["axes"])
    bc["axes"] = tuple(fix(np.asarray(a, dtype=float)) for a in bc["axes"])
    if case % 4 >= 2:   # irregular axes everywhere (the values on the nodes stay what they were: any table is a grid)
        model["axes"] = tuple(_irregular(np.asarray(a, dtype=float), rng) for a in model["axes"])
        bc["axes"] = tuple(_irregular(np.asarray(a, dtype=float), rng) for a in bc["axes"])
    return kind, n_stars, n_eep, model, bc, bands, rng


@pytest.mark.parametrize("case", range(12))
def test_random_grid_shapes_vs_oracle(case):
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn
    from oracle import oracle

    kind, N, n_eep, model, bc, bands, rng = _case(case)
    ctx = _lib.default_context()
    ic = ib.ichrone_from_arrays(kind, model, bc, ctx=ctx)
    og_m, og_b = oracle.Grid(model["grid"], model["axes"]), oracle.Grid(bc["grid"], bc["axes"])
    truth = syn.default_truth(kind, n_eep=n_eep, n_stars=N)

    # ---- standalone interpolation (DFInterpolator.__call__ / interp_values_3d) on every column ---------------------
    axes = [np.asarray(a, dtype=float) for a in model["axes"]]
    n = 3000
    pts = np.stack([a[0] + (a[-1] - a[0]) * (1.2 * rng.random_sample(n) - 0.1) for a in axes], axis=1)   # 17 % outside
    on = rng.randint(0, n, n // 4)
    for d, a in enumerate(axes):
        pts[on[d::3], d] = rng.choice(a, len(on[d::3]))                                                      # on nodes
    pts[rng.randint(0, n, 20), rng.randint(0, 3, 20)] = np.nan
    cols = list(range(len(model["columns"])))
    got = ic.model_grid.interp([pts[:, 0], pts[:, 1], pts[:, 2]], list(model["columns"]))
    want = og_m.interp_values([pts[:, 0], pts[:, 1], pts[:, 2]], cols)
    assert_same_special(got, want)
    m = np.isfinite(want)
    assert np.allclose(got[m], want[m], rtol=1e-11, atol=0.0)

    # ---- the fused lnpost kernel with a random set of observed properties ------------------------------------------
    prim = list(truth) if kind == "track" else [truth[0]] + list(truth[N:])
    _, _, _, mags = ic.interp_mag(prim, list(bands))
    kw = {b: (float(m_) - (0.3 if N > 1 else 0.0), 0.03) for b, m_ in zip(bands, mags) if np.isfinite(m_) and rng.rand() < 0.8}
    if rng.rand() < 0.7:
        kw["Teff"] = (5772.0, 90.0)
    if rng.rand() < 0.7:
        kw["logg"] = (4.44, 0.12)
    if rng.rand() < 0.7:
        kw["feh"] = (0.0, 0.15)
    if rng.rand() < 0.7:
        kw["parallax"] = (10.0, 0.2)
    if N == 1 and rng.rand() < 0.4:
        kw["nu_max"] = (3000.0, 50.0)
        kw["delta_nu"] = (135.0, 2.0)
    if not kw:
        kw["parallax"] = (10.0, 0.2)
    mod = ib.BasicStarModel(ic, N=N, **kw)
    om = oracle.StarModel(mod, model_grid=og_m, bc_grid=og_b)
    bounds = [mod.bounds(p) for p in mod.param_names]
    rows = np.concatenate([syn.posterior_like_batch(kind, 3000, truth, n_eep=n_eep, seed=case),
                           syn.prior_like_batch(kind, 3000, bounds, seed=100 + case),
                           syn.edge_batch(kind, 800, truth, model["axes"], bounds, seed=200 + case)])
    if N > 1:   # half of the walker-like rows respect the ordering prior
        rows[:1500, :N] = -np.sort(-rows[:1500, :N], axis=1)
    lnp, lnprior, lnlike = mod.lnpost_batch(rows, parts=True)
    w_lnp, w_lnprior, w_lnlike = om.lnpost_batch(rows, parts=True)
    for g, w, atol in ((lnp, w_lnp, 1e-4), (lnprior, w_lnprior, 1e-9), (lnlike, w_lnlike, 1e-4)):
        assert_same_special(g, w)
        f = np.isfinite(w)
        assert np.allclose(g[f], w[f], rtol=1e-12, atol=atol), float(np.max(np.abs(g[f] - w[f])))
    assert np.array_equal(mod.lnpost_batch(rows), lnp, equal_nan=True)   # lnpost-only launch (early outs) == full evaluation
    assert np.isfinite(w_lnp).sum() > 200


def test_packed_and_column_entry_points_agree():
    """iso_interp_mags (pars[5, N] parameter-major, the reference's layout at mags.py:86-87) and iso_interp_mags_cols
    (five separate arrays, what the Python mirror passes) are the same computation."""
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn

    ctx = _lib.default_context()
    trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=171)
    rows = syn.posterior_like_batch("track", 5000, truth, n_eep=171, seed=5)
    rows[::50, 1] = 1e9                                   # some rows outside the grid
    bands = ["V", "J", "H", "K"]
    teff, logg, feh, mags = ic.interp_mag([rows[:, j].copy() for j in range(5)], bands)
    p = np.ascontiguousarray(rows.T)                      # [5, N]
    n = p.shape[1]
    t2, l2, f2, m2 = np.empty(n), np.empty(n), np.empty(n), np.empty((n, 4))
    io = np.array(ic.param_index_order, dtype=np.int32)
    bc_cols = np.arange(4, dtype=np.int32)
    ctx.check(_lib.lib().iso_interp_mags(ctx.handle, ic.model_pack.handle, ic.bc_pack(bands).handle, _lib.ip(io), 0, 1, 2, 3,
                                         _lib.ip(bc_cols), 4, _lib.dp(p), n, _lib.dp(t2), _lib.dp(l2), _lib.dp(f2), _lib.dp(m2)))
    for a, b in ((teff, t2), (logg, l2), (feh, f2), (mags, m2)):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(t2[::50]).all() and np.isfinite(t2).sum() > 4000
    # scalars and mixed scalar / array parameters broadcast as in the reference
    s = ic.interp_mag(list(rows[3]), bands)
    assert s[0] == teff[3] and np.array_equal(s[3], mags[3])
    mixed = ic.interp_mag([rows[:, 0].copy(), rows[:, 1].copy(), 0.0, 100.0, 0.1], bands)
    want = ic.interp_mag([rows[:, 0].copy(), rows[:, 1].copy(), np.zeros(n), np.full(n, 100.0), np.full(n, 0.1)], bands)
    assert np.array_equal(mixed[3], want[3], equal_nan=True)


def test_device_buffer_interpolation_entry_points():
    """iso_interp_values_device / iso_interp_mags_device (inputs and outputs in HBM) == the host-pointer calls."""
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn

    ctx = _lib.default_context()
    trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=171)
    rows = syn.posterior_like_batch("track", 7001, truth, n_eep=171, seed=11)
    rows[::37, 0] = 1e6
    n = len(rows)
    cols = [np.ascontiguousarray(rows[:, j]) for j in range(5)]
    d_cols = []
    for c in cols:
        d = ctx.dev_alloc(c.nbytes)
        ctx.h2d(d, c)
        d_cols.append(d)
    bands = ["V", "J", "H", "K"]
    want = ic.interp_mag(cols, bands)
    d_out = [ctx.dev_alloc(n * 8) for _ in range(3)] + [ctx.dev_alloc(n * 8 * 4)]
    ic.interp_mag_device(d_cols, n, bands, *d_out)
    for w, d in zip(want, d_out):
        got = np.empty_like(w)
        ctx.d2h(got, d)
        assert np.array_equal(got, w, equal_nan=True)
    # interp_value on the model grid: axes order (feh, mass, eep) = parameters (2, 0, 1)
    props = ["Teff", "radius", "age"]
    grid = ic.model_grid.interp
    icols = [grid.column_index[p] for p in props]
    want_v = grid([cols[2], cols[0], cols[1]], props)
    d_v = ctx.dev_alloc(n * 8 * len(props))
    grid.device_grid.interp_values_device([d_cols[2], d_cols[0], d_cols[1]], n, icols, d_v)
    got_v = np.empty_like(want_v)
    ctx.d2h(got_v, d_v)
    assert np.array_equal(got_v, want_v, equal_nan=True)
    for d in d_cols + d_out + [d_v]:
        ctx.dev_free(d)


def test_interp_mags_generic_and_packed_kernels_agree():
    """iso_interp_mags on the FULL grids with arbitrary column indices (generic kernel: scalar gathers) gives the same
    numbers as the packed fast path the Python mirror uses (model pack + BC pack, vector gathers)."""
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn

    ctx = _lib.default_context()
    trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171)
    bands_all = ("V", "J", "H", "K", "G")
    bc = syn.make_bc_grid(bands=bands_all, n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=171)
    rows = syn.posterior_like_batch("track", 4000, truth, n_eep=171, seed=21)
    rows[::29, 2] = 9.0
    bands = ["K", "V", "G"]                                           # a subset, out of order
    want = ic.interp_mag([rows[:, j].copy() for j in range(5)], bands)    # packed kernel
    mgrid, bgrid = ic.model_grid.interp, ic.bc_grid.interp
    ci = mgrid.column_index
    p = np.ascontiguousarray(rows.T)
    n = p.shape[1]
    t2, l2, f2, m2 = np.empty(n), np.empty(n), np.empty(n), np.empty((n, len(bands)))
    io = np.array(ic.param_index_order, dtype=np.int32)
    bc_cols = np.array([bgrid.column_index[b] for b in bands], dtype=np.int32)
    ctx.check(_lib.lib().iso_interp_mags(ctx.handle, mgrid.device_grid.handle, bgrid.device_grid.handle, _lib.ip(io),
                                         ci["Teff"], ci["logg"], ci["feh"], ci["Mbol"], _lib.ip(bc_cols), len(bands),
                                         _lib.dp(p), n, _lib.dp(t2), _lib.dp(l2), _lib.dp(f2), _lib.dp(m2)))
    for a, b in zip(want, (t2, l2, f2, m2)):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(t2).sum() >= n // 29 and np.isfinite(m2).all(axis=1).sum() > 3000
