"""Deterministic MIST-*shaped* synthetic grids (no MIST data / network in this environment).

The reference builds three dense float64 arrays:

* evolution-track grid ``[feh, mass, EEP, 18 columns]`` (reference ``mist/models.py:167``),
* isochrone grid ``[log10 age, feh, EEP, 16 columns]`` (``mist/models.py:99``),
* bolometric-correction grid ``[Teff, logg, [Fe/H], Av, bands]`` (``bc.py:25``, ``mist/bc.py:161-163``).

This module produces arrays of the same shape, axis values, column names, NaN-tail
structure and physical value ranges from smooth analytic functions, so that every
code path of the hot path (in-bounds, NaN-padded tails, model-grid OOB, BC-grid OOB)
is exercised.  Sizes are parameters: the benchmark uses the full MIST v1.2 shape,
the tests use small ones.

The track and isochrone grids are mutually consistent: both are generated from the same
closed-form ``age(mass, feh, EEP)`` relation, inverted analytically for the isochrones.
"""
import numpy as np

# MIST v1.2 [Fe/H] nodes (reference mist/models.py:39-57)
MIST_FEHS = np.array(
    [-4.0, -3.5, -3.0, -2.5, -2.0, -1.75, -1.5, -1.25, -1.0, -0.75, -0.5, -0.25, 0.0, 0.25, 0.5]
)
MIST_N_EEP = 1710  # reference mist/models.py:63

# Av nodes of the MIST BC tables
MIST_AVS = np.array([0.0, 0.05, 0.1, 0.15, 0.2, 0.3, 0.4, 0.6, 0.8, 1.0, 2.0, 4.0, 6.0])

TRACK_COLUMNS = (
    "nu_max", "logg", "eep", "initial_mass", "radius", "logTeff", "mass", "density", "Mbol",
    "phase", "feh", "Teff", "logL", "delta_nu", "interpolated", "star_age", "age", "dt_deep",
)
ISO_COLUMNS = (
    "eep", "age", "feh", "mass", "initial_mass", "radius", "density", "logTeff", "Teff", "logg",
    "logL", "Mbol", "delta_nu", "nu_max", "phase", "dm_deep",
)

# grid limits the reference hard-codes for MIST (mist/models.py:37)
MIST_LIMITS = {"age": (5, 10.13), "feh": (-4, 0.5), "eep": (0, 1710), "mass": (0.1, 300)}

# extinction coefficient A_band / A_V and a zero point / colour slope per band
_BAND_COEFFS = {
    "V": (1.00, -0.10, 6.0), "B": (1.32, -0.60, 9.0), "J": (0.29, 1.40, -3.0), "H": (0.18, 1.90, -4.5),
    "K": (0.11, 2.00, -5.0), "G": (0.86, 0.05, 4.5), "BP": (1.07, -0.25, 7.0), "RP": (0.65, 0.60, 1.5),
    "W1": (0.07, 2.10, -5.3), "W2": (0.05, 2.15, -5.4), "W3": (0.03, 2.20, -5.5), "TESS": (0.60, 0.65, 1.2),
    "Kepler": (0.85, 0.10, 4.0),
}

_AGE0 = 5.0          # log10 age at the start of every track
_G_RATE = 4.0        # shape of the age(EEP) curve
_T_A, _T_B, _T_C = 10.2, 1.9, 0.1   # log10 lifetime = A - B log10(m) + C feh


def max_eep_table(mass, feh):
    """Last valid EEP of a MIST v1.2 track (data table restated from reference mist/utils.py:1-59)."""
    special = {
        -4.0: ((lambda m: m < 0.6, 454), (lambda m: m <= 0.94, 631), (lambda m: m < 3.8, 808),
               (lambda m: m <= 4.4, 1409), (lambda m: m >= 18, 631)),
        -3.5: ((lambda m: m == 0.65, 631), (lambda m: 0.65 < m < 1.78, 808), (lambda m: m == 1.78, 1409),
               (lambda m: 1.78 < m <= 3.4, 808), (lambda m: m >= 19, 707)),
        -3.0: ((lambda m: 0.7 <= m <= 2.48, 808), (lambda m: 2.5 <= m <= 4.4, 1409)),
        -2.5: ((lambda m: 0.7 <= m <= 2.32, 808), (lambda m: 2.32 < m <= 5.8, 1409)),
        0.5: ((lambda m: 0.7 <= m <= 0.75, 808),),
    }
    for cond, eep in special.get(float(feh), ()):
        if cond(mass):
            return eep
    if mass < 0.6:
        return 454
    if mass == 0.6:
        return 605
    if mass == 0.65:
        return 808
    if mass < 6.0:
        return 1710
    return 808


def _g(x):
    return (1.0 - np.exp(-_G_RATE * x)) / (1.0 - np.exp(-_G_RATE))


def _dg(x):
    return _G_RATE * np.exp(-_G_RATE * x) / (1.0 - np.exp(-_G_RATE))


def _stellar_columns(m, x, f, eep):
    """Smooth 'stellar' properties from initial mass m, phase x in (0,1), initial feh f."""
    lm = np.log10(m)
    bump = x ** 2 * (3.0 - 2.0 * x)                    # 0 -> 1 along the track (giant-branch cooling)
    logteff_ms = 3.80 + 0.45 * np.tanh(lm / 0.8) - 0.02 * f
    logteff = logteff_ms - 0.45 * (logteff_ms - 3.45) * bump + 0.01 * np.sin(6.0 * x)
    logg = 4.45 - 0.35 * lm - 3.6 * x ** 2 + 0.05 * f * x
    logl = 3.6 * np.tanh(lm / 1.4) * 1.6 + 2.4 * x ** 2 - 0.05 * f
    feh_s = f - 0.02 * x * (1.0 + 0.5 * np.tanh(lm)) + 0.005
    mass = m * (1.0 - 0.1 * x ** 4)
    logr = 0.5 * (np.log10(mass) - logg + 4.438)
    radius = 10.0 ** logr
    teff = 10.0 ** logteff
    density = 1.41 * mass / radius ** 3
    nu_max = 3090.0 * mass / radius ** 2 / np.sqrt(teff / 5777.0)
    delta_nu = 135.1 * np.sqrt(mass / radius ** 3)
    phase = np.floor(6.0 * x) - 1.0
    return {
        "nu_max": nu_max, "logg": logg, "eep": eep + 0.0 * m, "initial_mass": m + 0.0 * x, "radius": radius,
        "logTeff": logteff, "mass": mass, "density": density, "Mbol": 4.74 - 2.5 * logl, "phase": phase,
        "feh": feh_s, "Teff": teff, "logL": logl, "delta_nu": delta_nu, "interpolated": 0.0 * m * x,
    }


def mist_like_masses(n_mass=196, seed=0):
    """Sorted initial-mass axis in [0.1, 300], always containing 0.6, 0.65 and 1.0 when n_mass >= 5."""
    rng = np.random.RandomState(seed)
    base = np.logspace(np.log10(0.1), np.log10(300.0), n_mass)
    jitter = 1.0 + 0.02 * (rng.rand(n_mass) - 0.5)
    jitter[0] = jitter[-1] = 1.0
    masses = np.round(base * jitter, 4)
    masses[0], masses[-1] = 0.1, 300.0
    if n_mass >= 5:
        for special in (0.6, 0.65, 1.0):
            masses[np.argmin(np.abs(masses - special))] = special
    masses = np.unique(masses)
    while len(masses) < n_mass:  # extremely unlikely collision: fill in log-midpoints
        gaps = np.argmax(np.diff(np.log(masses)))
        masses = np.sort(np.append(masses, np.round(np.sqrt(masses[gaps] * masses[gaps + 1]), 5)))
    return masses


def make_track_grid(n_feh=15, n_mass=196, n_eep=MIST_N_EEP, columns=TRACK_COLUMNS, seed=0, nan_tails=True):
    """``{"grid": [n_feh, n_mass, n_eep, ncols] float64, "axes": (feh, mass, eep), "columns", "limits"}``."""
    if n_feh == len(MIST_FEHS):
        fehs = MIST_FEHS.copy()
    else:
        fehs = MIST_FEHS[np.round(np.linspace(0, len(MIST_FEHS) - 1, n_feh)).astype(int)]
    masses = mist_like_masses(n_mass, seed)
    eeps = np.arange(1, n_eep + 1, dtype=float)
    scale = MIST_N_EEP / float(n_eep)   # small test grids still span the 1..1710 phase range
    f = fehs[:, None, None]
    m = masses[None, :, None]
    e = eeps[None, None, :]
    x = (e * scale) / (MIST_N_EEP + 1.0)
    cols = _stellar_columns(m, x, f, e)
    t_max = _T_A - _T_B * np.log10(m) + _T_C * f
    age = _AGE0 + (t_max - _AGE0) * _g(x)
    cols["age"] = age
    cols["star_age"] = 10.0 ** age
    cols["dt_deep"] = (t_max - _AGE0) * _dg(x) * scale / (MIST_N_EEP + 1.0)
    shape = (len(fehs), len(masses), len(eeps))
    grid = np.empty(shape + (len(columns),))
    for i, c in enumerate(columns):
        grid[..., i] = np.broadcast_to(cols[c], shape)
    if nan_tails:
        for i_f, fv in enumerate(fehs):
            for i_m, mv in enumerate(masses):
                last = int(max_eep_table(mv, fv) / scale)
                if last < n_eep:
                    grid[i_f, i_m, last:, :] = np.nan
    limits = dict(MIST_LIMITS)
    limits["eep"] = (0, n_eep)
    return {"grid": grid, "axes": (fehs, masses, eeps), "columns": list(columns), "limits": limits, "kind": "track"}


def make_iso_grid(n_age=107, n_feh=15, n_eep=MIST_N_EEP, columns=ISO_COLUMNS):
    """``{"grid": [n_age, n_feh, n_eep, ncols], "axes": (age, feh, eep), ...}``; NaN where no star exists."""
    ages = np.round(np.linspace(5.0, 10.3, n_age), 6)
    if n_feh == len(MIST_FEHS):
        fehs = MIST_FEHS.copy()
    else:
        fehs = MIST_FEHS[np.round(np.linspace(0, len(MIST_FEHS) - 1, n_feh)).astype(int)]
    eeps = np.arange(1, n_eep + 1, dtype=float)
    scale = MIST_N_EEP / float(n_eep)
    a = ages[:, None, None]
    f = fehs[None, :, None]
    e = eeps[None, None, :]
    x = (e * scale) / (MIST_N_EEP + 1.0)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        t_max = _AGE0 + (a - _AGE0) / _g(x)
        lm = (_T_A + _T_C * f - t_max) / _T_B
        m = 10.0 ** lm
        valid = (m >= 0.1) & (m <= 300.0)
        m_safe = np.where(valid, m, 1.0)
        cols = _stellar_columns(m_safe, x + 0.0 * m_safe, f + 0.0 * m_safe, e)
        cols["age"] = a + 0.0 * m_safe
        dm_dx = m_safe * np.log(10.0) * (a - _AGE0) * _dg(x) / (_T_B * _g(x) ** 2)
        cols["dm_deep"] = dm_dx * scale / (MIST_N_EEP + 1.0)
    shape = (len(ages), len(fehs), len(eeps))
    grid = np.empty(shape + (len(columns),))
    for i, c in enumerate(columns):
        grid[..., i] = np.broadcast_to(cols[c], shape)
    grid[~np.broadcast_to(valid, shape)] = np.nan
    # MIST tracks of the corresponding mass also end early: reuse the track tail rule on the iso grid
    limits = dict(MIST_LIMITS)
    limits["eep"] = (0, n_eep)
    return {"grid": grid, "axes": (ages, fehs, eeps), "columns": list(columns), "limits": limits, "kind": "iso"}


def make_bc_grid(bands=("V", "J", "H", "K"), n_teff=70, n_logg=26, n_feh=18, n_av=13):
    """``{"grid": [n_teff, n_logg, n_feh, n_av, n_bands], "axes": (Teff, logg, feh, Av), "columns": bands}``."""
    teffs = np.round(np.logspace(np.log10(2500.0), np.log10(50000.0), n_teff), 3)
    loggs = np.round(np.linspace(-4.0, 9.5, n_logg), 6)
    fehs = np.round(np.linspace(-4.0, 0.75, n_feh), 6)
    if n_av == len(MIST_AVS):
        avs = MIST_AVS.copy()
    else:
        avs = MIST_AVS[np.round(np.linspace(0, len(MIST_AVS) - 1, n_av)).astype(int)]
    t = np.log10(teffs)[:, None, None, None] - 3.76
    g = loggs[None, :, None, None]
    f = fehs[None, None, :, None]
    av = avs[None, None, None, :]
    shape = (len(teffs), len(loggs), len(fehs), len(avs))
    grid = np.empty(shape + (len(bands),))
    for i, b in enumerate(bands):
        if b in _BAND_COEFFS:
            k, zp, slope = _BAND_COEFFS[b]
        else:  # unknown band names still get a deterministic smooth table
            h = sum(ord(ch) for ch in b)
            k, zp, slope = 0.05 + (h % 13) / 10.0, (h % 7) / 3.0 - 1.0, (h % 11) - 5.0
        bc = zp + slope * t - 2.5 * t ** 2 + 0.02 * g * (1.0 + t) + 0.05 * f * (1.0 - 0.3 * t)
        bc = bc - k * av * (1.0 + 0.02 * t - 0.003 * av)
        grid[..., i] = np.broadcast_to(bc, shape)
    return {"grid": grid, "axes": (teffs, loggs, fehs, avs), "columns": list(bands), "kind": "bc"}


# ---------------------------------------------------------------------------
# Synthetic parameter batches: rows are (mass, eep, feh, distance, AV) for track
# models and (eep_0[, eep_1[, eep_2]], age, feh, distance, AV) for isochrone models.
# ---------------------------------------------------------------------------

def default_truth(kind, n_eep=MIST_N_EEP, n_stars=1):
    """A Sun-like truth point, scaled to the EEP axis length of the grid."""
    s = n_eep / float(MIST_N_EEP)
    if kind == "track":
        return np.array([1.0, 350.0 * s, 0.0, 100.0, 0.1])
    eeps = [860.0 * s, 800.0 * s, 740.0 * s][:n_stars]
    return np.array(eeps + [9.5, 0.0, 100.0, 0.1])


def posterior_like_batch(kind, n, truth, n_eep=MIST_N_EEP, seed=2):
    """Gaussian ball around ``truth`` (sigma: mass .05, eep 15, age .1, feh .1, distance 2 pc, AV .05)."""
    rng = np.random.RandomState(seed)
    truth = np.asarray(truth, dtype=float)
    ndim = len(truth)
    s = n_eep / float(MIST_N_EEP)
    if kind == "track":
        sig = np.array([0.05, 15.0 * s, 0.1, 2.0, 0.05])
    else:
        n_stars = ndim - 4
        sig = np.array([15.0 * s] * n_stars + [0.1, 0.1, 2.0, 0.05])
    p = truth + sig * rng.standard_normal((n, ndim))
    p[:, -1] = np.abs(p[:, -1])          # AV >= 0 (reflected, keeps most rows inside the prior support)
    return p


def prior_like_batch(kind, n, bounds, seed=3):
    """Uniform over the model's ``bounds(par)`` box; ``bounds`` is a sequence of (lo, hi) per parameter."""
    rng = np.random.RandomState(seed)
    lo = np.array([b[0] for b in bounds], dtype=float)
    hi = np.array([b[1] for b in bounds], dtype=float)
    return lo + (hi - lo) * rng.random_sample((n, len(lo)))


def edge_batch(kind, n, truth, model_axes, bounds, seed=7):
    """Rows that sit exactly on grid nodes / edges, out of bounds, NaN, or next to NaN tails."""
    rng = np.random.RandomState(seed)
    truth = np.asarray(truth, dtype=float)
    ndim = len(truth)
    n_stars = ndim - 4 if kind == "iso" else 1
    p = np.tile(truth, (n, 1))
    if kind == "track":
        ax_of_par = {0: model_axes[1], 1: model_axes[2], 2: model_axes[0]}
    else:
        ax_of_par = {i: model_axes[2] for i in range(n_stars)}
        ax_of_par[n_stars] = model_axes[0]
        ax_of_par[n_stars + 1] = model_axes[1]
    for i in range(n):
        mode = i % 8
        j = rng.randint(ndim)
        lo, hi = bounds[j]
        if mode == 0 and j in ax_of_par:          # exactly on a node
            p[i, j] = rng.choice(ax_of_par[j])
        elif mode == 1 and j in ax_of_par:        # exactly on the lower / upper edge
            p[i, j] = ax_of_par[j][0] if rng.rand() < 0.5 else ax_of_par[j][-1]
        elif mode == 2:                           # outside the prior / grid support
            p[i, j] = lo - 0.1 * (hi - lo) * rng.rand() if rng.rand() < 0.5 else hi + 0.1 * (hi - lo) * rng.rand()
        elif mode == 3:                           # NaN coordinate
            p[i, j] = np.nan
        elif mode == 4:                           # all model coordinates on nodes
            for jj, ax in ax_of_par.items():
                p[i, jj] = rng.choice(ax)
        elif mode == 5:                           # uniform over the box
            for jj in range(ndim):
                p[i, jj] = bounds[jj][0] + (bounds[jj][1] - bounds[jj][0]) * rng.rand()
        elif mode == 6:                           # large extinction / tiny distance (BC-grid OOB, log of 0)
            p[i, -1] = rng.choice([0.0, 0.999, 1.0, 1.5])
            p[i, -2] = rng.choice([0.0, 1e-3, 50.0, 1999.0])
        else:                                     # jittered truth
            p[i] = truth * (1.0 + 0.02 * rng.standard_normal(ndim))
    return p
