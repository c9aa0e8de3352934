"""MIST bolometric-correction tables -> the dense ``(Teff, logg, [Fe/H], Av)`` grid the device path stages
(lets real MIST BC data flow into ``ichrone_from_arrays`` / ``BCGrid``).

The reference keeps these tables as one HDF file per photometric system (``bc.py:99-119``), which it writes itself
from the text tables of the MIST tarballs (``*.UBVRIplus``, ``*.WISE`` ... one file per [Fe/H]) with
``parse_table`` (``bc.py:72-83``: column names on the sixth line, ``#`` comments, whitespace-separated numbers,
index ``(Teff, logg, [Fe/H], Av, Rv)``), slices ``Rv = 3.1`` (``mist/bc.py:161-163``) and hands the frame to
``DFInterpolator`` (``is_full`` grid, ``bc.py:27``).  pytables / h5py do not exist in this environment, so this loader
starts from the text tables — the format the reference itself parses — with numpy only.

Band names: ``mist_band`` restates the shortcut rules of ``MISTBolometricCorrectionGrid.get_band``
(``mist/bc.py:165-240``) as a lookup table + the same ``System_band`` fallback.
"""
import glob
import os
import re

import numpy as np

INDEX_COLS = ("Teff", "logg", "[Fe/H]", "Av", "Rv")          # bc.py:25

# the photometric systems MIST tabulates (the keys of ``phot_bands``, mist/bc.py:7-158) and the prefix of the column
# names each system's tables use — enough to find the system of a full column name such as ``Bessell_V``
PHOT_SYSTEMS = ("UBVRIplus", "WISE", "CFHT", "DECam", "GALEX", "JWST", "LSST", "PanSTARRS", "SkyMapper", "SPITZER",
                "UKIDSS", "SDSSugriz", "HST_ACSWF", "HST_ACSHR", "HST_WFC3", "HST_WFPC2")
_COLUMN_PREFIX = {"2MASS": "UBVRIplus", "Bessell": "UBVRIplus", "Gaia": "UBVRIplus", "Hipparcos": "UBVRIplus",
                  "Kepler": "UBVRIplus", "Tycho": "UBVRIplus", "WISE": "WISE", "CFHT": "CFHT", "DECam": "DECam",
                  "GALEX": "GALEX", "LSST": "LSST", "PS": "PanSTARRS", "SkyMapper": "SkyMapper", "IRAC": "SPITZER",
                  "UKIDSS": "UKIDSS", "SDSS": "SDSSugriz", "ACS": "HST_ACSWF", "WFC3": "HST_WFC3", "WFC3vIR": "HST_WFC3",
                  "WFPC2": "HST_WFPC2"}

_SHORTCUTS = {}
_SHORTCUTS.update({b: ("SDSSugriz", "SDSS_" + b) for b in "ugriz"})
_SHORTCUTS.update({b: ("UBVRIplus", "Bessell_" + b) for b in "UBVRI"})
_SHORTCUTS.update({b: ("UBVRIplus", "2MASS_" + b) for b in ("J", "H", "Ks")})
_SHORTCUTS["K"] = ("UBVRIplus", "2MASS_Ks")
_SHORTCUTS.update({b: ("UBVRIplus", "Kepler_Kp") for b in ("kep", "Kepler", "Kp")})
_SHORTCUTS["TESS"] = ("UBVRIplus", "TESS")
_SHORTCUTS.update({b: ("WISE", "WISE_" + b) for b in ("W1", "W2", "W3", "W4")})
_SHORTCUTS.update({b: ("UBVRIplus", "Gaia_%s_DR2Rev" % b) for b in ("G", "BP", "RP")})
_SHORTCUTS.update({"Bp": ("UBVRIplus", "Gaia_BP_DR2Rev"), "Rp": ("UBVRIplus", "Gaia_RP_DR2Rev")})


def mist_band(b):
    """``(photometric system, table column)`` of a band name (mist/bc.py:165-240)."""
    if b in _SHORTCUTS:
        return _SHORTCUTS[b]
    m = re.match("([a-zA-Z]+)_([a-zA-Z_]+)", b)
    if m:
        if m.group(1) in PHOT_SYSTEMS:
            phot = m.group(1)
            return phot, ("PS_" + m.group(2)) if phot == "PanSTARRS" else m.group(0)
        if m.group(1) in ("UK", "UKIRT"):
            return "UKIDSS", "UKIDSS_" + m.group(2)
    # a full table column name: the system is the one whose tables use that prefix (the reference searches its
    # per-system column lists; JWST columns are bare filter names)
    prefix = b.split("_", 1)[0]
    if b.startswith("ACS_HRC_"):
        return "HST_ACSHR", b
    if "_" in b and prefix in _COLUMN_PREFIX:
        return _COLUMN_PREFIX[prefix], b
    if re.fullmatch(r"F\d{3}[WMN]2?", b):
        return "JWST", b
    raise ValueError("MIST grids cannot resolve band {}!".format(b))


def parse_bc_table(filename):
    """One text table -> ``(column names, float64 array [n_rows, n_columns])`` (bc.py:72-83)."""
    names = None
    with open(filename) as fin:
        for i, line in enumerate(fin):
            if i == 5:
                names = line[1:].split()
                break
    if not names:
        raise ValueError("%s: no column names on line 6" % filename)
    data = np.loadtxt(filename, comments="#", ndmin=2)
    if data.shape[1] != len(names):
        raise ValueError("%s: %d columns of numbers for %d names" % (filename, data.shape[1], len(names)))
    return names, data


def _dense(index, values):
    """Rows keyed by ``index[n_rows, ndim]`` -> dense ``[n0, .., n_{d-1}, ncols]`` array (NaN where a combination is
    absent) + the sorted axis values — what ``DFInterpolator._make_grid`` does with a MultiIndex frame."""
    axes, codes = [], []
    for d in range(index.shape[1]):
        a, c = np.unique(index[:, d], return_inverse=True)
        axes.append(a)
        codes.append(c)
    grid = np.full(tuple(len(a) for a in axes) + (values.shape[1],), np.nan)
    grid[tuple(codes)] = values
    return grid, tuple(axes)


def load_mist_bc_grid(datadir, bands, Rv=3.1):
    """Dense BC grid of ``bands`` from the MIST text tables under ``datadir`` (all ``*.<phot system>`` files of every
    system the bands need): ``{"grid": [nT, ng, nf, nA, n_bands], "axes": (Teff, logg, [Fe/H], Av), "columns": bands,
    "kind": "bc"}`` — the dict ``ichrone_from_arrays`` / ``BCGrid`` take."""
    bands = list(bands)
    datadir = os.path.expanduser(datadir)
    wanted = {}
    for b in bands:
        phot, col = mist_band(b)
        wanted.setdefault(phot, []).append((b, col))
    per_band, axes = {}, None
    for phot, cols in wanted.items():
        files = sorted(glob.glob(os.path.join(datadir, "*.{}".format(phot))))
        if not files:
            raise FileNotFoundError("no *.%s tables under %s (the reference downloads them from "
                                    "http://waps.cfa.harvard.edu/MIST/BC_tables/%s.txz)" % (phot, datadir, phot))
        names, blocks = None, []
        for f in files:
            n, data = parse_bc_table(f)
            if names is None:
                names = n
            elif n != names:
                raise ValueError("%s: columns differ from %s" % (f, files[0]))
            blocks.append(data)
        data = np.concatenate(blocks, axis=0)
        col_of = {n: i for i, n in enumerate(names)}
        missing = [c for c in INDEX_COLS if c not in col_of] + [c for _, c in cols if c not in col_of]
        if missing:
            raise KeyError("%s tables lack column(s) %s" % (phot, ", ".join(missing)))
        keep = data[:, col_of["Rv"]] == Rv                         # df.xs(3.1, level="Rv")  mist/bc.py:161-163
        if not keep.any():
            raise ValueError("%s tables hold no rows with Rv = %r" % (phot, Rv))
        data = data[keep]
        index = data[:, [col_of[c] for c in INDEX_COLS[:4]]]
        grid, ax = _dense(index, data[:, [col_of[c] for _, c in cols]])
        if axes is None:
            axes = ax
        elif any(len(a) != len(b_) or not np.array_equal(a, b_) for a, b_ in zip(axes, ax)):
            raise ValueError("photometric systems are tabulated on different (Teff, logg, [Fe/H], Av) nodes")
        for j, (b, _) in enumerate(cols):
            per_band[b] = grid[..., j]
    grid = np.stack([per_band[b] for b in bands], axis=-1)
    return {"grid": np.ascontiguousarray(grid), "axes": axes, "columns": bands, "kind": "bc"}
