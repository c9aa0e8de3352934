"""Real MIST data -> the dense arrays the device path stages.

What the reference leaves on disk under ``$ISOCHRONES`` (``config.py:5``: the environment variable, else
``~/.isochrones``) after its first use of a MIST grid, and what this module reads of it:

====================================  =========================================================  ==================
file                                  written by the reference at                                read here
====================================  =========================================================  ==================
``mist/full_grid_v1.2_vvcrit0.4_      ``DFInterpolator._make_grid`` ``interp.py:590-614`` via    yes (``grid``,
full_isos.npz``                       ``Grid.interp`` ``grid.py:132-137``, name                  ``columns``)
                                      ``models.py:163-165``: isochrone grid
                                      ``[n_age, 15, 1710, 16]``
``mist/tracks/full_grid_v1.2_         same, evolution-track grid ``[15, n_mass, 1710, 18]``      yes
vvcrit0.4.npz``
``mist/*.h5``, ``mist/tracks/*.h5``   pandas / pytables caches of the frames (``grid.py:99-      no (pytables /
                                      114``, ``models.py:120-149``, ``mist/models.py:395-430``)  h5py do not exist
                                                                                                 here)
``BC/mist/feh[mp]???.<system>``       text tables of the MIST BC tarballs, parsed by             yes
                                      ``bc.py:72-83``                                            (:mod:`.bcio`)
``BC/mist/<system>.h5``               HDF cache of those tables (``bc.py:99-118``)               no
====================================  =========================================================  ==================

The reference's ``.npz`` cache holds the dense array and the column names but NOT the axes — the reference takes those
from ``df.index.levels`` of the HDF frame (``interp.py:583``).  Here the axes come, in this order, from

1. a sidecar ``full_grid<tag>.axes.npz`` next to the cache (keys ``axis0``, ``axis1``, ``axis2``, ``index_names``) —
   written by :func:`write_axes_sidecar` / :func:`save_model_grid`, or by ``tools/export_mist_axes.py`` on a machine
   where the reference itself is installed;
2. the grid itself: the index levels are also columns of the frame — ``initial_mass`` and ``eep`` for track grids
   (``mist/models.py:167``), ``age`` and ``eep`` for isochrone grids (``mist/models.py:99``; ``age`` is the renamed
   ``log10_isochrone_age_yr`` index column) — so every axis value is the (constant) value of that column over the
   populated nodes of its slice; the [Fe/H] axis is the tabulated MIST list ``mist/models.py:39-57`` (the ``feh``
   column is the *surface* abundance, ``mist/models.py:81``, not the index).

Nothing here runs on the hot path: this is one-time host-side data preparation; the arrays go to HBM through
``ichrone_from_arrays``.
"""
import os

import numpy as np

from . import bcio

# mist/models.py:39-57
MIST_FEHS = np.array([-4.00, -3.50, -3.00, -2.50, -2.00, -1.75, -1.50, -1.25, -1.00, -0.75, -0.50, -0.25, 0.00, 0.25,
                      0.50])
# mist/models.py:37 — the `bounds` class attribute every MIST grid starts its `_limits` from (grid.py:56)
MIST_LIMITS = {"age": (5, 10.13), "feh": (-4, 0.5), "eep": (0, 1710), "mass": (0.1, 300)}
INDEX_NAMES = {"track": ("initial_feh", "initial_mass", "EEP"),              # mist/models.py:167
               "iso": ("log10_isochrone_age_yr", "feh", "EEP")}              # mist/models.py:99
DEFAULT_BANDS = ("J", "H", "K", "G", "BP", "RP", "W1", "W2", "W3", "TESS", "Kepler")   # mist/bc.py:159


class MistDataNotFound(FileNotFoundError):
    """No MIST data under ``$ISOCHRONES`` — there is no download here (no network) and no silent synthetic stand-in."""


def isochrones_root(root=None):
    """``config.ISOCHRONES`` of the reference (config.py:5)."""
    if root is None:
        root = os.getenv("ISOCHRONES", os.path.join("~", ".isochrones"))
    return os.path.abspath(os.path.expanduser(root))


def model_grid_files(kind, root=None, version="1.2", vvcrit=0.4, iso_kind="full_isos"):
    """``(datadir, dense-grid cache, axes sidecar)`` of a MIST grid as the reference names them: track grids live in
    ``mist/tracks`` with tag ``_v{version}_vvcrit{vvcrit}`` (mist/models.py:196-203), isochrone grids in ``mist`` with
    the grid kind appended (mist/models.py:103-106); the cache is ``full_grid{tag}.npz`` (models.py:163-165)."""
    if kind not in ("track", "iso"):
        raise ValueError("kind must be 'track' or 'iso'")
    tag = "_v{}_vvcrit{}".format(version, vvcrit)
    datadir = os.path.join(isochrones_root(root), "mist")
    if kind == "track":
        datadir = os.path.join(datadir, "tracks")
    else:
        tag = "{}_{}".format(tag, iso_kind)
    base = os.path.join(datadir, "full_grid{}".format(tag))
    return datadir, base + ".npz", base + ".axes.npz"


def _constant_over_slices(values, axis, what):
    """The value a column takes on every populated node of each slice along ``axis`` (NaN nodes ignored)."""
    other = tuple(a for a in range(values.ndim) if a != axis)
    with np.errstate(invalid="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)    # all-NaN slices are reported below
            lo = np.nanmin(values, axis=other)
            hi = np.nanmax(values, axis=other)
    if np.isnan(lo).any():
        raise ValueError("cannot derive the %s axis: slice(s) %s hold no populated node (write an axes sidecar with "
                         "write_axes_sidecar)" % (what, np.flatnonzero(np.isnan(lo)).tolist()))
    if not np.array_equal(lo, hi):
        raise ValueError("cannot derive the %s axis: the column is not constant over a slice" % what)
    if np.any(np.diff(lo) <= 0):
        raise ValueError("derived %s axis is not strictly increasing" % what)
    return lo.astype(float)


def derive_axes(kind, grid, columns):
    """Axes of a dense MIST grid from its own index-valued columns (see the module docstring, item 2)."""
    ci = {str(c): i for i, c in enumerate(columns)}
    if grid.ndim != 4:
        raise ValueError("a model grid is [n0, n1, n_eep, ncols]; got shape %r" % (grid.shape,))
    if "eep" not in ci:
        raise KeyError("grid has no 'eep' column")
    eeps = _constant_over_slices(grid[..., ci["eep"]], 2, "EEP")
    if kind == "track":
        if "initial_mass" not in ci:
            raise KeyError("track grid has no 'initial_mass' column")
        n_feh = grid.shape[0]
        if n_feh != len(MIST_FEHS):
            raise ValueError("track grid holds %d [Fe/H] slices, MIST tabulates %d: the axis cannot be derived (the 'feh' "
                             "column is the surface abundance); write an axes sidecar" % (n_feh, len(MIST_FEHS)))
        return MIST_FEHS.copy(), _constant_over_slices(grid[..., ci["initial_mass"]], 1, "initial mass"), eeps
    if "age" not in ci:
        raise KeyError("isochrone grid has no 'age' column")
    if grid.shape[1] != len(MIST_FEHS):
        raise ValueError("isochrone grid holds %d [Fe/H] slices, MIST tabulates %d; write an axes sidecar"
                         % (grid.shape[1], len(MIST_FEHS)))
    return _constant_over_slices(grid[..., ci["age"]], 0, "log10 age"), MIST_FEHS.copy(), eeps


def write_axes_sidecar(filename, axes, index_names=None):
    """``full_grid<tag>.axes.npz``: the index levels the reference's ``.npz`` cache lacks."""
    axes = [np.asarray(a, dtype=float) for a in axes]
    names = np.array([str(n) for n in (index_names or [""] * len(axes))])
    np.savez(filename, index_names=names, **{"axis%d" % i: a for i, a in enumerate(axes)})


def read_axes_sidecar(filename):
    with np.load(filename, allow_pickle=False) as d:
        n = len([k for k in d.files if k.startswith("axis")])
        return tuple(np.array(d["axis%d" % i], dtype=float) for i in range(n))


def save_model_grid(model, kind, root=None, sidecar=True, **tag):
    """Write a dense model grid (``{"grid", "columns", "axes"}``) in the reference's cache layout — ``np.savez(filename,
    grid=grid, columns=columns)``, interp.py:611-612 — at the path the reference would use, plus the axes sidecar.
    Returns the cache path.  (Export for users who build grids elsewhere; the tests use it to fabricate a
    ``$ISOCHRONES`` tree.)"""
    datadir, npz, axes_file = model_grid_files(kind, root, **tag)
    os.makedirs(datadir, exist_ok=True)
    np.savez(npz, grid=np.asarray(model["grid"], dtype=float), columns=[str(c) for c in model["columns"]])
    if sidecar:
        write_axes_sidecar(axes_file, model["axes"], INDEX_NAMES[kind])
    return npz


def load_model_grid(kind, root=None, limits=None, **tag):
    """The reference's dense-grid cache -> ``{"grid", "axes", "columns", "limits", "kind", "source"}`` (the dict
    ``ichrone_from_arrays`` takes).  Raises :class:`MistDataNotFound` when the cache is absent."""
    datadir, npz, axes_file = model_grid_files(kind, root, **tag)
    if not os.path.exists(npz):
        raise MistDataNotFound(
            "no MIST %s grid at %s.  The reference writes this cache the first time the grid is interpolated "
            "(isochrones.get_ichrone('mist'%s).initialize()); copy its $ISOCHRONES tree here or point ISOCHRONES at it.  "
            "For the synthetic MIST-shaped benchmark grids pass synthetic=True." % (
                "evolution-track" if kind == "track" else "isochrone", npz, ", tracks=True" if kind == "track" else ""))
    with np.load(npz, allow_pickle=False) as d:
        grid = np.ascontiguousarray(d["grid"], dtype=float)
        columns = [str(c) for c in d["columns"]]
    if grid.ndim != 4 or grid.shape[-1] != len(columns):
        raise ValueError("%s: grid of shape %r with %d column names" % (npz, grid.shape, len(columns)))
    if os.path.exists(axes_file):
        axes = read_axes_sidecar(axes_file)
        how = "sidecar"
    else:
        axes = derive_axes(kind, grid, columns)
        how = "derived from the grid's index-valued columns"
    if tuple(len(a) for a in axes) != grid.shape[:3]:
        raise ValueError("%s: axes of lengths %r for a grid of shape %r" % (npz, tuple(len(a) for a in axes), grid.shape))
    lim = dict(MIST_LIMITS)
    lim.update(limits or {})
    return {"grid": grid, "axes": tuple(axes), "columns": columns, "limits": lim, "kind": kind,
            "index_names": INDEX_NAMES[kind], "source": "%s (axes: %s)" % (npz, how)}


def bc_cache_file(bands, root=None):
    key = "_".join(bands)
    return os.path.join(isochrones_root(root), "BC", "mist", "b200_bc_%s.npz" % key)


def load_bc_grid(bands=None, root=None, cache=True):
    """Dense ``(Teff, logg, [Fe/H], Av)`` BC grid of ``bands`` from ``$ISOCHRONES/BC/mist`` (bc.py:61-63): the MIST text
    tables through :func:`bcio.load_mist_bc_grid` (Rv = 3.1 slice, mist/bc.py:161-163).  Parsing the ~18 tables of a
    photometric system takes a few seconds, so the dense result is cached next to them as ``b200_bc_<bands>.npz``
    (this package's own file: grid, columns, four axes)."""
    bands = list(bands) if bands is not None else list(DEFAULT_BANDS)
    datadir = os.path.join(isochrones_root(root), "BC", "mist")
    cfile = bc_cache_file(bands, root)
    if cache and os.path.exists(cfile):
        with np.load(cfile, allow_pickle=False) as d:
            if [str(c) for c in d["columns"]] == bands:
                return {"grid": np.ascontiguousarray(d["grid"], dtype=float), "columns": bands, "kind": "bc",
                        "axes": tuple(np.array(d["axis%d" % i], dtype=float) for i in range(4)), "source": cfile}
    if not os.path.isdir(datadir):
        raise MistDataNotFound("no MIST bolometric-correction tables under %s (the reference extracts "
                               "http://waps.cfa.harvard.edu/MIST/BC_tables/<system>.txz there, bc.py:61-97); for the "
                               "synthetic benchmark grids pass synthetic=True" % datadir)
    try:
        bc = bcio.load_mist_bc_grid(datadir, bands)
    except FileNotFoundError as e:
        raise MistDataNotFound(str(e))
    bc["source"] = datadir
    if cache:
        try:
            np.savez(cfile, grid=bc["grid"], columns=bands, **{"axis%d" % i: a for i, a in enumerate(bc["axes"])})
        except OSError:
            pass    # read-only data directory: parse again next time
    return bc
