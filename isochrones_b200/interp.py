"""``DFInterpolator`` — drop-in for the reference's ``isochrones/interp.py:571-698`` with the interpolation
running on the GPU.

Same constructor arguments, attributes (``grid``, ``index_columns``, ``columns``, ``column_index``,
``n_columns``, ``ndim``, ``index_names``) and call convention: ``interp(p, cols)`` with all-scalar ``p``
returns a 1-D array of ``len(cols)`` values, anything else is broadcast and returns ``(N, ncols)``
(interp.py:631-698).  ``NaN`` means "outside the grid".  The dense grid is staged once to HBM on first use;
each call is one CUDA launch (``iso_interp_values`` replaces ``interp_value(s)_{2,3,4}d``).
"""
import os

import numpy as np

from . import _lib


class DFInterpolator(object):
    """Interpolate column values of a DataFrame with a full-grid hierarchical index (interp.py:571-614)."""

    def __init__(self, df, filename=None, recalc=False, is_full=False, ctx=None):
        self.filename = filename
        self.is_full = is_full
        self.columns = list(df.columns)
        self.n_columns = len(self.columns)
        self.grid = self._make_grid(df, recalc=recalc)
        self.index_columns = tuple(np.array(l, dtype=float) for l in df.index.levels)
        self.index_names = df.index.names
        self.ndim = len(self.index_columns)
        self.column_index = {c: self.columns.index(c) for c in self.columns}
        self._ctx = ctx
        self._device_grid = None

    @classmethod
    def from_arrays(cls, grid, index_columns, columns, index_names=None, ctx=None):
        """Build directly from a dense ``grid[n0, .., ncols]`` and its axes (no DataFrame round trip)."""
        self = cls.__new__(cls)
        self.filename = None
        self.is_full = True
        self.columns = [str(c) for c in columns]
        self.n_columns = len(self.columns)
        self.grid = np.ascontiguousarray(grid, dtype=float)
        self.index_columns = tuple(np.array(a, dtype=float) for a in index_columns)
        self.index_names = list(index_names) if index_names is not None else [None] * len(self.index_columns)
        self.ndim = len(self.index_columns)
        self.column_index = {c: i for i, c in enumerate(self.columns)}
        assert self.grid.shape == tuple(len(a) for a in self.index_columns) + (self.n_columns,)
        self._ctx = ctx
        self._device_grid = None
        return self

    @classmethod
    def from_npz(cls, filename, index_columns, index_names=None, ctx=None):
        """Load a dense grid the reference cached with ``np.savez(filename, grid=grid, columns=columns)``
        (interp.py:611-612; e.g. ``full_grid*.npz`` written by ``Grid.interp``, grid.py:132-137).  The axes are not in
        that file (the reference takes them from ``df.index.levels``), so they are passed in."""
        d = np.load(filename, allow_pickle=False)
        grid = d["grid"]
        columns = [str(c) for c in d["columns"]]
        if grid.ndim != len(index_columns) + 1 or grid.shape[-1] != len(columns):
            raise ValueError("grid of shape %r does not match %d axes / %d columns"
                             % (grid.shape, len(index_columns), len(columns)))
        return cls.from_arrays(grid, index_columns, columns, index_names=index_names, ctx=ctx)

    def _make_grid(self, df, recalc=False):
        """Dense ``[n0, .., n_{d-1}, ncols]`` float64 array of a MultiIndex frame; index combinations the frame does
        not hold stay NaN.  Host-side, one-time data preparation (the reference's interp.py:590-614); a ``filename``
        caches the array in the reference's ``.npz`` layout (keys ``grid``, ``columns``)."""
        cached = self.filename is not None and os.path.exists(self.filename) and not recalc
        if cached:
            with np.load(self.filename) as d:
                grid, columns = d["grid"], d["columns"]
            if not all(columns == self.columns):
                raise ValueError("DataFrame columns do not match columns loaded from full grid!")
            return grid
        levels_shape = tuple(len(level) for level in df.index.levels)
        grid = np.full(levels_shape + (len(df.columns),), np.nan)
        # scatter the rows through the index codes: no reindexing of the frame, works for full and ragged grids alike
        grid[tuple(np.asarray(c) for c in df.index.codes)] = np.asarray(df.values, dtype=float)
        if self.filename is not None:
            np.savez(self.filename, grid=grid, columns=self.columns)
        return grid

    # ---- device residency -------------------------------------------------------------------------------
    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    @property
    def device_grid(self):
        """The grid staged in HBM (created on first use; dropped by ``add_column``)."""
        if self._device_grid is None:
            self._device_grid = _lib.DeviceGrid(self.ctx, self.grid, self.index_columns)
        return self._device_grid

    def add_column(self, values, name):
        """Append one column (``values`` broadcastable to the grid's node shape); drops the staged device copy."""
        extra = np.broadcast_to(np.asarray(values, dtype=float), self.grid.shape[:-1])[..., None]
        self.grid = np.concatenate([self.grid, extra], axis=-1)
        self.columns = self.columns + [name]
        self.column_index[name] = self.n_columns
        self.n_columns = len(self.columns)
        self._device_grid = None

    def __call__(self, p, cols="all", out=None):
        """``interp(p, cols)`` of the reference (interp.py:631-672).  ``out`` (optional, array calls): a C-contiguous
        float64 ``[N, len(cols)]`` array to fill (a page-locked one is written by DMA directly)."""
        if isinstance(cols, str) and cols == "all":
            icols = np.arange(self.n_columns)
        else:
            icols = np.array([self.column_index[col] for col in cols])
        if len(p) != self.ndim:
            raise ValueError("expected %d coordinates, got %d" % (self.ndim, len(p)))
        scalar = all(isinstance(x, (float, int)) for x in p)     # interp.py:638-666 (np.float64 is a float)
        if scalar:
            pp = [np.array([x], dtype=float) for x in p]
        else:
            b = np.broadcast(*p)
            # coordinates that already are full-size float64 arrays are passed as they are (no np.resize copy)
            pp = [np.ravel(x) if isinstance(x, np.ndarray) and x.dtype == np.float64 and x.size == b.size
                  else np.atleast_1d(np.resize(x, b.shape)).astype(float).ravel() for x in p]
        values = self.device_grid.interp_values(pp, icols, out=None if scalar else out)
        return values[0] if scalar else values
