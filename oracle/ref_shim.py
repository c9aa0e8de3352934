"""TEST INFRASTRUCTURE — loads the *unmodified* reference package in-process.

Only usable in the build container where ``/root/reference`` exists (it does not
travel to the GPU box).  Used by ``oracle/make_golden.py`` to generate the
committed golden vectors under ``tests/golden/`` and by the CPU-side tests that
pin the C oracle (``oracle/iso_oracle.c``) against the real reference.

Nothing under ``isochrones_b200/`` may import this module.

The reference's ``isochrones/__init__.py:8-10`` eagerly imports ``starmodel`` which
imports emcee/corner/tables/... (absent here), so the package ``__init__`` is
bypassed and the absent third-party modules are replaced by empty stand-ins
(recipe: SURVEY.md Appendix B).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ISOCHRONES_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "isochrones"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Const:
    # astropy.constants stand-in: reference models.py:19-21 reads ``.cgs.value``
    def __init__(self, v):
        self.cgs = types.SimpleNamespace(value=v)


_loaded = None


def load():
    """Return a namespace with the reference modules on the lnpost path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    import numpy, pandas, scipy, numba  # noqa: F401  (must be real, import first)

    if "astropy" not in sys.modules:
        _stub("astropy")
        _stub("astropy.constants", G=_Const(6.6743e-8), M_sun=_Const(1.98840987e33), R_sun=_Const(6.957e10))
        _stub("astropy.coordinates", SkyCoord=object)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    for name in ("emcee", "corner", "tables"):
        if name not in sys.modules:
            _stub(name)
    if "configobj" not in sys.modules:
        _stub("configobj", ConfigObj=object, Section=object)
    if "asciitree" not in sys.modules:
        _stub("asciitree", LeftAligned=object, Traversal=object)
        _stub("asciitree.drawing", BoxStyle=object, BOX_DOUBLE=None, BOX_BLANK=None)

    pkg = types.ModuleType("isochrones")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "isochrones")]
    pkg.__version__ = "2.1"
    sys.modules["isochrones"] = pkg

    ns = types.SimpleNamespace()
    for mod in ("interp", "mags", "likelihood", "utils", "priors", "models", "starmodel"):
        setattr(ns, mod, importlib.import_module("isochrones." + mod))
    _loaded = ns
    return ns


# ---------------------------------------------------------------------------
# Synthetic reference objects: the reference's own classes wired to in-memory
# grids (no MIST download).  Attribute use: interp.py:631-698, models.py:416-427,
# starmodel.py:1544-1552, 1577-1596.
# ---------------------------------------------------------------------------

def make_ref_interp(grid, index_columns, columns):
    """A reference ``DFInterpolator`` built directly from a dense array."""
    import numpy as np

    ref = load()

    class ArrayDFInterpolator(ref.interp.DFInterpolator):
        def __init__(self, grid, index_columns, columns):
            self.filename = None
            self.is_full = True
            self.columns = list(columns)
            self.n_columns = len(self.columns)
            self.grid = np.ascontiguousarray(grid, dtype=float)
            self.index_columns = tuple(np.array(a, dtype=float) for a in index_columns)
            self.index_names = [None] * len(self.index_columns)
            self.ndim = len(self.index_columns)
            self.column_index = {c: i for i, c in enumerate(self.columns)}

    return ArrayDFInterpolator(grid, index_columns, columns)


class _FakeGrid:
    def __init__(self, interp, limits, eep_replaces, name="synthetic"):
        self.interp = interp
        self._limits = dict(limits)
        self.eep_replaces = eep_replaces
        self.name = name
        # ModelGridInterpolator.__call__ asks the grid's frame for its column names only (models.py:477)
        import types
        self.df = types.SimpleNamespace(columns=list(interp.columns))

    def get_limits(self, prop):
        return self._limits[prop]


class _FakeBC:
    def __init__(self, interp, bands):
        self.interp = interp
        self.bands = list(bands)


def make_ref_ic(kind, model, bc, eep_bounds=(0, 1710)):
    """Reference ``EvolutionTrackInterpolator`` / ``IsochroneInterpolator`` on synthetic grids.

    ``model`` / ``bc`` are dicts with keys grid, axes, columns (+ limits for model).
    """
    ref = load()
    base = ref.models.EvolutionTrackInterpolator if kind == "track" else ref.models.IsochroneInterpolator
    replaces = "age" if kind == "track" else "mass"
    _eep_bounds = eep_bounds

    class SyntheticIC(base):
        eep_bounds = _eep_bounds

        def __init__(self):
            self.bands = list(bc["columns"])
            self.param_index_order = list(self._param_index_order)
            self.kwargs = {}
            self._fehs = self._ages = self._masses = None
            self._model_grid = _FakeGrid(
                make_ref_interp(model["grid"], model["axes"], model["columns"]), model["limits"], replaces
            )
            self._bc_grid = _FakeBC(make_ref_interp(bc["grid"], bc["axes"], bc["columns"]), bc["columns"])
            self._iso = None
            self._track = None

    return SyntheticIC()
