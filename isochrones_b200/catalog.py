"""``StarCatalog`` — host mirror of the reference's ``isochrones/catalog.py:19-139`` for the part that feeds the
lnpost path: a table of star measurements (``<band>_mag`` / ``<band>_mag_unc`` and ``<prop>`` / ``<prop>_unc``
columns) becomes one star model per row.

The reference yields one Python ``SingleStarModel`` per row (``iter_models``, catalog.py:126-139) and fits them one
after the other (or through a process pool / SLURM array).  Here ``compile(ic)`` packs ALL rows into one device array
of ``iso_model`` structs — built vectorised from one template model, whose prior objects (and ``set_prior``
overrides, catalog.py:116-124) every star shares — so that a single launch evaluates rows of many stars
(``model_of_row``) or runs one sampler chain per star (``DeviceEnsembleSampler(..., n_chains=len(catalog))``).
"""
import ctypes as C
import re

import numpy as np

from . import _lib
from .starmodel import BinaryStarModel, CompiledModel, SingleStarModel, TripleStarModel


class StarCatalog(object):
    def __init__(self, df, bands=None, props=None, no_uncs=False):
        self._df = df
        if bands is None:      # every "<band>_mag" column names a band
            bands = [m.group(1) for m in (re.fullmatch("(.+)_mag", str(c)) for c in df.columns) if m]
        self.bands = tuple(bands)
        self.band_cols = tuple(b + "_mag" for b in self.bands)
        self.props = tuple(props or ())
        if not no_uncs:
            missing = [c for c in self.band_cols + self.props if c not in df.columns]
            if missing:
                raise ValueError("{} not in DataFrame!".format(missing[0]))
            no_unc = [c for c in self.band_cols + self.props if c + "_unc" not in df.columns]
            if no_unc:
                raise ValueError("{0} uncertainty ({0}_unc) not in DataFrame!".format(no_unc[0]))
        self._prior_settings = {}

    df = property(lambda self: self._df)

    def __len__(self):
        return len(self._df)

    def get_measurement(self, prop, values=False):
        """``(values, uncertainties)`` of one measured quantity over the whole table."""
        return self._df[prop].values, self._df[prop + "_unc"].values

    def iter_bands(self, **kwargs):
        return ((b, self.get_measurement(c, **kwargs)) for b, c in zip(self.bands, self.band_cols))

    def iter_props(self, **kwargs):
        return ((p, self.get_measurement(p, **kwargs)) for p in self.props)

    def set_prior(self, **kwargs):
        """Prior settings applied to every model (catalog.py:116-124)."""
        self._prior_settings.update(kwargs)

    def _model(self, ic, i, N):
        mod_type = {1: SingleStarModel, 2: BinaryStarModel, 3: TripleStarModel}
        row = self.df.iloc[i]
        mags = {b: (row["{}_mag".format(b)], row["{}_mag_unc".format(b)]) for b in self.bands}
        props = {p: (row[p], row["{}_unc".format(p)]) for p in self.props}
        mod = mod_type[N](ic, **mags, **props, name=row.name)
        mod.set_prior(**self._prior_settings)
        return mod

    def iter_models(self, ic, N=1):
        """One host model object per row, as the reference does (slow path: each constructor integrates its priors)."""
        for i in range(len(self.df)):
            yield self._model(ic, i, N)

    # ---- the device path ---------------------------------------------------------------------------------------
    def build_structs(self, ic, N=1, maxAV=None, max_distance=None):
        """``(ctypes array of len(self) iso_model structs, bands)`` — equivalent to ``[m.to_struct() for m in
        self.iter_models(ic, N)]`` but vectorised: the struct of a template model is replicated and the per-star
        fields are written column-wise — observations, the NaN-means-absent rule (starmodel.py:1427-1430) and the
        parallax-dependent distance bound (starmodel.py:1468-1476).  Host-only (no GPU needed)."""
        n = len(self.df)
        if n == 0:
            raise ValueError("empty catalog")
        kw = {}
        if maxAV is not None:
            kw["maxAV"] = maxAV
        if max_distance is not None:
            kw["max_distance"] = max_distance
        mod_type = {1: SingleStarModel, 2: BinaryStarModel, 3: TripleStarModel}
        # template: every band / prop present (finite placeholders) so that the struct has all slots populated
        tmpl_kwargs = {b: (10.0, 0.1) for b in self.bands}
        tmpl_kwargs.update({p: (1.0, 0.1) for p in self.props})
        tmpl = mod_type[N](ic, **kw, **tmpl_kwargs)
        tmpl.set_prior(**self._prior_settings)
        bands = [b for b in tmpl.bands]                       # bands the BC grid knows, in kwargs order
        col = {b: i for i, b in enumerate(bands)}
        base = tmpl.to_struct(band_columns=col)
        arr = (_lib.IsoModel * n)()
        rec = np.frombuffer(arr, dtype=np.dtype(_lib.IsoModel))
        rec[:] = np.frombuffer((_lib.IsoModel * 1)(base), dtype=np.dtype(_lib.IsoModel))[0]
        for f in ("band_col", "mag_val", "mag_unc"):
            rec[f] = 0

        def meas(name):
            v, u = self.get_measurement(name)
            v, u = np.asarray(v, dtype=float), np.asarray(u, dtype=float)
            ok = ~(np.isnan(v) | np.isnan(u))
            return v, u, ok

        # photometry: a star's observed bands are packed to the front of its band list
        n_bands = np.zeros(n, dtype=np.int32)
        for b in bands:
            v, u, ok = meas("{}_mag".format(b))
            slot = n_bands.copy()
            idx = np.nonzero(ok)[0]
            rec["band_col"][idx, slot[idx]] = col[b]
            rec["mag_val"][idx, slot[idx]] = v[idx]
            rec["mag_unc"][idx, slot[idx]] = u[idx]
            n_bands[idx] += 1
        rec["n_bands"] = n_bands
        for i, p in enumerate(("Teff", "logg", "feh")):
            if p in self.props:
                v, u, ok = meas(p)
                rec["spec_val"][:, i] = np.where(ok, v, np.nan)
                rec["spec_unc"][:, i] = np.where(ok, u, np.nan)
            else:
                rec["spec_val"][:, i] = np.nan
                rec["spec_unc"][:, i] = np.nan
        for p, flag, fv, fu in (("parallax", "has_plax", "plax", "plax_unc"), ("nu_max", "has_nu_max", "nu_max", "nu_max_unc"),
                                ("delta_nu", "has_delta_nu", "delta_nu", "delta_nu_unc")):
            if p in self.props:
                v, u, ok = meas(p)
                rec[flag] = ok.astype(np.int32)
                rec[fv] = np.where(ok, v, np.nan)
                rec[fu] = np.where(ok, u, np.nan)
            else:
                rec[flag] = 0
        if "parallax" in self.props and max_distance is None and "distance" not in self._prior_settings:
            # starmodel.py:1468-1476: distance bounds (0, 2000 / parallax) for positive parallaxes, (0, 2000 / |unc|) for
            # negative ones; the template carries the bound of its placeholder parallax, undo that for absent ones
            v, u, ok = meas("parallax")
            default_hi = 10000.0
            # same expression as the reference (1.0 / value * 2000) so the bounds are bit-identical
            hi = np.where(ok & (v > 0), 1.0 / np.where(v > 0, v, 1.0) * 2000,
                          np.where(ok & (v < 0), 1.0 / np.where(u != 0, np.abs(u), 1.0) * 2000, default_hi))
            rec["distance"]["self"]["hi"] = hi
        return arr, bands

    def compile(self, ic, N=1, maxAV=None, max_distance=None):
        """All rows -> one ``CompiledModel`` with ``len(self)`` star models on the device (row i of the table = model i)."""
        arr, bands = self.build_structs(ic, N=N, maxAV=maxAV, max_distance=max_distance)
        n = len(arr)
        compiled = CompiledModel.from_struct_array(ic, arr, n, N, tuple(bands))
        return compiled
