"""CPU: the MIST bolometric-correction text-table loader (isochrones_b200/bcio.py).

No MIST download can exist here, so the test writes small tables in the MIST layout (five header lines, column names on
the sixth, `#` comments, one file per [Fe/H], several Rv per file) and checks the dense grid against (a) an independent
pandas parse following the reference's own recipe (bc.py:72-83 `parse_table`, mist/bc.py:161-163 the Rv = 3.1 slice,
DFInterpolator._make_grid) and (b) — in the build container, where the reference is importable — the reference's band
resolver `get_band` on every band name it knows.
"""
import os

import numpy as np
import pytest

from isochrones_b200 import bcio

REF = "/root/reference"


def _write_tables(tmp, phot, cols, fehs, seed):
    rng = np.random.RandomState(seed)
    teffs, loggs, avs, rvs = [3000.0, 4000.0, 5500.0, 8000.0], [-1.0, 2.0, 4.5], [0.0, 0.1, 0.5, 2.0], [2.0, 3.1, 4.1]
    truth = {}
    for feh in fehs:
        sign = "m" if feh < 0 else "p"
        path = os.path.join(tmp, "feh{0}{1:03.0f}.{2}".format(sign, abs(feh) * 100, phot))   # bc.py:66-70
        with open(path, "w") as f:
            f.write("# MIST version number  = 1.2\n# MESA revision number =     7503\n# photometric system    = %s\n"
                    "# ABUNDANCES: [Fe/H] = %.2f\n# number of filters = %d\n" % (phot, feh, len(cols)))
            f.write("#  " + "  ".join(["Teff", "logg", "[Fe/H]", "Av", "Rv"] + cols) + "\n")
            for t in teffs:
                for g in loggs:
                    if t == 8000.0 and g == -1.0:
                        continue                     # a hole in the table: the dense grid must hold NaN there
                    for av in avs:
                        for rv in rvs:
                            vals = rng.normal(size=len(cols)) + 0.001 * t / 1000 + rv
                            f.write("%10.1f %6.2f %6.2f %6.2f %4.1f " % (t, g, feh, av, rv) +
                                    " ".join("%12.6f" % v for v in vals) + "\n")
                            if rv == 3.1:
                                truth[(t, g, feh, av)] = [float("%12.6f" % v) for v in vals]
    return teffs, loggs, avs, truth


def test_dense_grid_from_text_tables(tmp_path):
    import pandas as pd

    tmp = str(tmp_path)
    fehs = [-1.0, 0.0, 0.5]
    ucols = ["Bessell_U", "Bessell_B", "Bessell_V", "2MASS_J", "2MASS_H", "2MASS_Ks", "Gaia_G_DR2Rev", "TESS"]
    wcols = ["WISE_W1", "WISE_W2", "WISE_W3", "WISE_W4"]
    teffs, loggs, avs, truth_u = _write_tables(tmp, "UBVRIplus", ucols, fehs, seed=1)
    _, _, _, truth_w = _write_tables(tmp, "WISE", wcols, fehs, seed=2)
    bands = ["V", "W2", "K", "G", "TESS", "J"]
    bc = bcio.load_mist_bc_grid(tmp, bands)
    assert bc["columns"] == bands and bc["kind"] == "bc"
    assert [list(a) for a in bc["axes"]] == [teffs, loggs, fehs, avs]
    assert bc["grid"].shape == (4, 3, 3, 4, len(bands))
    for (t, g, feh, av), vals in truth_u.items():
        i = (teffs.index(t), loggs.index(g), fehs.index(feh), avs.index(av))
        assert bc["grid"][i][0] == vals[ucols.index("Bessell_V")]
        assert bc["grid"][i][2] == vals[ucols.index("2MASS_Ks")]
        assert bc["grid"][i][1] == truth_w[(t, g, feh, av)][wcols.index("WISE_W2")]
    assert np.isnan(bc["grid"][3, 0]).all() and np.isfinite(bc["grid"][3, 1]).all()      # the hole, and only the hole

    # (a) the reference's recipe with pandas: read every table, concatenate, sort, slice Rv = 3.1, densify
    frames = []
    for f in sorted(os.listdir(tmp)):
        if f.endswith(".UBVRIplus"):
            path = os.path.join(tmp, f)
            with open(path) as fin:
                names = [line for i, line in enumerate(fin) if i == 5][0][1:].split()
            frames.append(pd.read_csv(path, names=names, sep=r"\s+", comment="#", index_col=list(bcio.INDEX_COLS)))
    df = pd.concat(frames).sort_index().xs(3.1, level="Rv")
    levels = [np.asarray(l, dtype=float) for l in df.index.levels]
    dense = np.full(tuple(len(l) for l in levels) + (len(df.columns),), np.nan)
    dense[tuple(np.asarray(c) for c in df.index.codes)] = df.values
    for a, b in zip(levels, bc["axes"]):
        assert np.array_equal(a, b)
    assert np.array_equal(dense[..., list(df.columns).index("Bessell_V")], bc["grid"][..., 0], equal_nan=True)
    assert np.array_equal(dense[..., list(df.columns).index("2MASS_J")], bc["grid"][..., 5], equal_nan=True)

    # the dict is what the interpolator mirror takes
    from isochrones_b200.interp import DFInterpolator
    it = DFInterpolator.from_arrays(bc["grid"], bc["axes"], bc["columns"])
    assert it.ndim == 4 and it.columns == bands

    with pytest.raises(FileNotFoundError):
        bcio.load_mist_bc_grid(tmp, ["SDSS_g"])
    with pytest.raises(ValueError):
        bcio.mist_band("no_such_band!")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference only exists in the build container")
def test_band_resolution_matches_reference():
    import importlib

    from oracle import ref_shim

    ref_shim.load()
    cls = importlib.import_module("isochrones.mist.bc").MISTBolometricCorrectionGrid
    names = list(bcio._SHORTCUTS) + [b for v in cls.phot_bands.values() for b in v]
    names += ["UK_J", "UKIRT_K", "PanSTARRS_g", "WISE_W1", "SkyMapper_u", "DECam_g", "SDSS_r"]
    assert set(bcio.PHOT_SYSTEMS) == set(cls.phot_bands)
    for b in names:
        assert bcio.mist_band(b) == cls.get_band(b), b
