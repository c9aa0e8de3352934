"""GPU: real-data path end to end — a fabricated ``$ISOCHRONES`` tree in the reference's on-disk
layout (``full_grid*.npz`` dense-grid cache + axes sidecar, MIST BC text tables) is loaded through
``get_ichrone("mist", root=...)`` and must reproduce the UNMODIFIED reference's lnprior / lnlike / lnpost
(tests/golden) on every golden star model, exactly as the in-memory path does."""
import numpy as np
import pytest

from tests.helpers import (assert_same_special, golden_grids, load_specs, product_model_from_spec, write_isochrones_tree)

pytestmark = pytest.mark.gpu


def test_get_ichrone_from_isochrones_tree(golden, tmp_path):
    import isochrones_b200 as ib
    from isochrones_b200 import mistio

    gi, gl = golden["interp"], golden["lnpost"]
    trk, iso, bc = golden_grids(gi)
    root = str(tmp_path)
    write_isochrones_tree(root, "track", trk, bc)
    write_isochrones_tree(root, "iso", iso, bc)
    with pytest.raises(mistio.MistDataNotFound):
        ib.get_ichrone("mist", root=str(tmp_path / "nothing_here"))
    ics = {"track": ib.get_ichrone("mist", tracks=True, bands=bc["columns"], root=root, limits={"eep": (0, 60)}),
           "iso": ib.get_ichrone("mist", bands=bc["columns"], root=root, limits={"eep": (0, 60)})}
    assert ics["track"].name == "mist" and "full_grid_v1.2_vvcrit0.4.npz" in ics["track"].model_grid.source
    for kind, model in (("track", trk), ("iso", iso)):
        it = ics[kind].model_grid.interp
        assert np.array_equal(it.grid, model["grid"], equal_nan=True) and it.columns == model["columns"]
        assert np.array_equal(ics[kind].bc_grid.interp.grid, bc["grid"])
    n = 0
    for name, spec in load_specs(gl).items():
        mod = product_model_from_spec(spec, ics[spec["kind"]])
        pars = gl["lp_%s_pars" % name]
        lnpost, lnprior, lnlike = mod.lnpost_batch(pars, parts=True)
        for got, key in ((lnpost, "lnpost"), (lnprior, "lnprior"), (lnlike, "lnlike")):
            want = gl["lp_%s_%s" % (name, key)]
            assert_same_special(got, want)
            m = np.isfinite(want)
            assert np.allclose(got[m], want[m], rtol=1e-12, atol=1e-9), (name, key)
        n += 1
    assert n >= 9
