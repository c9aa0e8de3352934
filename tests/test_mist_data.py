"""CPU: real-data ingestion (isochrones_b200/mistio.py) — the reference's dense-grid cache
``full_grid*.npz`` + axes, the MIST BC text tables, and ``get_ichrone`` refusing to fabricate data.

No MIST download can exist here, so the ``$ISOCHRONES`` tree is fabricated in the reference's layout from small
synthetic MIST-shaped grids; in the build container the cache file is additionally written by the UNMODIFIED
reference's own ``DFInterpolator`` (interp.py:590-614) and read back by the product."""
import os

import numpy as np
import pytest

from isochrones_b200 import mistio, synthetic as syn
from tests.helpers import write_bc_text_tables, write_isochrones_tree

REF = "/root/reference"


def test_paths_follow_the_reference(tmp_path, monkeypatch):
    monkeypatch.setenv("ISOCHRONES", str(tmp_path))
    d, npz, axes = mistio.model_grid_files("track")
    assert npz == os.path.join(str(tmp_path), "mist", "tracks", "full_grid_v1.2_vvcrit0.4.npz")     # models.py:163-165
    d, npz, axes = mistio.model_grid_files("iso")
    assert npz == os.path.join(str(tmp_path), "mist", "full_grid_v1.2_vvcrit0.4_full_isos.npz")
    monkeypatch.delenv("ISOCHRONES")
    assert mistio.isochrones_root() == os.path.expanduser("~/.isochrones")                          # config.py:5


@pytest.mark.parametrize("kind", ["track", "iso"])
def test_axes_derived_from_the_grid(tmp_path, kind):
    """The reference's cache holds no axes; with the 15 MIST [Fe/H] values they are recovered from the grid's own
    index-valued columns (initial_mass / age, eep)."""
    model = syn.make_track_grid(n_mass=14, n_eep=57) if kind == "track" else syn.make_iso_grid(n_age=12, n_eep=57)
    if kind == "iso":   # keep the ages that have at least one populated node (the derivation needs one per slice)
        keep = ~np.isnan(model["grid"][..., 0]).all(axis=(1, 2))
        model = dict(model, grid=model["grid"][keep], axes=(model["axes"][0][keep],) + model["axes"][1:])
    mistio.save_model_grid(model, kind, root=str(tmp_path), sidecar=False)
    got = mistio.load_model_grid(kind, root=str(tmp_path))
    assert "derived" in got["source"]
    for a, b in zip(got["axes"], model["axes"]):
        assert np.array_equal(a, b)
    assert np.array_equal(got["grid"], model["grid"], equal_nan=True) and got["columns"] == model["columns"]
    assert got["limits"] == mistio.MIST_LIMITS
    # with a sidecar the axes are read, not derived (non-MIST [Fe/H] lists need it)
    small = syn.make_track_grid(n_feh=5, n_mass=9, n_eep=40)
    root2 = str(tmp_path / "two")
    mistio.save_model_grid(small, "track", root=root2, sidecar=False)
    with pytest.raises(ValueError):
        mistio.load_model_grid("track", root=root2)
    mistio.save_model_grid(small, "track", root=root2, sidecar=True)
    got = mistio.load_model_grid("track", root=root2)
    assert "sidecar" in got["source"] and all(np.array_equal(a, b) for a, b in zip(got["axes"], small["axes"]))


def test_missing_data_raises(tmp_path):
    import isochrones_b200 as ib

    with pytest.raises(mistio.MistDataNotFound):
        mistio.load_model_grid("iso", root=str(tmp_path))
    with pytest.raises(mistio.MistDataNotFound):
        mistio.load_bc_grid(["J"], root=str(tmp_path))
    with pytest.raises(FileNotFoundError):          # the drop-in entry point does not fabricate physics
        ib.get_ichrone("mist", root=str(tmp_path))
    with pytest.raises(FileNotFoundError):
        ib.get_ichrone("mist", tracks=True, root=str(tmp_path))


def test_bc_tables_and_cache(tmp_path):
    bc = syn.make_bc_grid(bands=("J", "K", "W1", "G"), n_teff=6, n_logg=4, n_feh=3, n_av=4)
    root = str(tmp_path)
    write_bc_text_tables(os.path.join(root, "BC", "mist"), bc)
    got = mistio.load_bc_grid(["K", "W1", "J"], root=root)
    assert got["columns"] == ["K", "W1", "J"]
    for a, b in zip(got["axes"], bc["axes"]):
        assert np.array_equal(a, b)
    assert np.array_equal(got["grid"][..., 0], bc["grid"][..., 1]) and np.array_equal(got["grid"][..., 1], bc["grid"][..., 2])
    assert os.path.exists(mistio.bc_cache_file(["K", "W1", "J"], root))
    again = mistio.load_bc_grid(["K", "W1", "J"], root=root)        # second call: from the dense cache
    assert again["source"].endswith(".npz") and np.array_equal(again["grid"], got["grid"])
    with pytest.raises(mistio.MistDataNotFound):
        mistio.load_bc_grid(["SDSS_g"], root=root, cache=False)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference only exists in the build container")
def test_cache_written_by_the_reference(tmp_path):
    """The UNMODIFIED reference's DFInterpolator writes the dense-grid cache from a MultiIndex frame (ragged: NaN-tail
    rows dropped, as the real MIST frame is); the product reads that file and recovers the index levels."""
    import importlib

    import pandas as pd

    from oracle import ref_shim

    ref_shim.load()
    ref_interp = importlib.import_module("isochrones.interp")
    model = syn.make_track_grid(n_mass=10, n_eep=33)
    fehs, masses, eeps = model["axes"]
    idx = pd.MultiIndex.from_product([fehs, masses, eeps], names=mistio.INDEX_NAMES["track"])
    df = pd.DataFrame(model["grid"].reshape(-1, len(model["columns"])), index=idx, columns=model["columns"])
    df = df[~df["Teff"].isna()]
    d, npz, _ = mistio.model_grid_files("track", root=str(tmp_path))
    os.makedirs(d)
    ref = ref_interp.DFInterpolator(df, filename=npz, is_full=False)
    assert os.path.exists(npz)
    got = mistio.load_model_grid("track", root=str(tmp_path))
    assert got["columns"] == list(ref.columns)
    assert np.array_equal(got["grid"], ref.grid, equal_nan=True)
    for a, b in zip(got["axes"], ref.index_columns):
        assert np.array_equal(a, b)
