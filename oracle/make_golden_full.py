"""TEST INFRASTRUCTURE — golden vectors from the UNMODIFIED reference on the FULL-SIZE benchmark grids.

    python oracle/make_golden_full.py          (build container only: needs /root/reference)

The full MIST-shaped grids (322 MB pack) are too large to commit, but `isochrones_b200.synthetic` regenerates them
deterministically; the file stores SHA-256 digests of the grids so a test can prove it rebuilt the same bits, the
seeded parameter rows, and the reference's own lnprior / lnlike / lnpost (BasicStarModel, scalar calls through its
numba kernels) for the bench's single-star track model, an isochrone single and an isochrone binary.
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from isochrones_b200 import synthetic as syn  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import fix_ub  # noqa: E402

ISO_COLS = ("Teff", "logg", "feh", "Mbol", "mass", "dm_deep", "nu_max", "delta_nu")
BANDS = ("V", "J", "H", "K")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rows_for(kind, mod, truth, axes, n, N):
    bounds = [tuple(float(v) for v in mod.bounds(p)) for p in mod.param_names]
    p = np.concatenate([syn.posterior_like_batch(kind, n, truth, seed=61),
                        syn.prior_like_batch(kind, n // 2, bounds, seed=62),
                        syn.edge_batch(kind, n // 4, truth, axes, bounds, seed=63)])
    if kind == "track":
        p = np.concatenate([p, bench.scattered_batch(n // 2, seed=64)])
    # rows whose corner reads would leave the reference's array (undefined behaviour there: a segfault on a grid this
    # size) are moved into the first cell of the leading axis, exactly as oracle/make_golden.py does
    lead = 2 if kind == "track" else N
    for _ in range(2):
        for k in range(N):
            cols = [2, 0, 1] if kind == "track" else [N, N + 1, k]
            coords = p[:, cols].copy()
            fix_ub(axes, coords)
            p[:, lead] = coords[:, 0]
    return p


def main():
    ref = ref_shim.load()
    out = {}
    trk = syn.make_track_grid(columns=bench.PACK_COLUMNS)
    iso = syn.make_iso_grid(columns=ISO_COLS)
    bc = syn.make_bc_grid(bands=BANDS)
    meta = {"track_sha256": digest(trk["grid"]), "iso_sha256": digest(iso["grid"]), "bc_sha256": digest(bc["grid"]),
            "track_columns": list(bench.PACK_COLUMNS), "iso_columns": list(ISO_COLS), "bands": list(BANDS), "cases": {}}
    for name, kind, model, N in (("track_single", "track", trk, 1), ("iso_single", "iso", iso, 1), ("iso_binary", "iso", iso, 2)):
        ic = ref_shim.make_ref_ic(kind, model, bc, eep_bounds=(0, 1710))
        truth = syn.default_truth(kind, n_stars=N)
        prim = list(truth) if kind == "track" else [truth[0]] + list(truth[N:])
        _, _, _, mags = ic.interp_mag(prim, list(BANDS))
        obs = {b: (float(np.round(m, 3)) - (0.35 if N > 1 else 0.0), 0.02) for b, m in zip(BANDS, mags)}
        kw = dict(Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1))
        mod = ref.starmodel.BasicStarModel(ic, N=N, **kw, **obs)
        p = rows_for(kind, mod, truth, model["axes"], 600, N)
        lnprior, lnlike, lnpost = np.empty(len(p)), np.empty(len(p)), np.empty(len(p))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i, row in enumerate(p):
                lnprior[i], lnlike[i], lnpost[i] = mod.lnprior(row), mod.lnlike(row), mod.lnpost(row)
        out[name + "_pars"], out[name + "_lnprior"], out[name + "_lnlike"], out[name + "_lnpost"] = p, lnprior, lnlike, lnpost
        meta["cases"][name] = {"kind": kind, "N": N, "obs": {k: list(v) for k, v in {**kw, **obs}.items()},
                               "n_finite": int(np.isfinite(lnpost).sum()), "n_nan": int(np.isnan(lnpost).sum())}
        print(name, len(p), "rows; finite", meta["cases"][name]["n_finite"], "nan", meta["cases"][name]["n_nan"])
    out["meta_json"] = np.array(json.dumps(meta))
    path = os.path.join(ROOT, "tests", "golden", "golden_full.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
