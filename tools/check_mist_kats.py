#!/usr/bin/env python
"""MIST-valued known answers of the reference, checked against the product on REAL MIST data.

    ISOCHRONES=/path/to/.isochrones python tools/check_mist_kats.py [--root DIR] [--rtol 1e-6]

The build environment has no MIST data (no network), so these checks cannot run there; this script is for anyone who
holds the reference's ``$ISOCHRONES`` tree (the ``full_grid*.npz`` caches it writes + the MIST bolometric-correction
tables).  Every number below is an output printed in the reference's own tests / notebooks:

  K1  tests/test_basic.py:16-18            MIST_Isochrone.logg at three points
  K2  docs/grid_interpolator.ipynb [3,5]   isochrone grid: interp_value, interp_mag
  K3  docs/grid_interpolator.ipynb [7,8]   track grid: interp_value, interp_mag
  K4  docs/grid_interpolator.ipynb [18,22,26]  get_eep (interp_eep) = 343.8; the accurate EEP puts the age back
  K5  docs/modelgrids.ipynb:820            track_grid.interp([-0.12, 1.01, 353.1], [mass, radius, logg, Teff])
  K6  docs/bc.ipynb:205                    bc_grid.interp([5770, 4.44, 0.0, 0.], ['G', 'K'])
  K7  docs/starmodel.ipynb [2-8]           SingleStarModel lnprior / lnlike / lnpost at two points
  K8  docs/multiple.ipynb [3-9]            BinaryStarModel lnpost = -645802.2025506602 (needs generate(accurate=True) inputs)

Interpolated properties must agree to ``--rtol`` (1e-6, BASELINE.json), lnpost values to 1e-4 absolute.  K7 / K8 go
through an EEP the reference finds with a scipy Nelder-Mead search (tolerance ~1e-4 in EEP) where the product brackets the
root; they are compared at 2e-2 absolute and the difference is printed.
"""
import argparse
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

RESULTS = []


def check(name, got, want, rtol=1e-6, atol=0.0):
    got, want = np.atleast_1d(np.asarray(got, dtype=float)), np.atleast_1d(np.asarray(want, dtype=float))
    ok = got.shape == want.shape and bool(np.allclose(got, want, rtol=rtol, atol=atol))
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))) if got.shape == want.shape else np.inf
    RESULTS.append((name, ok, err))
    print("%-58s %s  max rel diff %.3g" % (name, "ok  " if ok else "FAIL", err))
    if not ok:
        print("      got  %r\n      want %r" % (got.tolist(), want.tolist()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=None, help="the reference's $ISOCHRONES directory (default: $ISOCHRONES or ~/.isochrones)")
    ap.add_argument("--rtol", type=float, default=1e-6)
    args = ap.parse_args()
    import isochrones_b200 as ib
    from isochrones_b200 import mistio

    bands = ["J", "H", "K", "G", "BP", "RP", "W1", "W2", "W3", "TESS", "Kepler"]
    try:
        iso = ib.get_ichrone("mist", bands=bands, root=args.root)
        trk = ib.get_ichrone("mist", bands=bands, tracks=True, root=args.root)
    except mistio.MistDataNotFound as e:
        print("no MIST data: %s" % e)
        return 2
    print("isochrone grid: %s\ntrack grid:     %s\n" % (iso.model_grid.source, trk.model_grid.source))
    r = args.rtol

    # K1
    check("K1 logg(632, 7.55, -1.75)", iso.logg(632, 7.55, -1.75), 2.4117770214014103, r)
    check("K1 logg(355, 9.653, 0.0)", iso.logg(355, 9.653, 0.0), 4.4124675, 1e-7)
    check("K1 logg(700, 9.3, -0.03)", iso.logg(700, 9.3, -0.03), 2.24831956, 1e-7)
    # K2
    pars = [353, 9.78, -1.24]
    check("K2 iso interp_value(mass, radius, Teff)", iso.interp_value(pars, ["mass", "radius", "Teff"]),
          [7.93829519e-01, 7.91444054e-01, 6.30305932e+03], 1e-8)
    teff, logg, feh, mags = iso.interp_mag(pars + [200, 0.11], ["K", "BP", "RP"])
    check("K2 iso interp_mag Teff, logg, feh", [teff, logg, feh], [6303.059322477636, 4.540738764316164, -1.377262817643937], r)
    check("K2 iso interp_mag K, BP, RP", mags, [10.25117074, 11.73997159, 11.06529993], 1e-8)
    # K3
    pars = [0.794, 353, -1.24]
    check("K3 track interp_value(mass, radius, Teff, age)", trk.interp_value(pars, ["mass", "radius", "Teff", "age"]),
          [7.93843749e-01, 7.91818696e-01, 6.31006708e+03, 9.77929505e+00], 1e-8)
    teff, logg, feh, mags = trk.interp_mag(pars + [200, 0.11], ["K", "BP", "RP"])
    check("K3 track interp_mag Teff, logg, feh", [teff, logg, feh], [6310.067080800683, 4.54076772643659, -1.372925841944066], r)
    check("K3 track interp_mag K, BP, RP", mags, [10.24893319, 11.73358578, 11.06056746], 1e-8)
    check("K3 track mass(*pars)", trk.mass(*pars), 0.79384375, 1e-8)
    # K4
    check("K4 get_eep(1.01, 9.51, 0.01) [interp_eep]", trk.get_eep(1.01, 9.51, 0.01), 343.8, 1e-9)
    e = trk.get_eep(1.01, 9.51, 0.01, accurate=True)
    check("K4 accurate EEP (reference: Nelder-Mead, 343.19635)", e, 343.1963539123535, 3e-5)
    check("K4 age at the accurate EEP", trk.interp_value([1.01, e, 0.01], ["age"]), [9.51], 1e-7)
    # K5
    check("K5 track_grid.interp([-0.12, 1.01, 353.1])", trk.model_grid.interp([-0.12, 1.01, 353.1], ["mass", "radius", "logg", "Teff"]),
          [1.00983180, 1.04691913, 4.40266419, 6033.83320], 1e-8)
    # K6
    check("K6 bc_grid.interp([5770, 4.44, 0.0, 0.], [G, K])", iso.bc_grid.interp([5770.0, 4.44, 0.0, 0.0], ["G", "K"]),
          [0.0819599, 1.45398088], 1e-6)
    # K7: the star of docs/starmodel.ipynb, observations as printed by the reference's track.generate(1.0, 9.74, -0.05, 100, 0.02)
    true_props = {"Teff": 5934.703385987951, "logg": 4.370219109480715, "feh": -0.09685557997282962,
                  "J": 8.435233804866742, "H": 8.124109062114325, "K": 8.09085566863133}
    gen = trk.generate(1.0, 9.74, -0.05, distance=100, AV=0.02, return_dict=True, accurate=True)
    check("K7 generate(1.0, 9.74, -0.05, 100, 0.02): Teff logg feh J H K",
          [gen["Teff"], gen["logg"], gen["feh"], gen["J_mag"], gen["H_mag"], gen["K_mag"]],
          [true_props[k] for k in ("Teff", "logg", "feh", "J", "H", "K")], 1e-5)
    uncs = dict(Teff=80, logg=0.1, feh=0.1)
    props = {p: (true_props[p], uncs[p]) for p in ("Teff", "logg", "feh")}
    props.update({b: (true_props[b], 0.02) for b in "JHK"})
    props["parallax"] = (10.0, 0.1)
    mod = ib.SingleStarModel(iso, name="demo", **props)
    eep = iso.get_eep(1.0, 9.74, -0.05, accurate=True)
    p1 = [eep, 9.74, -0.05, 100, 0.02]
    check("K7 (lnprior, lnlike, lnpost) at the truth", [mod.lnprior(p1), mod.lnlike(p1), mod.lnpost(p1)],
          [-23.05503287088296, -20.716150242083508, -43.77118311296647], 0.0, 2e-2)
    p2 = [eep + 3, 9.74 - 0.05, -0.05 + 0.02, 100, 0.02]
    check("K7 (lnprior, lnlike, lnpost) at the shifted point", [mod.lnprior(p2), mod.lnlike(p2), mod.lnpost(p2)],
          [-23.251706955307853, -85.08590699022739, -108.33761394553524], 0.0, 5e-2)
    # K8: docs/multiple.ipynb — magnitudes of a 1.0 + 0.5 Msun pair at 500 pc, AV 0.2, from generate(accurate=True)
    bands6 = ["J", "H", "K", "BP", "RP", "G"]
    a = trk.generate(1.0, 9.6, 0.0, distance=500, AV=0.2, bands=bands6, return_dict=True, accurate=True)
    b = trk.generate(0.5, 9.6, 0.0, distance=500, AV=0.2, bands=bands6, return_dict=True, accurate=True)
    unc = dict(J=0.02, H=0.02, K=0.02, BP=0.002, RP=0.002, G=0.001)
    tot = {k: (-2.5 * np.log10(10 ** (-0.4 * a[k + "_mag"]) + 10 ** (-0.4 * b[k + "_mag"])), unc[k]) for k in bands6}
    mb = ib.BinaryStarModel(iso, parallax=(2, 0.05), name="demo_binary", **tot)
    mb.set_bounds(eep=(1, 600), age=(8, 10))
    check("K8 BinaryStarModel lnpost([350, 300, 9.7, 0.0, 300, 0.1])", mb.lnpost([350.0, 300.0, 9.7, 0.0, 300.0, 0.1]),
          -645802.2025506602, 0.0, 5.0)      # |lnpost| ~ 6e5 and sigma_G = 1 mmag: 1e-5 mag in a generated magnitude moves it by ~1
    from isochrones_b200.priors import GaussianPrior

    mb.set_prior(age=GaussianPrior(9.6, 1, bounds=(8, 10)))
    check("K8 ... with GaussianPrior(9.6, 1) on age: difference", mb.lnpost([350.0, 300.0, 9.7, 0.0, 300.0, 0.1]) -
          (-645802.7700077017), 0.0, 0.0, 5.0)
    bad = [n for n, ok, _ in RESULTS if not ok]
    print("\n%d of %d known answers reproduced%s" % (len(RESULTS) - len(bad), len(RESULTS), "" if not bad else "; FAILED: " + "; ".join(bad)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
