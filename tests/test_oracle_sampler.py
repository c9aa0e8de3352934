"""CPU: the C oracle's stretch-move driver (oracle/iso_oracle.c::orc_stretch_move — the checker of the device samplers
and the CPU baseline of BASELINE configs 3 / 4) against an independent numpy restatement of the same published
algorithm and random stream (tests/helpers.py::replay_stretch_move).  Both use the oracle's lnpost, so chains must
coincide bit for bit, single- and multi-threaded, one chain or several."""
import numpy as np

from tests.helpers import golden_grids, load_specs, ns_model_from_spec, replay_stretch_move


def _model(golden, name):
    from oracle import oracle

    gi, gl = golden["interp"], golden["lnpost"]
    trk, iso, bc = golden_grids(gi)
    spec = load_specs(gl)[name]
    om = oracle.StarModel(ns_model_from_spec(spec, trk if spec["kind"] == "track" else iso, bc))
    pars = gl["lp_%s_pars" % name]
    good = pars[np.isfinite(gl["lp_%s_lnpost" % name])]
    return om, good


def test_c_driver_matches_numpy_replay(golden):
    from oracle import oracle

    for name in ("track_full", "iso_binary"):
        om, good = _model(golden, name)
        nw, steps, seed = 40, 25, 4242
        p0 = good[:nw]
        want_chain, want_lp, want_acc = replay_stretch_move(lambda r: om.lnpost_batch(np.ascontiguousarray(r)), p0, steps, seed)
        for threads in (1, 4):
            chain, lnp, pos, lp, acc = oracle.stretch_move(om, p0, steps, seed, n_threads=threads)
            assert np.array_equal(chain[:, 0], want_chain) and np.array_equal(lnp[:, 0], want_lp), (name, threads)
            assert acc[0] == want_acc and np.array_equal(pos[0], want_chain[-1])
        assert 0 < want_acc < steps * nw


def test_chains_are_independent_streams(golden):
    from oracle import oracle

    om, good = _model(golden, "iso_single")
    nw, steps, seed = 24, 12, 7
    p0 = np.stack([good[c * nw:(c + 1) * nw] for c in range(3)])
    chain, lnp, pos, lp, acc = oracle.stretch_move(om, p0, steps, seed, n_threads=3)
    for c in range(3):
        want_chain, want_lp, want_acc = replay_stretch_move(lambda r: om.lnpost_batch(np.ascontiguousarray(r)), p0[c], steps,
                                                            seed, chain=c)
        assert np.array_equal(chain[:, c], want_chain) and acc[c] == want_acc
    # continuing a run (step0) equals one longer run
    a1 = oracle.stretch_move(om, p0[0], 5, seed)
    a2 = oracle.stretch_move(om, a1[2][0], 7, seed, step0=5)
    assert np.array_equal(a2[2][0], chain[-1, 0])
