"""CPU: host-side helpers of the prior classes — every named prior of the star models integrates to one over its bounds
and its sampler follows its density (the checks the reference makes in its tests/test_priors.py).  The device
evaluation of lnpdf / pdf is covered by tests/test_gpu_golden.py::test_priors and ::test_feh_prior_rebounded."""
import numpy as np
import pytest

from isochrones_b200 import priors as P

CASES = {
    "age": lambda: P.AgePrior(),
    "distance": lambda: P.DistancePrior(),
    "AV": lambda: P.AVPrior(),
    "q": lambda: P.QPrior(),
    "salpeter": lambda: P.SalpeterPrior(),
    "feh": lambda: P.FehPrior(),
    "chabrier": lambda: P.ChabrierPrior(),
    "gauss_truncated": lambda: P.GaussianPrior(9.6, 0.4, bounds=(8.0, 10.0)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_unit_integral_and_sampling(name):
    np.random.seed(7)
    prior = CASES[name]()
    prior.test_integral()
    prior.test_sampling()
    draws = prior.sample(1000)
    lo, hi = prior.bounds
    assert draws.shape == (1000,) and np.all(draws >= lo) and np.all(draws <= hi)


def test_feh_rebounded():
    np.random.seed(8)
    prior = P.FehPrior()
    prior.bounds = (-3, 0.25)
    assert 0.8 < prior._norm < 1          # the mass of the unbounded density inside the new bounds
    prior.test_integral()
    prior.test_sampling()
    assert prior.density(-3.5) == 0 and prior.density(0.4) == 0 and prior.density(0.0) > 0


def test_rebounding_a_bounded_prior_must_keep_unit_mass():
    with pytest.raises(ValueError):
        P.GaussianPrior(0.0, 1.0).bounds = (-0.5, 0.5)      # a BoundedPrior is not re-normalised by new bounds
    ok = P.AVPrior()
    ok.bounds = (0, 1.0)


def test_lognormal_and_eep_sampling_shapes():
    np.random.seed(9)
    ln = P.LogNormalPrior(np.log(0.3), 0.5)
    x = ln.sample(20000)
    assert abs(np.median(x) - 0.3) < 0.01 and x.min() > 0
