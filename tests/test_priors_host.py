"""CPU: the reference's tests/test_priors.py, verbatim in structure, against the product's prior classes — each prior
integrates to 1 over its bounds and its sampler follows its pdf (host-side construction helpers; the device evaluation
of lnpdf / pdf is covered by tests/test_gpu_golden.py::test_priors)."""
import numpy as np


def test_age():
    from isochrones_b200.priors import AgePrior

    age_prior = AgePrior()
    age_prior.test_integral()
    age_prior.test_sampling()


def test_distance():
    from isochrones_b200.priors import DistancePrior

    distance_prior = DistancePrior()
    distance_prior.test_integral()
    distance_prior.test_sampling()


def test_AV():
    from isochrones_b200.priors import AVPrior

    AV_prior = AVPrior()
    AV_prior.test_integral()
    AV_prior.test_sampling()


def test_q():
    from isochrones_b200.priors import QPrior

    q_prior = QPrior()
    q_prior.test_integral()
    q_prior.test_sampling()


def test_salpeter():
    from isochrones_b200.priors import SalpeterPrior

    salpeter_prior = SalpeterPrior()
    salpeter_prior.test_integral()
    salpeter_prior.test_sampling()


def test_feh():
    from isochrones_b200.priors import FehPrior

    feh_prior = FehPrior()
    feh_prior.test_integral()
    feh_prior.test_sampling()
    feh_prior.bounds = (-3, 0.25)
    feh_prior.test_integral()
    feh_prior.test_sampling()
    # `feh_prior(-3.5) == 0` of the reference test is a device evaluation here: tests/test_gpu_golden.py::test_priors
    assert feh_prior._host_call(-3.5) == 0 and feh_prior._host_call(0.4) == 0


def test_chabrier():
    from isochrones_b200.priors import ChabrierPrior

    chabrier_prior = ChabrierPrior()
    chabrier_prior.test_integral()
    chabrier_prior.test_sampling()
