"""CPU: the oracle reproduces the analytic answers the reference's examples compare against (SURVEY.md section 4:
Taylor-Green decay, Couette profile, the four Poiseuille recipes, ABC flow).  The reference ships no tests; these
known-answer cases pin the physics of the restated algorithm next to the fixtures that pin its arithmetic.
tests/test_gpu_z_analytic.py runs the same cases through the CUDA stepper."""

import pytest

import analytic_cases as ac
from oracle import recipes


def run(case):
    spec, f0, steps, check = case
    f, _ = recipes.run(spec, f0, steps)
    check(f)


@pytest.mark.parametrize("prepared", [False, True])
def test_taylor_green_vortex_decay(prepared):
    run(ac.taylor_green(prepared))


def test_couette_profile():
    run(ac.couette())


@pytest.mark.parametrize("kind", ac.POISEUILLE_KINDS)
def test_poiseuille_profile(kind):
    run(ac.poiseuille(kind))


def test_abc_flow_decay():
    run(ac.abc_flow())
