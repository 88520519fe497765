"""GPU: the CUDA stepper reproduces the analytic answers the reference's examples compare against
-- the cases of tests/analytic_cases.py, whose checks the CPU suite runs on the oracle."""

import pytest

import analytic_cases as ac

pytestmark = pytest.mark.gpu


def run(case, **kw):
    from vivsim_b200 import Stepper
    spec, f0, steps, check = case
    st = Stepper(spec, **kw).set_f(f0)
    st.step(steps)
    check(st.get_f().detach().cpu().numpy())


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
