"""Thermostat oracle pinned on the reference: known answers of test/test_thermostat.py:117-143 and
outputs of the real hymd/thermostat.py (tests/golden/thermostat_golden.npz)."""
import os

import numpy as np
import pytest

from oracle import thermostat_oracle as to

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "thermostat_golden.npz"))


def groups_of(names, groups):
    """thermostat.py:177-183: a particle belongs to group i if its name is in groups[i]; an empty
    configuration means one group of everything."""
    names = np.asarray(names)
    if not groups:
        return np.zeros(len(names), dtype=np.int32), 1
    out = np.full(len(names), -1, dtype=np.int32)
    for i, g in enumerate(groups):
        for t in g:
            out[names == np.bytes_(t)] = i
    return out, len(groups)


CASES = {
    "all": [],
    "abcd": [["A"], ["B"], ["C"], ["D"]],
    "abc_d": [["A", "B", "C"], ["D"]],
}


@pytest.mark.parametrize("remove", [False, True], ids=["nocom", "com"])
@pytest.mark.parametrize("case", list(CASES))
def test_csvr_matches_reference(case, remove):
    pre = f"mws/{case}_{'com' if remove else 'nocom'}"
    mass, gas, T0, dt, inner, tau = G["mws/params"]
    v = G["mws/velocities"].copy()
    grp, ng = groups_of(G["mws/names"], CASES[case])
    work = to.csvr_thermostat(v, grp, ng, mass=mass, gas_constant=gas, target_temperature=T0,
                              time_step=dt, respa_inner=int(inner), tau=tau,
                              draws=list(zip(G[pre + "/gauss"], G[pre + "/chi2"])),
                              remove_center_of_mass_momentum=remove)
    np.testing.assert_allclose(v, G[pre + "/v"], rtol=1e-13, atol=1e-15)
    assert work == pytest.approx(float(G[pre + "/work"]), abs=1e-11)
    if not remove:
        K = 0.5 * mass * np.sum(v ** 2)
        if case == "all":      # test_thermostat.py:117-118
            assert K == pytest.approx(171.25339969021243, abs=1e-11)
            assert work == pytest.approx(2.797848031552252, abs=1e-11)
        if case == "abcd":     # test_thermostat.py:143
            assert work == pytest.approx(-21.670791766960217, abs=1e-11)


def test_kinetic_energy_regression_of_the_fixture():
    """test_thermostat.py:74-78."""
    v = G["mws/velocities"]
    assert 0.5 * 72.0 * np.sum(v ** 2) == pytest.approx(168.45555165866017, abs=1e-12)


def test_larger_system_and_cancel_com():
    mass, gas, T0, dt, inner, tau = G["rand/params"]
    v = G["rand/v0"].copy()
    grp, ng = groups_of(G["rand/names"], [["A", "B"], ["W"]])
    work = to.csvr_thermostat(v, grp, ng, mass=mass, gas_constant=gas, target_temperature=T0,
                              time_step=dt, respa_inner=int(inner), tau=tau,
                              draws=list(zip(G["rand/gauss"], G["rand/chi2"])))
    np.testing.assert_allclose(v, G["rand/v"], rtol=1e-12, atol=1e-14)
    assert work == pytest.approx(float(G["rand/work"]), rel=1e-11)
    out = to.cancel_com_momentum(G["mws/velocities"].copy(), len(G["mws/velocities"]))
    np.testing.assert_allclose(out, G["mws/cancel_com"], rtol=1e-14, atol=1e-16)
    assert np.abs(out.sum(axis=0)).max() < 1e-13
