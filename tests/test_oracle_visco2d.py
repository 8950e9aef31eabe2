"""CPU tests of the 2-D viscoelastic oracle (oracle/cpml_oracle_visco2d.c): the independent numpy
restatement, the elastic limit, the optional energy, and the golden vectors."""
import os

import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from oracle.np_restatement_visco2d import run_2d_visco_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIELDS = ("vx", "vy", "sigmaxx", "sigmayy", "sigmaxy", "e1", "e11", "e13")


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("material", ["homogeneous", "layered"])
def test_numpy_restatement_matches_c_oracle_bit_for_bit(order, material):
    c = refcfg.cfgv2d(order=order, material=material, nstep=150)
    r = O.run_2d_visco(**c, want_fields=True)
    n = run_2d_visco_np(**c)
    assert np.abs(r["sisvy"]).max() > 0 and np.abs(r["e13"]).max() > 0
    for f in FIELDS:
        assert np.array_equal(r[f], n[f]), f
    for s in ("sisvx", "sisvy", "sispressure"):
        assert np.array_equal(r[s], n[s]), s


def test_elastic_branch_equals_unit_relaxation_times():
    """VISCOELASTIC_ATTENUATION = .false. (2D-visco-4th :713-760) is the viscoelastic branch with
    tau_epsilon == tau_sigma (phi = 0, memory variables stay zero): the library relies on this."""
    c = refcfg.cfgv2d(nstep=120, material="layered")
    e = O.run_2d_visco(**c, viscoelastic_attenuation=False, want_fields=True)
    ones = dict(tau_epsilon_nu1=(1.0,) * 3, tau_sigma_nu1=(1.0,) * 3, tau_epsilon_nu2=(1.0,) * 3, tau_sigma_nu2=(1.0,) * 3)
    u = O.run_2d_visco(**{**c, **ones}, want_fields=True)
    for f in ("vx", "vy", "sigmaxx", "sigmayy", "sigmaxy"):
        assert np.array_equal(e[f], u[f]), f
    assert not np.any(u["e1"]) and not np.any(u["e13"])
    v = O.run_2d_visco(**c)
    assert 0.01 < refcfg.rel_l2(v["sisvy"], e["sisvy"]) < 1.0       # attenuation changes the answer


def test_energy_only_when_asked():
    c = refcfg.cfgv2d(nstep=80)
    off = O.run_2d_visco(**c)
    on = O.run_2d_visco(**c, compute_energy=True)
    assert not off["energy_kinetic"].any() and not off["energy_potential"].any()        # COMPUTE_ENERGY = .false., :201
    assert on["energy_kinetic"][-1] > 0 and on["energy_potential"][-1] > 0
    assert np.array_equal(off["sisvx"], on["sisvx"])


@pytest.mark.parametrize("order", [2, 4])
def test_golden_visco2d(order):
    g = np.load(os.path.join(GOLD, f"cpml2d_visco_order{order}.npz"))
    o = O.run_2d_visco(**refcfg.cfgv2d(order=order, material="layered", nstep=300))
    for s in ("sisvx", "sisvy", "sispressure"):
        assert np.array_equal(o[s], g[s]), s
