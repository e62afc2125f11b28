"""
Pins the multilayer oracle (``oracle/multilayer.py``) to the reference's golden
vectors: the IMD tables of ``optika/materials/_tests/test_multilayers.py:178-287``
(committed as ``tests/golden/imd_*.npz`` by ``tests/golden/make_golden.py``) at the
reference's own tolerance ``rtol=1e-4``, plus its identities.
"""

import pathlib

import numpy as np
import pytest

from oracle import multilayer as orm

GOLDEN = pathlib.Path(__file__).parent / "golden"
NK = pathlib.Path(__file__).parent.parent / "optika_b200" / "data" / "nk"


def n_of(chemical, wavelength_angstrom):
    return orm.interp_nk(wavelength_angstrom, orm.load_nk(NK / f"{chemical}.nk"))


# (file, layers [(chemical, thickness A)], substrate, is_mirror): test_multilayers.py:188-222
CASES = [
    ("Si", [], "Si", False),
    ("SiO2", [("SiO2", 50.0)], "Si", False),
    ("SiO2_100A", [("SiO2", 100.0)], "Si", False),
    ("SiC_Cr", [("SiC", 250.0), ("Cr", 50.0)], "SiO2", True),
]


@pytest.mark.parametrize("name,layers,substrate,is_mirror", CASES)
def test_multilayer_efficiency_vs_imd_file(name, layers, substrate, is_mirror):
    g = np.load(GOLDEN / f"imd_{name}.npz")
    w = g["wavelength_angstrom"]
    efficiency_file = g["columns"][0]
    stack = [(n_of(c, w), t, 0, 0.0) for c, t in layers]
    r_s, r_p, t_s, t_p = orm.multilayer_efficiency(w, 1.0, 1.0, stack, (n_of(substrate, w), 0, 0, 0.0))
    efficiency = (r_s + r_p) / 2 if is_mirror else (t_s + t_p) / 2
    assert np.allclose(efficiency, efficiency_file, rtol=1e-4)  # test_multilayers.py:287


def test_rough_case_is_known_to_disagree_with_imd():
    # test_multilayers.py:223-248 marks SiC_Cr_Rough xfail ("IMD incorrectly uses the vacuum wavelength")
    g = np.load(GOLDEN / "imd_SiC_Cr_Rough.npz")
    w = g["wavelength_angstrom"]
    stack = [(n_of("SiC", w), 250.0, 1, 20.0), (n_of("Cr", w), 50.0, 1, 20.0)]
    r_s, r_p, _, _ = orm.multilayer_efficiency(w, 1.0, 1.0, stack, (n_of("SiO2", w), 0, 1, 20.0))
    assert not np.allclose((r_s + r_p) / 2, g["columns"][0], rtol=1e-4)


@pytest.mark.parametrize("direction", [1.0, 0.8, 0.3])
def test_energy_conservation_and_bounds(direction):
    # test_multilayers.py:88-109: 0 <= R, T and R + T <= 1
    w = np.linspace(100, 300, 50)
    stack = [(n_of("SiO2", w), 30.0, 1, 5.0), (n_of("Mo", w), 40.0, 0, 0.0)]
    r_s, r_p, t_s, t_p = orm.multilayer_efficiency(w, direction, 1.0, stack, (n_of("Si", w), 0, 0, 0.0))
    for r, t in ((r_s, t_s), (r_p, t_p)):
        assert np.all(r >= 0) and np.all(t >= 0)
        assert np.all(r + t <= 1 + 1e-12)


def test_periodic_equals_explicit():
    # optika/materials/_tests/test_layers.py:240-291
    w = np.linspace(120, 140, 40)
    mo = (n_of("Mo", w), 27.0, 1, 7.0)
    si = (n_of("Si", w), 40.0, 1, 7.0)
    explicit = [si, mo] * 7
    periodic = [("periodic", [si, mo], 7)]
    sub = (n_of("SiO2", w), 0, 0, 0.0)
    a = orm.multilayer_efficiency(w, 0.95, 1.0, explicit, sub)
    b = orm.multilayer_efficiency(w, 0.95, 1.0, periodic, sub)
    for x, y in zip(a, b):
        assert np.allclose(x, y)


def test_snells_law_scalar_identity():
    # optika/materials/_tests/test_snells_law.py:29-46: cos(arcsin(n1 sin(arccos c) / n2))
    c = np.linspace(0.05, 1, 20)
    for n1, n2 in ((1.0, 1.5), (1.0, 0.9 + 0.1j), (1.2 + 0.01j, 2.0)):
        expected = np.cos(np.arcsin(n1 * np.sin(np.arccos(c + 0j)) / n2))
        assert np.allclose(orm.snells_law_scalar(c, n1, n2), expected)


def test_unused_rough_transmission_file_agrees_loosely():
    """
    ``_data/SiO2_rough.txt`` (R, T, A for 50 A SiO2 on Si with 10 A erf interfaces) is shipped
    with the reference but used by none of its tests.  IMD applies the roughness factor with
    the vacuum wavelength (the reason ``SiC_Cr_Rough`` is an xfail there), so it cannot pin
    the model at 1e-4; it still bounds the transmissivity: within 2 % everywhere, 1e-4 typical.
    """
    g = np.load(GOLDEN / "imd_SiO2_rough.npz")
    w = g["wavelength_angstrom"]
    stack = [(n_of("SiO2", w), 50.0, 1, 10.0)]
    _, _, t_s, t_p = orm.multilayer_efficiency(w, 1.0, 1.0, stack, (n_of("Si", w), 0, 1, 10.0))
    err = np.abs((t_s + t_p) / 2 / g["columns"][1] - 1)
    assert err.max() < 2e-2 and np.median(err) < 2e-4
    assert np.allclose(g["columns"].sum(axis=0), 1.0, atol=1e-6)  # the file's own R + T + A
