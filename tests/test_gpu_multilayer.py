"""
Parity of the multilayer transfer-matrix kernel (kernel 3) through
``optika_b200.materials.multilayer_efficiency`` against the oracle, against the
reference's IMD golden tables (``rtol=1e-4``, ``optika/materials/_tests/test_multilayers.py:287``)
and through its size-independent identities.
"""

import pathlib

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from oracle import multilayer as orm

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).parent / "golden"
M = optika.materials
TOL = 1e-9


def oracle_stack(layers, w):
    """Product layer objects -> oracle tuples (n, thickness, kind, width), all in mm."""
    out = []
    for layer in layers:
        if isinstance(layer, M.PeriodicLayerSequence):
            out.append(("periodic", oracle_stack(layer.layers, w), layer.num_periods))
        else:
            n = layer.n(w)
            n = n.ndarray if isinstance(n, na.ScalarArray) else n
            kind = 0 if layer.interface is None else layer.interface.kind
            width = 0.0 if layer.interface is None else layer.interface.width
            out.append((n, 0.0 if layer.thickness is None else layer.thickness, kind, width))
    return out


def assert_close(got, want, tol=TOL):
    got = np.asarray(got)
    want = np.broadcast_to(np.asarray(want), got.shape)
    scale = np.maximum(np.abs(want), 1e-300)
    err = np.abs(got - want) / scale
    ok = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), ok)
    assert err[ok].max() <= tol, f"max relative error {err[ok].max():.3e}"


CASES = [
    ("Si", None, M.Layer("Si"), False),
    ("SiO2", M.Layer("SiO2", thickness=50 * u.AA), M.Layer("Si"), False),
    ("SiO2_100A", M.Layer("SiO2", thickness=100 * u.AA), M.Layer("Si"), False),
    ("SiC_Cr", [M.Layer("SiC", thickness=25 * u.nm), M.Layer("Cr", thickness=5 * u.nm)], M.Layer("SiO2"), True),
]


@pytest.mark.parametrize("name,layers,substrate,is_mirror", CASES)
def test_vs_imd_golden_file(name, layers, substrate, is_mirror, cuda_device):
    g = np.load(GOLDEN / f"imd_{name}.npz")
    w = na.ScalarArray(g["wavelength_angstrom"] * u.AA, "wavelength")
    reflectivity, transmissivity = M.multilayer_efficiency(w, 1, 1, layers, substrate)
    efficiency = reflectivity.average if is_mirror else transmissivity.average
    assert np.allclose(efficiency.ndarray, g["columns"][0], rtol=1e-4)


def test_vs_oracle_layers_angles_roughness(cuda_device):
    w = na.linspace(80 * u.AA, 400 * u.AA, axis="wavelength", num=200)
    cos = na.linspace(0.05, 1.0, axis="angle", num=64)
    profiles = M.profiles
    layers = [
        M.Layer("SiO2", thickness=1.5 * u.nm, interface=profiles.ErfInterfaceProfile(0.4 * u.nm)),
        M.Layer("Mo", thickness=3.0 * u.nm, interface=profiles.ExponentialInterfaceProfile(0.5 * u.nm)),
        M.Layer("Si", thickness=4.0 * u.nm, interface=profiles.LinearInterfaceProfile(0.6 * u.nm)),
        M.Layer("Cr", thickness=2.0 * u.nm, interface=profiles.SinusoidalInterfaceProfile(0.3 * u.nm)),
    ]
    substrate = M.Layer("SiC", interface=profiles.ErfInterfaceProfile(0.7 * u.nm))
    r, t = M.multilayer_efficiency(w, cos, 1, layers, substrate)
    assert r.s.shape == {"wavelength": 200, "angle": 64}
    ww = w.ndarray[:, None]
    cc = cos.ndarray[None, :]
    stack = [(n[:, None], th, k, wd) for n, th, k, wd in oracle_stack(layers, w)]
    sub = oracle_stack([substrate], w)[0]
    want = orm.multilayer_efficiency(ww, cc, 1.0, stack, (sub[0][:, None], 0, sub[2], sub[3]))
    for got, exp in zip((r.s, r.p, t.s, t.p), want):
        assert_close(got.ndarray, exp)


def test_complex_ambient_index_and_direction(cuda_device):
    w = na.linspace(100 * u.AA, 200 * u.AA, axis="wavelength", num=50)
    n_amb = 1.2 + 0.01j
    direction = 0.7 + 0.02j
    layers = [M.Layer("Mo", thickness=3 * u.nm), M.Layer("Si", thickness=4 * u.nm)]
    r, t = M.multilayer_efficiency(w, direction, n_amb, layers, M.Layer("SiO2"))
    want = orm.multilayer_efficiency(w.ndarray, direction, n_amb, oracle_stack(layers, w), oracle_stack([M.Layer("SiO2")], w)[0])
    for got, exp in zip((r.s, r.p, t.s, t.p), want):
        assert_close(got.ndarray, exp)


def test_periodic_equals_explicit_and_oracle(cuda_device):
    # optika/materials/_tests/test_layers.py:240-291
    w = na.linspace(125 * u.AA, 140 * u.AA, axis="wavelength", num=128)
    cos = na.linspace(0.9, 1.0, axis="angle", num=8)
    si = M.Layer("Si", thickness=4.0 * u.nm, interface=M.profiles.ErfInterfaceProfile(0.7 * u.nm))
    mo = M.Layer("Mo", thickness=2.7 * u.nm, interface=M.profiles.ErfInterfaceProfile(0.7 * u.nm))
    periodic = M.PeriodicLayerSequence([si, mo], num_periods=30)
    explicit = M.LayerSequence([si, mo] * 30)
    substrate = M.Layer("SiO2")
    rp, tp = M.multilayer_efficiency(w, cos, 1, periodic, substrate)
    re, te = M.multilayer_efficiency(w, cos, 1, explicit, substrate)
    assert np.allclose(rp.s.ndarray, re.s.ndarray, rtol=1e-9)
    assert np.allclose(tp.p.ndarray, te.p.ndarray, rtol=1e-9, atol=1e-300)
    ww, cc = w.ndarray[:, None], cos.ndarray[None, :]
    stack = [(n[:, None], th, k, wd) for n, th, k, wd in oracle_stack([si, mo], w)] * 30
    sub = oracle_stack([substrate], w)[0]
    want = orm.multilayer_efficiency(ww, cc, 1.0, stack, (sub[0][:, None], 0, 0, 0.0))
    assert_close(re.s.ndarray, want[0])
    assert_close(re.p.ndarray, want[1])
    # a real Mo/Si mirror reflects strongly near 13.5 nm
    assert re.s.ndarray.max() > 0.5


def test_configuration_axis_on_thickness(cuda_device):
    """cfg 4 in miniature: thickness scaled along a named configuration axis."""
    w = na.linspace(125 * u.AA, 145 * u.AA, axis="wavelength", num=32)
    cos = na.linspace(0.87, 1.0, axis="angle", num=16)
    scale = na.linspace(0.95, 1.05, axis="config", num=5)
    si = M.Layer("Si", thickness=scale * 4.0 * u.nm)
    mo = M.Layer("Mo", thickness=scale * 2.7 * u.nm)
    layers = M.LayerSequence([si, mo] * 10)
    r, t = M.multilayer_efficiency(w, cos, 1, layers, M.Layer("SiO2"))
    assert r.s.shape == {"wavelength": 32, "angle": 16, "config": 5}
    n_si, n_mo = si.n(w).ndarray[:, None, None], mo.n(w).ndarray[:, None, None]
    sc = scale.ndarray[None, None, :]
    stack = [(n_si, sc * 4.0e-6, 0, 0.0), (n_mo, sc * 2.7e-6, 0, 0.0)] * 10
    sub = M.Layer("SiO2").n(w).ndarray[:, None, None]
    want = orm.multilayer_efficiency(w.ndarray[:, None, None], cos.ndarray[None, :, None], 1.0, stack, (sub, 0, 0, 0.0))
    assert_close(r.s.ndarray, want[0])
    assert_close(t.s.ndarray, want[2])


def test_identities_at_scale(cuda_device):
    """R + T <= 1, 0 <= R, T (test_multilayers.py:88-109) on a 512 x 256 grid; thick absorber -> T = 0 guard."""
    w = na.linspace(50 * u.AA, 500 * u.AA, axis="wavelength", num=512)
    cos = na.linspace(0.02, 1.0, axis="angle", num=256)
    layers = M.LayerSequence([M.Layer("Si", thickness=4 * u.nm), M.Layer("Mo", thickness=3 * u.nm)] * 20)
    r, t = M.multilayer_efficiency(w, cos, 1, layers, M.Layer("SiO2"))
    for a, b in ((r.s, t.s), (r.p, t.p)):
        assert (a.ndarray >= 0).all() and (b.ndarray >= 0).all()
        assert (a.ndarray + b.ndarray <= 1 + 1e-9).all()
    # a 1 mm thick chromium slab: |exp(-i beta)| overflows the 1e10 guard (_layers.py:271), t = 0
    r2, t2 = M.multilayer_efficiency(w, 1, 1, [M.Layer("Cr", thickness=1.0)], M.Layer("SiO2"))
    assert (t2.s.ndarray == 0).all() and np.isfinite(r2.s.ndarray).all()
    want = orm.multilayer_efficiency(w.ndarray, 1.0, 1.0, [(M.Layer("Cr").n(w).ndarray, 1.0, 0, 0.0)], (M.Layer("SiO2").n(w).ndarray, 0, 0, 0.0))
    assert_close(r2.s.ndarray, want[0])
