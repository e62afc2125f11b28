"""
Run-time specialised kernels (``csrc/jit.cu``): the walk compiled by NVRTC for one surface list
must give bit-identical results to the table-driven kernels, on every variant it serves
(strided / dense / on-device grid input, with and without the fused image, efficiencies).
"""

import numpy as np
import pytest
import torch

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import _engine, _grid, _lib

import configs
from test_gpu_trace import torch_equal

pytestmark = pytest.mark.gpu


@pytest.fixture()
def jit():
    lib = _lib.lib()

    class Mode:
        def __call__(self, mode):
            _lib.check(lib.optk_jit_mode(mode))

        @property
        def compiled(self):
            return lib.optk_jit_compiled()

    mode = Mode()
    yield mode
    mode(-1)


def same(a: _engine.DeviceRays, b: _engine.DeviceRays):
    for name in a.fields:
        assert torch_equal(a.fields[name], b.fields[name]), name
    assert torch_equal(a.unvignetted, b.unvignetted)


def same_to_rounding(a: _engine.DeviceRays, b: _engine.DeviceRays, rtol=1e-12):
    """
    Two compilations of the same expressions may contract different multiply-add pairs into
    FMAs (`a * b + c * d` leaves the choice to the compiler), so across ALL element kinds the
    specialised kernels are held to rounding-level agreement: identical masks, identical NaN / inf
    patterns, finite values within `rtol` (a few ulp through the chain of one surface).
    """
    assert torch_equal(a.unvignetted, b.unvignetted)
    for name in a.fields:
        x, y = a.fields[name].reshape(-1), b.fields[name].reshape(-1)
        assert torch.equal(torch.isnan(x), torch.isnan(y)), name
        inf = torch.isinf(y)
        assert torch.equal(torch.isinf(x), inf), name
        assert torch.equal(x[inf], y[inf]), name
        ok = torch.isfinite(y)
        scale = torch.clamp(y[ok].abs(), min=1.0)
        err = ((x[ok] - y[ok]).abs() / scale).max().item() if ok.any() else 0.0
        assert err <= rtol, (name, err)


SYSTEMS = {
    "newtonian": lambda: configs.newtonian(num_field=5, num_pupil=12, num_pixel=64),
    "grating": lambda: configs.spherical_grating(num_field=4, num_pupil=10, num_wavelength=3, num_pixel=256),
    "toroidal_vls": lambda: configs.toroidal_vls(4, 10, 2),
    "misaligned": lambda: configs.misaligned_telescope(4, 8, 64, 3),
}


@pytest.mark.parametrize("name", list(SYSTEMS))
def test_specialised_kernels_are_bit_identical(cuda_device, jit, name):
    system = SYSTEMS[name]()
    _, rays = system._input(None, None, None, None, False, False)
    order = system._ray_axes_order
    edges = na.ScalarArray(np.array([1e-6, 1e-2]), "wavelength")
    jit(0)
    want = _engine.trace(system._compiled, rays, ray_axes_order=order)
    dense_in = _engine.trace(system._compiled, rays, ray_axes_order=order, surf_count=0)
    want_dense = _engine.trace(system._compiled, dense_in)
    want_image = system.image_rays(edges, counts=True, **configs.PHYSICAL)
    before = jit.compiled
    jit(1)
    got = _engine.trace(system._compiled, rays, ray_axes_order=order)
    got_dense = _engine.trace(system._compiled, dense_in)
    got_image = system.image_rays(edges, counts=True, **configs.PHYSICAL)
    assert jit.compiled > before, "nothing was compiled: NVRTC unavailable?"
    same(got, want)
    same(got_dense, want_dense)
    assert torch.equal(got_image.counts, want_image.counts)
    assert torch.allclose(got_image.flux, want_image.flux, rtol=1e-12, atol=0)
    # a second call reuses the cached kernels
    count = jit.compiled
    _engine.trace(system._compiled, rays, ray_axes_order=order)
    assert jit.compiled == count


def test_specialised_grid_kernel_and_efficiencies(cuda_device, jit):
    from test_oracle_efficiency import measured

    system = configs.spherical_grating(num_field=4, num_pupil=10, num_wavelength=3, num_pixel=256)
    w = np.linspace(150 * u.AA, 650 * u.AA, 9)
    system.surfaces[0].material = optika.materials.MeasuredMirror(measured(np.linspace(0.2, 0.9, 9), w))
    system.surfaces[0].rulings = optika.rulings.SawtoothRulings(
        spacing=(1 / 1200) * u.mm, depth=10 * u.nm, diffraction_order=1
    )
    system.invalidate()
    deg = u.deg
    v = [
        np.linspace(17 * u.nm, 63 * u.nm, 4), np.linspace(-0.05 * deg, 0.05 * deg, 7), np.linspace(-0.05 * deg, 0.05 * deg, 6),
        np.linspace(-45, 45, 13), np.linspace(-45, 45, 12),
    ]
    grid = _grid.RayGrid(v, seed=4)
    compiled = system._compiled_local
    ex, ey = system.sensor.pixel_edges()
    ew = np.array([v[0][0], v[0][-1]])
    results = []
    for mode in (0, 1):
        jit(mode)
        image = _engine.DeviceImage.zeros(ew, ex, ey, cuda_device, moments=True, counts=True)
        rays = _grid.trace_grid(compiled, grid, image=image)
        results.append((rays, image))
    same(results[1][0], results[0][0])
    assert torch.equal(results[1][1].counts, results[0][1].counts)
    assert results[0][1].counts.sum().item() > 0
    assert float(results[0][0].fields["intensity"].max()) < 1.0  # the efficiencies were applied


def _element_surfaces():
    """One surface per element kind that the specialised kernels inline (test_gpu_trace's zoo)."""
    import test_gpu_trace as zoo

    S = optika.surfaces.Surface
    surfaces = []
    for sag in zoo.SAGS:
        if getattr(sag, "transformation", None) is None:  # sag transformations take the generic kernel
            surfaces.append((f"sag-{type(sag).__name__}-mirror", S(sag=sag, material=optika.materials.Mirror(), transformation=zoo.T_LIST)))
            surfaces.append((f"sag-{type(sag).__name__}-glass", S(sag=sag, material=optika.materials.Glass.n_bk7())))
    for k, aperture in enumerate(zoo.APERTURES):
        surfaces.append((f"aperture-{k}-{type(aperture).__name__}", S(aperture=aperture, transformation=zoo.T_LIST)))
    for k, rulings in enumerate(zoo.RULINGS):
        surfaces.append(
            (
                f"rulings-{k}-{type(rulings.spacing_).__name__}",
                S(sag=optika.sags.ToroidalSag(400.0, 450.0), rulings=rulings, material=optika.materials.Mirror(), transformation=zoo.T_LIST),
            )
        )
    return surfaces


@pytest.mark.parametrize("name,surface", _element_surfaces(), ids=[n for n, _ in _element_surfaces()])
def test_every_element_kind_agrees_when_specialised(cuda_device, jit, name, surface):
    """
    The specialised kernels inline the element kinds that the table-driven kernels call out of
    line (conic, cylinder, toroid with the paired Newton loop, polynomial and holographic
    rulings, polygon / sector / elliptical apertures, Sellmeier glass) and fold loop lengths and
    exponents: every one of them must reproduce the table-driven result to rounding (see
    `same_to_rounding`; the BASELINE systems above are bit-identical), rays that miss the surface
    (NaN / inf) included.
    """
    import test_gpu_trace as zoo

    rays = zoo.random_rays(n=6001, spread=60.0, wavelength=300 * u.nm)  # odd count: the last thread holds one ray
    system = _engine.CompiledSystem([surface])
    dense_in = _engine.trace(system, rays, surf_count=0)
    broadcast = name.startswith("rulings")  # strided-input variant too (one more compilation) for a few
    jit(0)
    want = _engine.trace(system, dense_in)
    want_broadcast = _engine.trace(system, rays) if broadcast else None
    before = jit.compiled
    jit(1)
    got = _engine.trace(system, dense_in)
    got_broadcast = _engine.trace(system, rays) if broadcast else None
    # (surfaces of the same shape -- kinds, flags, loop lengths -- share one compiled kernel)
    assert jit.compiled >= before and jit.compiled > 0
    same_to_rounding(got, want)
    if broadcast:
        same_to_rounding(got_broadcast, want_broadcast)
