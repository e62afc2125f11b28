"""
Multilayer coatings as surface materials: ``MultilayerMirror`` /
``MultilayerFilm`` efficiency evaluated for every ray inside a system trace
(``optika/materials/_multilayers.py:839-866, 908-935``; chained launches around the coated
surface) against the oracle, which applies ``multilayer_efficiency`` to the ray arrays.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from oracle import raytrace as ora, binning as orb

import configs
import parity

pytestmark = pytest.mark.gpu
M = optika.materials


def mo_si(num_periods=10, scale=1.0):
    d, gamma = 6.65 * u.nm * scale, 0.6
    rough = M.profiles.ErfInterfaceProfile(0.7 * u.nm)
    return M.MultilayerMirror(
        layers=[
            M.Layer("SiO2", thickness=1 * u.nm),
            M.PeriodicLayerSequence(
                [M.Layer("Si", thickness=d * gamma, interface=rough), M.Layer("Mo", thickness=d * (1 - gamma), interface=rough)],
                num_periods=num_periods,
            ),
        ],
        substrate=M.Layer("SiO2", interface=rough),
    )


def coated_grating(material, rulings=None, num_wavelength=6):
    system = configs.spherical_grating(num_field=4, num_pupil=10, num_wavelength=num_wavelength, num_pixel=512)
    system.grid_input.wavelength = na.linspace(12.5 * u.nm, 14.5 * u.nm, axis="wavelength", num=num_wavelength)
    grating = system.surfaces[0]
    grating.material = material
    # 13.5 nm leaves this grating where 40 nm leaves the 1200 / mm one: onto the sensor
    spacing = (13.5 / 40 / 1200) * u.mm
    grating.rulings = optika.rulings.Rulings(spacing=spacing, diffraction_order=1) if rulings is None else rulings(spacing)
    return system


def states_of(system, out, n_surfaces, order):
    get = lambda a: na.as_named_array(a).numpy(("surface",) + order).reshape(n_surfaces, -1)  # noqa: E731
    return dict(
        wavelength=get(out.wavelength), px=get(out.position.x), py=get(out.position.y), pz=get(out.position.z),
        dx=get(out.direction.x), dy=get(out.direction.y), dz=get(out.direction.z), intensity=get(out.intensity),
        attenuation=get(out.attenuation), index_refraction=get(out.index_refraction),
        unvignetted=get(out.unvignetted).astype(bool),
    )


def test_unit_operation_matches_multilayer_efficiency(cuda_device):
    mirror = mo_si()
    n = 300
    rng = np.random.default_rng(0)
    d = rng.normal(size=(3, n)) * 0.2
    d[2] = 1
    d /= np.linalg.norm(d, axis=0)
    rays = optika.rays.RayVectorArray(
        wavelength=na.ScalarArray(rng.uniform(12 * u.nm, 15 * u.nm, n), "ray"),
        direction=na.Cartesian3dVectorArray(*[na.ScalarArray(c, "ray") for c in d]),
    )
    normal = na.Cartesian3dVectorArray(0.0, 0.0, -1.0)
    got = mirror.efficiency(rays, normal)
    r0, _ = configs.flatten_rays(rays)
    want = ora.material_efficiency(mirror, r0, (0.0, 0.0, -1.0))
    assert np.allclose(got.ndarray, want, rtol=1e-9, atol=1e-15)
    assert want.max() > 0.1  # the stack is tuned near 13.5 nm


@pytest.mark.parametrize("with_profile", [False, True])
def test_coated_grating_trace(cuda_device, with_profile):
    rulings = None
    if with_profile:
        rulings = lambda spacing: optika.rulings.SawtoothRulings(spacing=spacing, depth=4 * u.nm, diffraction_order=1)  # noqa: E731
    system = coated_grating(mo_si(), rulings)
    result = system.raytrace(accumulate=True, **configs.PHYSICAL)
    _, rays0 = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays0)
    states = ora.accumulate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
    got = states_of(system, result.outputs, len(system.surfaces_all), tuple(rays0.shape))
    parity.compare_states(got, states, system.surfaces_all)
    final = states["intensity"][-1]
    assert 0 < final.max() < 1 and np.ptp(final) > 1e-4
    # without accumulate: the same final state
    last = system.raytrace(accumulate=False, **configs.PHYSICAL).outputs
    assert np.allclose(
        na.as_named_array(last.intensity).numpy(tuple(rays0.shape)).reshape(-1), final, rtol=1e-9, atol=1e-15
    )


def test_film_and_mirror_in_one_system_with_a_configuration_axis(cuda_device):
    scale = na.ScalarArray(np.array([0.97, 1.0, 1.03]), "period")
    system = coated_grating(mo_si(num_periods=8, scale=scale))
    film = optika.surfaces.Surface(
        name="filter",
        material=M.MultilayerFilm(layers=[M.Layer("Si", thickness=100 * u.nm), M.Layer("SiO2", thickness=2 * u.nm)]),
        transformation=optika.transformations.Cartesian3dTranslation(z=500 * u.mm),
    )
    system.surfaces = [film] + list(system.surfaces)
    system.invalidate()
    assert system.shape == {"period": 3}
    result = system.raytrace(accumulate=True, **configs.PHYSICAL)
    _, rays0 = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays0)
    flat0 = {k: v.reshape(-1) for k, v in r0.items()}
    out = result.outputs
    order = tuple(rays0.shape)
    n_s = len(system.surfaces_all)
    totals = []
    for c in range(3):
        surfaces = ora.select_config(system.surfaces_all, {"period": c})
        states = ora.accumulate_rays(surfaces, flat0, extended=True)
        sel = lambda a: na.as_named_array(a).numpy(("period", "surface") + order)[c].reshape(n_s, -1)  # noqa: E731
        got = dict(
            wavelength=sel(out.wavelength), px=sel(out.position.x), py=sel(out.position.y), pz=sel(out.position.z),
            dx=sel(out.direction.x), dy=sel(out.direction.y), dz=sel(out.direction.z), intensity=sel(out.intensity),
            attenuation=sel(out.attenuation), index_refraction=sel(out.index_refraction),
            unvignetted=sel(out.unvignetted).astype(bool),
        )
        parity.compare_states(got, states, surfaces)
        assert np.all(states["intensity"][1] < 1)  # the film absorbs
        totals.append(states["intensity"][-1].sum())
    assert len({round(t, 6) for t in totals}) == 3  # the period scale changes the reflectivity


def test_fused_image_of_a_coated_system(cuda_device):
    system = coated_grating(mo_si())
    edges = na.ScalarArray(np.array([12 * u.nm, 13.5 * u.nm, 15 * u.nm]), "wavelength")
    image = system.image_rays(edges, counts=True, **configs.PHYSICAL)
    _, rays0 = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays0)
    out = ora.propagate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    want, _, _ = orb.collect(local, edges.ndarray, ex, ey)
    flux = image.flux.cpu().numpy()
    assert want.sum() > 0 and np.isclose(flux.sum(), want.sum(), rtol=1e-9)
    assert (~np.isclose(flux, want, rtol=1e-9, atol=1e-9 * want.max())).sum() <= 8
    assert np.array_equal(image.counts.cpu().numpy(), orb.counts(local, edges.ndarray, ex, ey))


def test_image_of_a_scene_through_a_coated_system(cuda_device):
    from oracle import grid as og

    system = coated_grating(mo_si())
    nf = 6
    field = na.Cartesian2dVectorLinearSpace(
        -0.05 * u.deg, 0.05 * u.deg, na.Cartesian2dVectorArray("field_x", "field_y"), nf + 1
    )
    w = na.linspace(12.8 * u.nm, 14.2 * u.nm, "wavelength", 4)
    rng = np.random.default_rng(4)
    scene = na.FunctionArray(
        inputs=optika.vectors.SpectralPositionalVectorArray(wavelength=w, position=field),
        outputs=na.ScalarArray(rng.uniform(1e9, 2e9, (3, nf, nf)), ("wavelength", "field_x", "field_y")),
    )
    pupil = na.Cartesian2dVectorLinearSpace(
        -40 * u.mm, 40 * u.mm, na.Cartesian2dVectorArray("pupil_x", "pupil_y"), 15
    )
    image = system.image(scene, pupil=pupil, noise=False, normalized_pupil=False, seed=6)
    v = [w.ndarray, field.x.ndarray, field.y.ndarray, pupil.x.ndarray, pupil.y.ndarray]
    aw, af, ap = og.cell_area(v, True, False)
    rays0 = og.input_rays(v, weight_scene=scene.outputs.ndarray * aw[:, None, None] * af[None], weight_pupil=ap, seed=6)
    out = ora.propagate_rays(system.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    want, _, _ = orb.collect(local, np.array([w.ndarray.min(), w.ndarray.max()]), ex, ey)
    got = image.outputs.ndarray
    assert want.sum() > 0 and np.isclose(got.sum(), want.sum(), rtol=1e-9)
    assert (~np.isclose(got, want, rtol=1e-9, atol=1e-9 * want.max())).sum() <= 8
    # the coating matters: far from unit efficiency
    assert got.sum() < 0.5 * rays0["intensity"].sum()


def test_image_through_two_coated_surfaces(cuda_device):
    """Filter film + multilayer grating: the grid chain applies two coatings (`trace_grid_coated`)."""
    from oracle import grid as og

    system = coated_grating(mo_si(num_periods=6))
    film = optika.surfaces.Surface(
        name="filter",
        material=M.MultilayerFilm(layers=[M.Layer("Si", thickness=80 * u.nm)]),
        transformation=optika.transformations.Cartesian3dTranslation(z=300 * u.mm),
    )
    system.surfaces = [film] + list(system.surfaces)
    system.invalidate()
    assert sorted(system._compiled_local.coatings) == [1, 2]
    nf = 4
    field = na.Cartesian2dVectorLinearSpace(
        -0.05 * u.deg, 0.05 * u.deg, na.Cartesian2dVectorArray("field_x", "field_y"), nf + 1
    )
    w = na.linspace(13.0 * u.nm, 14.0 * u.nm, "wavelength", 3)
    scene = na.FunctionArray(
        inputs=optika.vectors.SpectralPositionalVectorArray(wavelength=w, position=field),
        outputs=na.ScalarArray(np.full((2, nf, nf), 1e9), ("wavelength", "field_x", "field_y")),
    )
    pupil = na.Cartesian2dVectorLinearSpace(
        -40 * u.mm, 40 * u.mm, na.Cartesian2dVectorArray("pupil_x", "pupil_y"), 13
    )
    image = system.image(scene, pupil=pupil, noise=False, normalized_pupil=False, seed=3)
    v = [w.ndarray, field.x.ndarray, field.y.ndarray, pupil.x.ndarray, pupil.y.ndarray]
    aw, af, ap = og.cell_area(v, True, False)
    rays0 = og.input_rays(v, weight_scene=scene.outputs.ndarray * aw[:, None, None] * af[None], weight_pupil=ap, seed=3)
    out = ora.propagate_rays(system.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    want, _, _ = orb.collect(local, np.array([w.ndarray.min(), w.ndarray.max()]), ex, ey)
    got = image.outputs.ndarray
    assert want.sum() > 0 and np.isclose(got.sum(), want.sum(), rtol=1e-9)
    assert (~np.isclose(got, want, rtol=1e-9, atol=1e-9 * want.max())).sum() <= 8


# ---------------------------------------------------------------------------
# efficiency tables: coating="table" against the exact per-ray chain
# ---------------------------------------------------------------------------
def _intensity(system, **kwargs):
    out = system.raytrace(accumulate=False, **configs.PHYSICAL, **kwargs).outputs
    return na.as_named_array(out.intensity), out


@pytest.mark.parametrize("num_periods", [6, 30])
def test_coating_table_matches_the_exact_chain_to_the_stated_tolerance(cuda_device, num_periods):
    system = coated_grating(mo_si(num_periods), num_wavelength=16)
    exact, rays_exact = _intensity(system)
    system.coating = "table"
    table, rays_table = _intensity(system)
    scale = float(np.nanmax(np.abs(exact.ndarray)))
    assert scale > 0.05  # the stack reflects near 13.5 nm
    assert np.nanmax(np.abs(table.ndarray - exact.ndarray)) <= 1e-6 * scale  # the contract of coating_tolerance
    # one exact node per wavelength of the grid: only the cosine is interpolated; the walk is the same up to
    # rounding (one fused launch here, chained launches of other kernel instantiations there)
    for name in ("x", "y", "z"):
        a, b = getattr(rays_table.position, name), getattr(rays_exact.position, name)
        assert np.allclose(na.as_named_array(a).ndarray, na.as_named_array(b).ndarray, rtol=1e-12, atol=1e-12, equal_nan=True)
    assert np.array_equal(rays_table.unvignetted.ndarray, rays_exact.unvignetted.ndarray)
    tabled = next(iter(system._compiled.__dict__["_tabled"].values()))
    (error_cos, error_wavelength), = set(tabled.errors.values())
    assert error_cos <= 0.5e-6 and error_wavelength == 0.0
    assert all(t.exact_in_wavelength for t in tabled.tables.values())
    # a tighter tolerance is honoured (more cosine nodes), and the tables are rebuilt for it
    system.coating_tolerance = 1e-9
    tight, _ = _intensity(system)
    assert np.nanmax(np.abs(tight.ndarray - exact.ndarray)) <= 1e-9 * scale


def test_coating_table_in_the_fused_image_and_in_the_pupil_reductions(cuda_device):
    system = coated_grating(mo_si(10), num_wavelength=8)
    edges = na.ScalarArray(np.array([12 * u.nm, 15 * u.nm]), "wavelength")
    exact = system.image_rays(edges, counts=True, **configs.PHYSICAL)
    moments_exact = system.pupil_moments(**configs.PHYSICAL)
    system.coating = "table"
    table = system.image_rays(edges, counts=True, **configs.PHYSICAL)
    moments_table = system.pupil_moments(**configs.PHYSICAL)
    assert np.array_equal(table.counts.cpu().numpy(), exact.counts.cpu().numpy())
    a, b = table.flux.cpu().numpy(), exact.flux.cpu().numpy()
    assert b.max() > 0 and np.abs(a - b).max() <= 2e-6 * b.max()
    ia, ib = moments_table["intensity"].ndarray, moments_exact["intensity"].ndarray
    assert np.abs(ia - ib).max() <= 2e-6 * ib.max()
    assert np.array_equal(moments_table["where"].ndarray, moments_exact["where"].ndarray)


def test_coating_table_over_a_continuous_wavelength_range(cuda_device):
    """Dense device rays carry arbitrary wavelengths: nodes on every kink of the optical constants, linear in between."""
    import torch
    from optika_b200 import _engine

    system = coated_grating(mo_si(10), num_wavelength=5)
    system.coating_tolerance = 2e-5  # what a 2-D table of this band reaches within the default size limit
    _, rays = system._input(None, None, None, None, False, False)
    start = _engine.trace(system._compiled, rays, surf_count=0, ray_axes_order=system._ray_axes_order)  # dense copies
    n = start.size
    rng = torch.Generator(device="cuda").manual_seed(3)
    start.fields["wavelength"] = (12.6e-6 + 1.8e-6 * torch.rand(n, generator=rng, device="cuda", dtype=torch.float64))
    exact = _engine.trace(system._compiled, start)
    compiled = system._compiled
    compiled.coating = "table"
    table = _engine.trace(compiled, start)
    a, b = table.fields["intensity"].cpu().numpy(), exact.fields["intensity"].cpu().numpy()
    assert np.nanmax(b) > 0.05
    assert np.nanmax(np.abs(a - b)) <= 2e-5 * np.nanmax(b)
    tabled = [t for t in compiled.__dict__["_tabled"].values() if t != "exact"]
    assert tabled and not any(t.exact_in_wavelength for t in tabled[0].tables.values())


def test_rays_outside_the_tabulated_cosines_fail_loudly(cuda_device):
    system = coated_grating(mo_si(6), num_wavelength=4)
    system.coating = "table"
    compiled = system._compiled
    compiled.coating_cos_range = (0.2, 0.5)  # the rays arrive at near-normal incidence: cos ~ 1
    with pytest.raises(ValueError, match="outside its efficiency table"):
        system.raytrace(accumulate=False, **configs.PHYSICAL)


def test_glass_before_the_coating_takes_the_exact_chain(cuda_device):
    """A ray that reaches the coating inside a medium has a complex ambient index: not a function of two numbers."""
    system = coated_grating(mo_si(6), num_wavelength=4)
    window = optika.surfaces.Surface(
        name="window", material=M.Glass.n_bk7(),
        transformation=optika.transformations.Cartesian3dTranslation(z=10 * u.mm),
    )
    system.surfaces = [window] + list(system.surfaces)
    exact, _ = _intensity(system)
    system.coating = "table"
    table, _ = _intensity(system)
    assert np.array_equal(table.ndarray, exact.ndarray, equal_nan=True)  # the same (exact) route ran
    assert "_tabled" not in system._compiled.__dict__
