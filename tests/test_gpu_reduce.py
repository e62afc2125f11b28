"""
Reductions over the pupil on the device (``optk_reduce_groups``): the
sums ``SequentialSystem.distortion`` / ``vignetting`` / ``area_effective`` take from the ray arrays
(``optika/systems/_sequential.py:1266-1285, 1351-1368, 1501-1506``) against the same reductions
done with NumPy on the rays brought back to the host.
"""

import ctypes as C

import numpy as np
import pytest
import torch

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import _engine, _lib

import configs

pytestmark = pytest.mark.gpu


def coated_system():
    from test_gpu_coatings import coated_grating, mo_si  # a MultilayerMirror on the cfg-2 grating

    return coated_grating(mo_si(6))


def host_moments(system, **kwargs):
    """The reference's expressions on host arrays (NumPy, named axes resolved by hand)."""
    result = system.rayfunction(**configs.PHYSICAL, **kwargs)
    rays = result.outputs
    axis_pupil = tuple(na.shape(result.inputs.pupil))
    shape_ = rays.shape
    names = list(shape_)
    k = tuple(names.index(ax) for ax in axis_pupil)

    def full(a):
        return na.broadcast_to(na.as_named_array(a), shape_).ndarray

    unv = full(rays.unvignetted).astype(bool)
    where = unv.any(axis=k)
    use = unv | ~np.expand_dims(where, k)
    with np.errstate(invalid="ignore"):
        x = np.mean(full(rays.position.x), axis=k, where=use)
        y = np.mean(full(rays.position.y), axis=k, where=use)
    illumination = unv.mean(axis=k)
    intensity = np.sum(full(rays.intensity), axis=k, where=unv)
    outer = [ax for ax in names if ax not in axis_pupil]
    return outer, where, x, y, illumination, intensity


@pytest.mark.parametrize(
    "make,kwargs",
    [
        (lambda: configs.newtonian(num_field=7, num_pupil=24), {}),
        (lambda: configs.toroidal_vls(num_field=5, num_pupil=40, num_wavelength=3), {}),  # 43 % vignetted by the octagon
        (lambda: configs.misaligned_telescope(num_field=5, num_pupil=16, num_pixel=64, num_tilt=3), {}),  # configuration axis
        # a multilayer-coated grating: chained launches, reduced from the dense rays (intensities carry the coating)
        (lambda: coated_system(), {}),
        # field angles far off axis: whole field points without a single surviving ray (the `~where` fallback)
        (
            lambda: configs.newtonian(num_field=5, num_pupil=16),
            dict(field=na.Cartesian2dVectorLinearSpace(
                -2.0 * u.deg, 2.0 * u.deg, axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=5, centers=True)),
        ),
    ],
    ids=["newtonian", "toroidal_vls", "misaligned_telescope", "coated_grating", "fully_vignetted_field_points"],
)
def test_pupil_moments_match_host_reductions(cuda_device, make, kwargs):
    system = make()
    got = system.pupil_moments(**configs.PHYSICAL, **kwargs)
    outer, where, x, y, illumination, intensity = host_moments(system, **kwargs)
    shape_ = {ax: n for ax, n in zip(outer, where.shape)}

    def dev(a):
        return na.broadcast_to(na.as_named_array(a), shape_).ndarray

    assert np.array_equal(dev(got["where"]), where)
    assert np.array_equal(dev(got["illumination"]), illumination)  # integer counts: exact
    finite = np.isfinite(x)
    assert np.array_equal(np.isfinite(dev(got["position"].x)), finite)
    assert np.allclose(dev(got["position"].x)[finite], x[finite], rtol=1e-12, atol=1e-12)
    assert np.allclose(dev(got["position"].y)[finite], y[finite], rtol=1e-12, atol=1e-12)
    assert np.allclose(dev(got["intensity"]), intensity, rtol=1e-12, atol=0)
    if kwargs:
        assert (~where).any() and where.any()


def test_reduce_groups_through_the_bare_abi(cuda_device):
    """Groups that straddle warps and CTAs, more than 512 CTAs (strided CTA order), optional outputs absent."""
    g = torch.Generator(device=cuda_device).manual_seed(3)
    n_groups, n_inner = 1531, 197  # 301 607 rays = 1179 CTAs
    n = n_groups * n_inner
    x = torch.rand(n, generator=g, device=cuda_device, dtype=torch.float64) - 0.5
    y = torch.rand(n, generator=g, device=cuda_device, dtype=torch.float64)
    w = torch.rand(n, generator=g, device=cuda_device, dtype=torch.float64)
    mask = (torch.rand(n, generator=g, device=cuda_device) > 0.4).to(torch.uint8)
    mask.reshape(n_groups, n_inner)[::7] = 0  # whole groups dropped
    out = {k: torch.zeros(n_groups, dtype=torch.float64, device=cuda_device) for k in ("i", "x", "y")}
    count = torch.zeros(n_groups, dtype=torch.int64, device=cuda_device)
    _lib.check(
        _lib.lib().optk_reduce_groups(
            n_groups, n_inner, x.data_ptr(), y.data_ptr(), w.data_ptr(), mask.data_ptr(), out["i"].data_ptr(),
            out["x"].data_ptr(), out["y"].data_ptr(), count.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream,
        )
    )
    m = mask.reshape(n_groups, n_inner).to(torch.float64)
    assert torch.equal(count, mask.reshape(n_groups, n_inner).sum(dim=1, dtype=torch.int64))
    assert int(count[::7].sum().item()) == 0
    for key, v in (("i", w), ("x", x), ("y", y)):
        want = (v.reshape(n_groups, n_inner) * m).sum(dim=1)
        assert torch.allclose(out[key], want, rtol=1e-12, atol=1e-12), key
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().optk_reduce_groups(1, 0, x.data_ptr(), y.data_ptr(), None, None, None, None, None, None, None, None, None))
