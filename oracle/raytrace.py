"""
ORACLE (test infrastructure, not product code): NumPy fp64 restatement of the
reference's sequential raytrace, ``optika.propagators.propagate_rays`` ->
``AbstractSurface.propagate_rays`` -> sag / rulings / Snell / aperture.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module; the product (``optika_b200``) never does.

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).  Rays are a plain ``dict`` of equally shaped float64 arrays
``wavelength, px, py, pz, dx, dy, dz, intensity, attenuation, index_refraction``
plus the boolean ``unvignetted``; lengths in mm, angles in radians.  Optical
elements are *duck typed*: the oracle dispatches on the class **name**
(``"SphericalSag"``, ``"RectangularAperture"``, ...) and reads the attributes the
reference classes define, so it never imports the product's classes.  All
parameters must be scalars (one configuration); use :func:`select_config` to
slice objects whose parameters carry named configuration axes.

Pinning: the reference cannot be imported in the build
container (``named_arrays``/``astropy`` missing), so this restatement is pinned
by the reference's own known-answer tests, restated in ``tests/test_oracle_*.py``:
Snell identities (``optika/materials/_tests/test_snells_law.py:82-120``), the
``direction`` convention (``optika/_util_test.py:20-32``), sag on-surface and
closed-form == iterative (``optika/sags/_tests/_abc_test.py:90-102``), the grazing
conic regression (``optika/sags/_tests/_conic_test.py:57-87``), parabola normal ==
conic normal (``optika/sags/_tests/_parabolic_test.py:25-37``), Glass n_d
(``optika/materials/_tests/test_materials.py:126-150``) and the holographic
refocusing example (``optika/rulings/_spacing.py:141-219``).
Parity UNPINNED (third-party ``named_arrays`` behaviour assumed, not verifiable
here): ``na.geometry.point_in_polygon`` on polygon *edges*,
``na.optimize.root_secant`` iteration details, ``na.histogram`` edge handling.
"""

from __future__ import annotations
import dataclasses
import numpy as np

__all__ = [
    "FIELDS",
    "make_rays",
    "copy_rays",
    "select_config",
    "transform_forward",
    "transform_inverse",
    "sag_value",
    "sag_normal",
    "sag_intercept",
    "sag_propagate",
    "ruling_vector",
    "incident_effective",
    "index_refraction",
    "snells_law",
    "aperture_mask",
    "aperture_margin",
    "surface_propagate",
    "propagate_rays",
    "accumulate_rays",
    "direction",
    "angles",
]

FIELDS = (
    "wavelength",
    "px",
    "py",
    "pz",
    "dx",
    "dy",
    "dz",
    "intensity",
    "attenuation",
    "index_refraction",
)


def make_rays(n: int | tuple, **kwargs) -> dict:
    """Rays with the defaults of ``optika/rays/_ray_vectors.py:256-278``."""
    shape = (n,) if np.isscalar(n) else tuple(n)
    defaults = dict(
        wavelength=0.0, px=0.0, py=0.0, pz=0.0, dx=0.0, dy=0.0, dz=0.0,
        intensity=1.0, attenuation=0.0, index_refraction=1.0,
    )
    rays = {}
    for k, v in defaults.items():
        rays[k] = np.broadcast_to(np.asarray(kwargs.get(k, v), dtype=np.float64), shape).copy()
    rays["unvignetted"] = np.broadcast_to(
        np.asarray(kwargs.get("unvignetted", True), dtype=bool), shape
    ).copy()
    return rays


def copy_rays(rays: dict) -> dict:
    return {k: np.array(v, copy=True) for k, v in rays.items()}


def _name(obj) -> str:
    return type(obj).__name__


def _f(a) -> float:
    """Scalar parameter -> float (accepts 0-d arrays and 0-d named arrays)."""
    if hasattr(a, "ndarray"):
        a = a.ndarray
    a = np.asarray(a)
    if a.ndim != 0:
        raise ValueError("oracle parameters must be scalars; use select_config()")
    return a.item()


def select_config(obj, index: dict):
    """
    Copy of a (dataclass) optical element with every named-array parameter
    indexed by ``index`` (``{axis: i}``), recursively.  Third-party
    ``named_arrays`` indexing ``a[dict(axis=i)]`` restated for duck-typed arrays
    exposing ``.ndarray`` and ``.axes``.
    """
    if obj is None or isinstance(obj, (bool, int, float, complex, str)):
        return obj
    if hasattr(obj, "ndarray") and hasattr(obj, "axes"):
        nd = np.asarray(obj.ndarray)
        idx = tuple(index.get(ax, slice(None)) for ax in obj.axes)
        axes = tuple(ax for ax in obj.axes if ax not in index)
        return type(obj)(nd[idx], axes)
    if isinstance(obj, dict):
        return {k: select_config(v, index) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(select_config(v, index) for v in obj)
    if dataclasses.is_dataclass(obj):
        changes = {
            f.name: select_config(getattr(obj, f.name), index)
            for f in dataclasses.fields(obj)
            if f.init
        }
        return dataclasses.replace(obj, **changes)
    return obj


# ---------------------------------------------------------------------------
# transformations (third-party na.transformations; call sites
# optika/surfaces.py:141-142, 195-196; optika/rays/_ray_vectors.py:105-166)
# ---------------------------------------------------------------------------
def _rotation(name: str, angle: float) -> np.ndarray:
    c, s = np.cos(angle), np.sin(angle)
    if name.endswith("X"):
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=float)
    if name.endswith("Y"):
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=float)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=float)


def _primitives(t) -> list:
    """Flatten a transformation into (matrix, vector) steps, first applied first."""
    if t is None:
        return []
    name = _name(t)
    if name == "TransformationList":
        result = []
        for item in t.transformations:
            result += _primitives(item)
        return result
    if name in ("Cartesian3dTranslation",):
        return [(np.eye(3), np.array([_f(t.x), _f(t.y), _f(t.z)]))]
    if name.startswith("Cartesian3dRotation"):
        return [(_rotation(name, _f(t.angle)), np.zeros(3))]
    if name == "IdentityTransformation":
        return []
    raise NotImplementedError(f"oracle: transformation {name}")


def _apply(m, v, x, y, z, is_direction):
    rx = m[0, 0] * x + m[0, 1] * y + m[0, 2] * z
    ry = m[1, 0] * x + m[1, 1] * y + m[1, 2] * z
    rz = m[2, 0] * x + m[2, 1] * y + m[2, 2] * z
    if not is_direction:
        rx, ry, rz = rx + v[0], ry + v[1], rz + v[2]
    return rx, ry, rz


def transform_forward(t, x, y, z, is_direction=False):
    """``t(vector)``: each primitive in list order, position affine / direction linear."""
    for m, v in _primitives(t):
        x, y, z = _apply(m, v, x, y, z, is_direction)
    return x, y, z


def transform_inverse(t, x, y, z, is_direction=False):
    """``t.inverse(vector)``: inverse primitives in reverse order, ``R^T (p - t)``."""
    for m, v in reversed(_primitives(t)):
        if not is_direction:
            x, y, z = x - v[0], y - v[1], z - v[2]
        x, y, z = _apply(m.T, np.zeros(3), x, y, z, True)
    return x, y, z


def _rays_transform(t, rays, inverse=False):
    if t is None:
        return rays
    f = transform_inverse if inverse else transform_forward
    rays = dict(rays)
    rays["px"], rays["py"], rays["pz"] = f(t, rays["px"], rays["py"], rays["pz"], False)
    rays["dx"], rays["dy"], rays["dz"] = f(t, rays["dx"], rays["dy"], rays["dz"], True)
    return rays


# ---------------------------------------------------------------------------
# sags
# ---------------------------------------------------------------------------
def _sag_radius_conic(sag):
    name = _name(sag)
    if name == "ParabolicSag":
        # optika/sags/_parabolic.py:40-46: radius = 2 f, conic = -1
        return 2 * _f(sag.focal_length), -1.0
    if name == "ConicSag":
        return _f(sag.radius), _f(sag.conic)
    raise NotImplementedError(name)


def _sag_value_local(sag, x, y):
    """z(x, y) in the sag's own frame."""
    name = _name(sag)
    with np.errstate(invalid="ignore", divide="ignore"):
        if name == "NoSag":
            # optika/sags/_flat.py:26-41
            return np.zeros(np.broadcast(x, y).shape)
        if name == "SphericalSag":
            # optika/sags/_spherical.py:100-125
            c = 1 / _f(sag.radius)
            r2 = np.square(x) + np.square(y)
            return c * r2 / (1 + np.sqrt(1 - np.square(c) * r2))
        if name == "CylindricalSag":
            # optika/sags/_cylindrical.py:75-92
            c = 1 / _f(sag.radius)
            r2 = np.square(x)
            return c * r2 / (1 + np.sqrt(1 - np.square(c) * r2)) + 0 * y
        if name in ("ConicSag", "ParabolicSag"):
            # optika/sags/_conic.py:36-54
            radius, conic = _sag_radius_conic(sag)
            c = 1 / radius
            r2 = np.square(x) + np.square(y)
            return c * r2 / (1 + np.sqrt(1 - (1 + conic) * np.square(c) * r2))
        if name == "ToroidalSag":
            # optika/sags/_toroidal.py:38-58
            c = 1 / _f(sag.radius)
            r = _f(sag.radius_of_rotation)
            x2 = np.square(x)
            y2 = np.square(y)
            zy = c * y2 / (1 + np.sqrt(1 - np.square(c) * y2))
            return r - np.sqrt(np.square(r - zy) - x2)
    raise NotImplementedError(f"oracle: sag {name}")


def sag_value(sag, x, y, z=0.0):
    """``sag(position)``; the sag's own transformation is inverted first."""
    t = getattr(sag, "transformation", None)
    x, y, z = np.broadcast_arrays(*[np.asarray(a, dtype=float) for a in (x, y, z)])
    if t is not None:
        x, y, z = transform_inverse(t, x, y, z)
    if _name(sag) == "NoSag" and t is not None:
        # optika/sags/_flat.py:31-41: z of the transformed (x, y, 0)
        _, _, zz = transform_forward(t, x, y, np.zeros_like(x))
        return zz
    return _sag_value_local(sag, x, y)


def sag_normal(sag, x, y, z=0.0):
    """``sag.normal(position)`` -> (nx, ny, nz); not rotated back, as in the reference."""
    name = _name(sag)
    t = getattr(sag, "transformation", None)
    x, y, z = np.broadcast_arrays(*[np.asarray(a, dtype=float) for a in (x, y, z)])
    if name == "NoSag":
        # optika/sags/_flat.py:43-47 (transformation ignored)
        return np.zeros_like(x), np.zeros_like(x), -np.ones_like(x)
    if t is not None:
        x, y, z = transform_inverse(t, x, y, z)
    with np.errstate(invalid="ignore", divide="ignore"):
        if name == "SphericalSag":
            # optika/sags/_spherical.py:127-146
            c = 1 / _f(sag.radius)
            nx = c * x
            ny = c * y
            nz = -np.sqrt(1 - nx**2 - ny**2)
            return nx, ny, nz
        if name == "CylindricalSag":
            # optika/sags/_cylindrical.py:94-113
            nx = x / _f(sag.radius)
            return nx, np.zeros_like(x), -np.sqrt(1 - nx**2)
        if name == "ParabolicSag":
            # optika/sags/_parabolic.py:48-63: (x, y, -R) / sqrt((x/R)^2 + (y/R)^2 + 1) / R
            r = 2 * _f(sag.focal_length)
            d = np.sqrt((x / r) ** 2 + (y / r) ** 2 + 1)
            return x / d / r, y / d / r, -r / d / r + 0 * x
        if name == "ConicSag":
            # optika/sags/_conic.py:56-81
            radius, conic = _sag_radius_conic(sag)
            c = 1 / radius
            g = np.sqrt(1 - (1 + conic) * np.square(c) * (np.square(x) + np.square(y)))
            dzdx, dzdy = c * x / g, c * y / g
            length = np.sqrt(np.square(dzdx) + np.square(dzdy) + 1)
            return dzdx / length, dzdy / length, -1 / length
        if name == "ToroidalSag":
            # optika/sags/_toroidal.py:60-88
            c = 1 / _f(sag.radius)
            r = _f(sag.radius_of_rotation)
            x2 = np.square(x)
            y2 = np.square(y)
            c2 = np.square(c)
            g = np.sqrt(1 - c2 * y2)
            zy = c * y2 / (1 + g)
            f = np.sqrt(np.square(r - zy) - x2)
            dzdx = x / f
            dzydy = c * y / g
            dzdy = (r - zy) * dzydy / f
            length = np.sqrt(np.square(dzdx) + np.square(dzdy) + 1)
            return dzdx / length, dzdy / length, -1 / length
    raise NotImplementedError(f"oracle: sag {name}")


def _intercept_secant(sag, o, d, min_step_size=1e-6, max_iterations=100, converge=False):
    """
    Generic intercept, ``optika/sags/_abc.py:76-107``: root of
    ``f(t) = z(t) - sag(x(t), y(t))`` by ``na.optimize.root_secant(guess=0 mm,
    min_step_size=1e-6 mm)``.  ``root_secant`` is third-party (named_arrays ~= 2.1,
    not in /root/reference): restated as the textbook secant iteration started at
    ``t0 = 0`` with a first step of ``10 * min_step_size``, iterating the WHOLE
    array until every ray's last step is below `min_step_size` (rays that have
    already converged keep iterating, as whole-array code does).
    ``converge=True`` instead iterates every ray to machine precision
    (step < 4 ulp-ish); this is the "tightened oracle" of SURVEY.md section 7.
    """
    ox, oy, oz = o
    ux, uy, uz = d

    def func(t):
        # f(t) = a.z - sag(a)   (sag's own transformation handled by sag_value)
        x, y, z = ox + ux * t, oy + uy * t, oz + uz * t
        return z - sag_value(sag, x, y, z)

    tol = min_step_size if not converge else 0.0
    t0 = np.zeros(np.broadcast(ox, ux).shape)
    t1 = t0 + 10 * min_step_size
    with np.errstate(invalid="ignore", divide="ignore"):
        f0 = func(t0)
        for _ in range(max_iterations):
            f1 = func(t1)
            df = f1 - f0
            step = np.where(df != 0, f1 * (t1 - t0) / df, 0.0)
            t0, f0 = t1, f1
            t1 = t1 - step
            scale = 4 * np.finfo(float).eps * np.maximum(np.abs(t1), 1.0) if converge else tol
            if np.all(~(np.abs(step) > scale)):
                break
    return t1


def sag_intercept(sag, rays: dict, converge: bool = False, generic: bool = False, extended: bool = False) -> dict:
    """
    ``sag.intercept(rays)``: rays moved to the surface, direction unchanged.
    ``generic=True`` forces the iterative ``AbstractSag.intercept`` for any sag
    (used to restate ``optika/sags/_tests/_abc_test.py:100-102``).

    ``extended=True`` evaluates the SAME closed-form expressions of the parabolic
    and conic sags in ``numpy.longdouble`` (80-bit on x86), polishes the root they
    select by Newton steps on the quadratic it solves (still in 80 bits), and rounds
    the path length to float64 at the end.  Those two reference formulas subtract nearly
    equal numbers for near-axial rays -- in float64 they carry a rounding noise of
    about ``eps * |2 f| / (ux^2 + uy^2)`` (1e-6 mm in the Newtonian example), which
    any 1-ulp change upstream re-samples -- so "the reference's answer" is only
    defined to that noise.  The extended evaluation is the reference's formula
    without that noise; parity tests compare against it and report the float64
    noise separately ("conditioning of the reference's closed forms").
    """
    name = _name(sag)
    t = getattr(sag, "transformation", None)
    if generic or name == "ToroidalSag":
        # optika/sags/_abc.py:76-107 (no transformation handling of its own: the
        # sag value function inverts the sag transformation internally)
        o = (rays["px"], rays["py"], rays["pz"])
        d = (rays["dx"], rays["dy"], rays["dz"])
        tt = _intercept_secant(sag, o, d, converge=converge)
        result = dict(rays)
        result["px"] = o[0] + d[0] * tt
        result["py"] = o[1] + d[1] * tt
        result["pz"] = o[2] + d[2] * tt
        return result

    r = _rays_transform(t, rays, inverse=True)
    ox, oy, oz = r["px"], r["py"], r["pz"]
    ux, uy, uz = r["dx"], r["dy"], r["dz"]
    o64, u64 = (ox, oy, oz), (ux, uy, uz)
    if extended and name in ("ParabolicSag", "ConicSag"):
        ox, oy, oz, ux, uy, uz = [np.asarray(v, dtype=np.longdouble) for v in (ox, oy, oz, ux, uy, uz)]
    with np.errstate(invalid="ignore", divide="ignore"):
        if name == "NoSag":
            # optika/sags/_flat.py:49-64
            tt = -oz / uz
        elif name == "SphericalSag":
            # optika/sags/_spherical.py:148-191
            rad = _f(sag.radius)
            px, py, pz = ox, oy, oz - rad
            up = ux * px + uy * py + uz * pz
            tt = -up - np.sign(rad * uz) * np.sqrt(up**2 - (px**2 + py**2 + pz**2 - rad**2))
        elif name == "ParabolicSag":
            # optika/sags/_parabolic.py:65-158
            f = _f(sag.focal_length)
            if extended:
                f = np.longdouble(f)
            tt = np.where(
                (ux**2 + uy**2) > 1e-10,
                (
                    -ox * ux - oy * uy + 2 * f * uz
                    - np.sign(f * uz)
                    * np.sqrt(
                        -((oy * ux) ** 2) - (ox * uy) ** 2
                        + 2 * oy * uy * (ox * ux - 2 * f * uz)
                        + 4 * f * (oz * (ux**2 + uy**2) - ox * ux * uz + f * uz**2)
                    )
                )
                / (ux**2 + uy**2),
                (ox**2 + oy**2 - 4 * f * oz) / (4 * f * uz),
            )
            if extended:
                # 80-bit evaluation still leaves ~eps_80 |2 f| / (ux^2 + uy^2) (1e-7 mm just above
                # the 1e-10 switch): polish the root the formula selected on its own equation,
                # z(t) = (x(t)^2 + y(t)^2) / (4 f), a quadratic in t with well-conditioned
                # coefficients.  Newton from 1e-7 away converges to rounding in two steps.
                qa = ux**2 + uy**2
                qb = 2 * (ox * ux + oy * uy) - 4 * f * uz
                qc = ox**2 + oy**2 - 4 * f * oz
                quadratic = (qa > 1e-10) & np.isfinite(tt)
                for _ in range(3):
                    step = (qa * tt**2 + qb * tt + qc) / (2 * qa * tt + qb)
                    tt = np.where(quadratic & np.isfinite(step), tt - step, tt)
        elif name == "ConicSag":
            # optika/sags/_conic.py:83-170
            radius, conic = _sag_radius_conic(sag)
            c = 1 / radius
            kp1 = 1 + conic
            if extended:
                c, kp1 = np.longdouble(c), np.longdouble(kp1)
            a = c * (np.square(ux) + np.square(uy) + kp1 * np.square(uz))
            b = 2 * (c * (ox * ux + oy * uy + kp1 * oz * uz) - uz)
            cc = c * (np.square(ox) + np.square(oy) + kp1 * np.square(oz)) - 2 * oz
            discriminant = np.square(b) - 4 * a * cc
            real = discriminant >= 0
            sqrt_discriminant = np.sqrt(np.where(real, discriminant, 0))
            degenerate = np.abs(a) < 1e-12
            denominator = np.where(degenerate, 1.0, 2 * a)
            t_linear = -cc / b

            def root(sign):
                t_ = np.where(degenerate, t_linear, (-b + sign * sqrt_discriminant) / denominator)
                x_, y_, z_ = ox + ux * t_, oy + uy * t_, oz + uz * t_
                r2 = np.square(x_) + np.square(y_)
                on_vertex_sheet = (z_ * (c * r2 - z_)) >= 0
                valid = real & on_vertex_sheet
                return np.where(valid, t_, np.inf)

            t_a = root(-1)
            t_b = root(+1)
            tt = np.where(np.abs(t_a) <= np.abs(t_b), t_a, t_b)
            if extended:
                # as for the parabola: polish the selected root on a t^2 + b t + c = 0
                quadratic = ~degenerate & np.isfinite(tt)
                for _ in range(3):
                    step = (a * tt**2 + b * tt + cc) / (2 * a * tt + b)
                    tt = np.where(quadratic & np.isfinite(step), tt - step, tt)
        elif name == "CylindricalSag":
            # optika/sags/_cylindrical.py:115-160, cross products written out with a = y-hat:
            # n x a = (-n_z, 0, n_x);  b = (0,0,r) - o;  b x a = (-b_z, 0, b_x)
            rad = _f(sag.radius)
            bx, by, bz = -ox, -oy, rad - oz
            ncx, ncz = -uz, ux
            n_cross_a_squared = ncx * ncx + ncz * ncz
            negative_b = ncx * (-bz) + ncz * bx
            b_squared = n_cross_a_squared * np.square(rad)
            four_ac = np.square(bx * ncx + bz * ncz)
            discriminant = b_squared - four_ac
            sgn = np.sign(rad * uz)
            tt = np.where(
                discriminant > 0,
                (negative_b - sgn * np.sqrt(discriminant)) / n_cross_a_squared,
                -oz / uz,
            )
        else:
            raise NotImplementedError(f"oracle: sag {name}")
        tt = np.asarray(tt, dtype=np.float64)
        (ox, oy, oz), (ux, uy, uz) = o64, u64
        r = dict(r)
        r["px"], r["py"], r["pz"] = ox + ux * tt, oy + uy * tt, oz + uz * tt
    return _rays_transform(t, r, inverse=False)


def sag_propagate(sag, rays: dict, converge: bool = False, extended: bool = False) -> dict:
    """``AbstractSag.propagate_rays``, ``optika/sags/_abc.py:109-122``."""
    result = sag_intercept(sag, rays, converge=converge, extended=extended)
    with np.errstate(invalid="ignore", over="ignore"):
        length = np.sqrt(
            np.square(result["px"] - rays["px"])
            + np.square(result["py"] - rays["py"])
            + np.square(result["pz"] - rays["pz"])
        )
        f = np.exp(-result["attenuation"] * length)
        result = dict(result)
        result["intensity"] = f * result["intensity"]
    return result


# ---------------------------------------------------------------------------
# rulings
# ---------------------------------------------------------------------------
def _vec3(v):
    return _f(v.x), _f(v.y), _f(v.z)


def ruling_vector(spacing, position, normal):
    """``spacing_(position, normal)`` -> kappa (kx, ky, kz)."""
    name = _name(spacing)
    x, y, z = position
    with np.errstate(invalid="ignore", divide="ignore"):
        if name == "ConstantRulingSpacing":
            # optika/rulings/_spacing.py:69-74
            g = _vec3(spacing.normal)
            c = _f(spacing.constant)
            one = np.ones(np.broadcast(x, y, z).shape)
            return c * g[0] * one, c * g[1] * one, c * g[2] * one
        if name == "Polynomial1dRulingSpacing":
            # optika/rulings/_spacing.py:109-128 (transformation applied FORWARDS)
            g = _vec3(spacing.normal)
            t = spacing.transformation
            if t is not None:
                x, y, z = transform_forward(t, x, y, z)
            s = x * g[0] + y * g[1] + z * g[2]
            result = 0.0
            for power, coefficient in spacing.coefficients.items():
                result = result + _f(coefficient) * (s**power)
            return result * g[0], result * g[1], result * g[2]
        if name == "HolographicRulingSpacing":
            # optika/rulings/_spacing.py:295-328
            x1 = _vec3(spacing.x1)
            x2 = _vec3(spacing.x2)
            w = _f(spacing.wavelength)
            d1 = 2 * float(bool(_f(spacing.is_diverging_1))) - 1
            d2 = 2 * float(bool(_f(spacing.is_diverging_2))) - 1
            nx, ny, nz = normal
            r1 = (x - x1[0], y - x1[1], z - x1[2])
            r2 = (x - x2[0], y - x2[1], z - x2[2])
            l1 = np.sqrt(r1[0] ** 2 + r1[1] ** 2 + r1[2] ** 2)
            l2 = np.sqrt(r2[0] ** 2 + r2[1] ** 2 + r2[2] ** 2)
            r1 = tuple(d1 * (c / l1) for c in r1)
            r2 = tuple(d2 * (c / l2) for c in r2)
            dr = tuple(a - b for a, b in zip(r1, r2))
            aq = (
                ny * dr[2] - nz * dr[1],
                nz * dr[0] - nx * dr[2],
                nx * dr[1] - ny * dr[0],
            )
            a = np.sqrt(aq[0] ** 2 + aq[1] ** 2 + aq[2] ** 2)
            q = tuple(c / a for c in aq)
            sp = w / a
            return (
                sp * (q[1] * nz - q[2] * ny),
                sp * (q[2] * nx - q[0] * nz),
                sp * (q[0] * ny - q[1] * nx),
            )
    raise NotImplementedError(f"oracle: ruling spacing {name}")


class ConstantRulingSpacing:
    """Local stand-in used when `Rulings.spacing` is a bare length."""

    def __init__(self, constant):
        self.constant = constant
        self.normal = _Vec(1.0, 0.0, 0.0)


class _Vec:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z


def _spacing_of(rulings):
    # optika/rulings/_rulings.py:156-168: bare spacing => constant along x-hat
    spacing = rulings.spacing
    if not _name(spacing).endswith("RulingSpacing"):
        spacing = ConstantRulingSpacing(spacing)
    return spacing


def incident_effective(rulings, rays: dict, normal) -> dict:
    """
    ``AbstractRulings.incident_effective``, ``optika/rulings/_rulings.py:170-204``
    with the kernel ``:107-128``:
    ``a + sign(a . n) * m * w * g / (n * d)``, ``d = |kappa|``, ``g = kappa / d``.
    """
    kappa = ruling_vector(_spacing_of(rulings), (rays["px"], rays["py"], rays["pz"]), normal)
    with np.errstate(invalid="ignore", divide="ignore"):
        d = np.sqrt(kappa[0] ** 2 + kappa[1] ** 2 + kappa[2] ** 2)
        g = tuple(k / d for k in kappa)
        m = _f(rulings.diffraction_order)
        w = rays["wavelength"]
        n = rays["index_refraction"]
        ax, ay, az = rays["dx"], rays["dy"], rays["dz"]
        ux, uy, uz = normal
        s = np.sign(ax * ux + ay * uy + az * uz)
        result = dict(rays)
        result["dx"] = ax + s * m * w * g[0] / (n * d)
        result["dy"] = ay + s * m * w * g[1] / (n * d)
        result["dz"] = az + s * m * w * g[2] / (n * d)
    return result


# ---------------------------------------------------------------------------
# materials
# ---------------------------------------------------------------------------
def is_mirror(material) -> bool:
    # optika/materials/_materials.py:114-116, 155-157, 453-455
    return _name(material) in ("Mirror", "MeasuredMirror", "MultilayerMirror")


def _passes_through(material) -> bool:
    # AbstractMultilayerMaterial: index and attenuation of the incoming ray (_multilayers.py:795-805)
    return _name(material) in ("MultilayerMirror", "MultilayerFilm")


def _oracle_layers(layers, wavelength):
    """Product-style layer objects -> the tuples of ``oracle.multilayer`` with ``n`` at the ray wavelengths."""
    from . import multilayer as orm

    out = []
    if layers is None:
        return out
    if not isinstance(layers, (list, tuple)):
        layers = [layers]
    for layer in layers:
        name = _name(layer)
        if name == "PeriodicLayerSequence":
            out.append(("periodic", _oracle_layers(list(layer.layers), wavelength), int(layer.num_periods)))
        elif name == "LayerSequence":
            out += _oracle_layers(list(layer.layers), wavelength)
        else:
            if layer.chemical is None:
                n = 1.0 + 0j  # optika/materials/_layers.py:218-227
            else:
                chemical = layer._chemical
                table = orm.load_nk(_nk_file(chemical.file_nk))  # optika/chemicals/_chemicals.py:116-142
                n = orm.interp_nk(wavelength / 1e-7, table)
            kind = 0 if layer.interface is None else layer.interface.kind
            width = 0.0 if layer.interface is None else _f(layer.interface.width)
            out.append((n, 0.0 if layer.thickness is None else _f(layer.thickness), kind, width))
    return out


def _nk_file(name: str) -> str:
    import os
    import pathlib

    roots = [pathlib.Path(p) for p in os.environ.get("OPTIKA_NK_PATH", "").split(os.pathsep) if p]
    roots.append(pathlib.Path(__file__).resolve().parent.parent / "optika_b200" / "data" / "nk")  # data, not code
    for root in roots:
        if (root / name).exists():
            return str(root / name)
    raise FileNotFoundError(name)


def index_refraction(material, rays: dict):
    name = _name(material)
    if name in ("Vacuum", "IdealSensorMaterial"):
        return np.ones_like(rays["wavelength"])  # _materials.py:95-99
    if is_mirror(material) or _passes_through(material):
        return rays["index_refraction"]  # _materials.py:135-139, _multilayers.py:795-799
    if name == "Glass":
        # optika/materials/_materials.py:428-438
        w2 = np.square(rays["wavelength"])
        n2 = 1 + (
            _f(material.b1) * w2 / (w2 - _f(material.c1))
            + _f(material.b2) * w2 / (w2 - _f(material.c2))
            + _f(material.b3) * w2 / (w2 - _f(material.c3))
        )
        return np.sqrt(n2)
    raise NotImplementedError(f"oracle: material {name}")


def attenuation(material, rays: dict):
    if is_mirror(material) or _passes_through(material):
        return rays["attenuation"]  # _materials.py:141-145, _multilayers.py:801-805
    return np.zeros_like(rays["wavelength"])  # _materials.py:101-105, 440-444


def _measured_interp(measured, wavelength):
    """
    ``na.interp(x=rays.wavelength, xp=wavelength, fp=efficiency)`` of a measured efficiency
    (``optika/materials/_materials.py:283-305``, ``optika/rulings/_rulings.py:291-313``);
    third-party ``na.interp`` assumed to be ``numpy.interp`` (linear, clamped ends).
    """
    inputs = measured.inputs
    xp = np.asarray(inputs.wavelength.ndarray, dtype=np.float64)
    fp = np.asarray(measured.outputs.ndarray, dtype=np.float64)
    if xp.ndim != 1:
        raise ValueError(f"wavelength must be one dimensional, got shape {xp.shape}")
    if fp.shape != xp.shape:
        raise NotImplementedError("oracle: select the configuration first (select_config)")
    order = np.argsort(xp)
    return np.interp(wavelength, xp[order], fp[order])


def material_efficiency(material, rays: dict, normal):
    """``material.efficiency(rays, normal)``."""
    name = _name(material)
    if name == "MeasuredMirror":
        return _measured_interp(material.efficiency_measured, rays["wavelength"])  # _materials.py:279-305
    if name in ("Vacuum", "IdealSensorMaterial", "Mirror", "Glass", "_FixedIndex"):
        return 1.0  # _materials.py:107-112, 147-152, 446-451
    if name in ("MultilayerMirror", "MultilayerFilm"):
        # optika/materials/_multilayers.py:839-866 (film: T.average, substrate None), 908-935 (mirror: R.average)
        from . import multilayer as orm

        w = rays["wavelength"]
        k = rays["attenuation"] * w / (4 * np.pi)
        n = rays["index_refraction"] + k * 1j
        cos = -(rays["dx"] * normal[0] + rays["dy"] * normal[1] + rays["dz"] * normal[2])
        stack = _oracle_layers(material.layers, w)
        substrate = None
        if name == "MultilayerMirror" and material.substrate is not None:
            (substrate,) = _oracle_layers([material.substrate], w)
        r_s, r_p, t_s, t_p = orm.multilayer_efficiency(w, cos, n, stack, substrate)
        return (r_s + r_p) / 2 if name == "MultilayerMirror" else (t_s + t_p) / 2
    raise NotImplementedError(f"oracle: efficiency of material {name}")


def rulings_efficiency(rulings, rays: dict, normal):
    """
    ``rulings.efficiency(rays, normal)``: the groove efficiencies of
    ``optika/rulings/_rulings.py`` (Magnusson & Gaylord 1978, Table 1), line by line,
    including ``direction - direction @ parallel_rulings`` (a scalar subtracted from a
    vector, ``:446-447``), the unsquared Bessel function of the sinusoidal profile
    (``:455``) and the ``+ i^2`` of the triangular profile (``:901``).
    """
    import scipy.special

    name = _name(rulings)
    if name == "Rulings":
        return 1.0  # :246-251
    if name == "MeasuredRulings":
        return _measured_interp(rulings.efficiency_measured, rays["wavelength"])  # :287-313
    with np.errstate(invalid="ignore", divide="ignore"):
        kappa = ruling_vector(_spacing_of(rulings), (rays["px"], rays["py"], rays["pz"]), normal)
        length = np.sqrt(kappa[0] ** 2 + kappa[1] ** 2 + kappa[2] ** 2)
        g = tuple(k / length for k in kappa)  # normal_rulings
        nx, ny, nz = normal
        p = (ny * g[2] - nz * g[1], nz * g[0] - nx * g[2], nx * g[1] - ny * g[0])  # normal.cross(normal_rulings)
        length = np.sqrt(p[0] ** 2 + p[1] ** 2 + p[2] ** 2)
        p = tuple(c / length for c in p)
        d = (rays["dx"], rays["dy"], rays["dz"])
        dp = d[0] * p[0] + d[1] * p[1] + d[2] * p[2]
        d = tuple(c - dp for c in d)
        wavelength = rays["wavelength"]
        cos_theta = -(d[0] * nx + d[1] * ny + d[2] * nz)
        depth = _f(rulings.depth)
        i = _f(rulings.diffraction_order)
        pi = np.pi
        if name == "SinusoidalRulings":  # :442-457
            gamma = pi * depth / (wavelength * cos_theta)
            return scipy.special.jv(i, 2 * gamma)
        if name == "SquareRulings":  # :590-614
            gamma = pi * (depth / (pi / 4)) / (wavelength * cos_theta)
            result = np.where(i % 2 == 0, 0, np.square(2 * np.sin(pi * gamma / 2) / (i * pi)) if i != 0 else 0)
            return np.where(i == 0, np.square(np.cos(pi * gamma / 2)), result)
        if name == "SawtoothRulings":  # :741-758
            gamma = pi * (depth / (pi / 2)) / (wavelength * cos_theta)
            return np.square(np.sin(pi * gamma) / (pi * (gamma + i)))
        if name == "TriangularRulings":  # :887-911
            gamma = pi * (depth / (np.square(pi) / 8)) / (wavelength * cos_theta)
            a = gamma / (np.square(pi * gamma / 2) + np.square(i))
            return np.where(
                i % 2 == 0,
                np.square(a * np.sin(np.square(pi) * gamma / 4)),
                np.square(a * np.cos(np.square(pi) * gamma / 4)),
            )
        if name == "RectangularRulings":  # :1046-1073
            a = 2 * pi * _f(rulings.ratio_duty)
            amplitude = pi / (2 * np.sqrt(2 * (1 - np.cos(a))))
            gamma = pi * (depth / amplitude) / (wavelength * cos_theta)
            b = np.square(np.sin(pi * gamma / np.sqrt(2 * (1 - np.cos(a)))))
            if i == 0:
                return 1 - ((2 * a / pi) - np.square(a / pi)) * b
            return (2 / np.square(i * pi)) * (1 - np.cos(i * a)) * b
    raise NotImplementedError(f"oracle: rulings {name}")


# How the surface operator evaluates Snell's law: None -- the NumPy expression below; "parallel" /
# "serial" -- its numba ``guvectorize`` twin (oracle/snell_numba.py), which is how the reference itself
# runs this one step (``_snells_law.py:294-302``: ``target="parallel"``).  Bit-identical results; set by
# the CPU arm of bench.py through :func:`use_numba_snell`.
_SNELL_MODE = None


def use_numba_snell(mode: str | None) -> str | None:
    """Select the Snell implementation of :func:`surface_propagate`; returns the mode in effect (None without numba)."""
    global _SNELL_MODE
    if mode is not None:
        try:
            from . import snell_numba  # noqa: F401
        except Exception:
            mode = None
    _SNELL_MODE = mode
    return mode


def snells_law(ax, ay, az, n1, n2, ux, uy, uz, mirror: bool):
    """
    Vector Snell's law, ``optika/materials/_snells_law.py:341-366``
    (the numba kernel body; ``|a|^2`` is NOT assumed to be 1, which matters after
    ``incident_effective``).
    """
    if _SNELL_MODE is not None:
        from . import snell_numba

        kernel = snell_numba.snells_law_parallel if _SNELL_MODE == "parallel" else snell_numba.snells_law_serial
        args = np.broadcast_arrays(*[np.asarray(v, dtype=np.float64) for v in (ax, ay, az, n1, n2, ux, uy, uz)])
        with np.errstate(invalid="ignore", divide="ignore"):
            return kernel(*args, bool(mirror))
    with np.errstate(invalid="ignore", divide="ignore"):
        a2 = ax * ax + ay * ay + az * az
        r = n1 / n2
        r2 = r * r
        au = ax * ux + ay * uy + az * uz
        au2 = au * au
        sgn = -np.copysign(1.0, au)
        d = -au + sgn * (2 * float(mirror) - 1) * np.sqrt(1 / r2 + au2 - a2)
        return r * (ax + d * ux), r * (ay + d * uy), r * (az + d * uz)


# ---------------------------------------------------------------------------
# apertures
# ---------------------------------------------------------------------------
def point_in_polygon(x, y, vx, vy):
    """
    ``na.geometry.point_in_polygon`` (third-party, named_arrays ~= 2.1; source not
    in /root/reference).  Restated as the even-odd crossing test with the
    published "is_inside_sm" rule set: a point exactly on an edge or vertex counts
    as inside.  Interior/exterior behaviour is pinned by
    ``optika/apertures/_apertures_test.py:62-80, 344-373``; behaviour ON edges is
    parity-unpinned, so tests enumerate edge rays via :func:`aperture_margin`.
    """
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    nv = len(vx)
    inside = np.zeros(x.shape, dtype=bool)
    on_edge = np.zeros(x.shape, dtype=bool)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(nv):
            x0, y0 = vx[i], vy[i]
            x1, y1 = vx[(i + 1) % nv], vy[(i + 1) % nv]
            # point-on-segment test
            cross = (x1 - x0) * (y - y0) - (y1 - y0) * (x - x0)
            within = (
                (np.minimum(x0, x1) <= x) & (x <= np.maximum(x0, x1))
                & (np.minimum(y0, y1) <= y) & (y <= np.maximum(y0, y1))
            )
            on_edge |= (cross == 0) & within
            # half-open crossing rule: the point is left of the edge, x < x0 + (y - y0) (x1 - x0) / (y1 - y0),
            # written without the division: the sign of the same cross product, oriented by the edge
            straddles = (y0 > y) != (y1 > y)
            inside ^= straddles & ((cross > 0) == (y1 > y0))
    return inside | on_edge


def _polygon_vertices(aperture):
    v = aperture.vertices
    vx = np.asarray(v.x.ndarray if hasattr(v.x, "ndarray") else v.x, dtype=float)
    vy = np.asarray(v.y.ndarray if hasattr(v.y, "ndarray") else v.y, dtype=float)
    return np.atleast_1d(vx), np.atleast_1d(vy)


def _aperture_local(aperture, x, y, z):
    t = getattr(aperture, "transformation", None)
    if t is not None:
        x, y, z = transform_inverse(t, x, y, z)
    return x, y, z


def _half_width(aperture):
    h = aperture.half_width
    if hasattr(h, "x") and hasattr(h, "y"):
        return _f(h.x), _f(h.y)
    return _f(h), _f(h)


def aperture_mask(aperture, x, y, z=0.0):
    """``aperture(position)`` -> bool mask."""
    name = _name(aperture)
    x, y, z = np.broadcast_arrays(*[np.asarray(a, dtype=float) for a in (x, y, z)])
    x, y, z = _aperture_local(aperture, x, y, z)
    active = bool(_f(aperture.active))
    inverted = bool(_f(aperture.inverted))
    with np.errstate(invalid="ignore", divide="ignore"):
        if name == "CircularAperture":
            # optika/apertures/_apertures.py:292-314
            mask = np.sqrt(np.square(x) + np.square(y)) <= _f(aperture.radius)
        elif name == "CircularSectorAperture":
            # optika/apertures/_apertures.py:438-480
            mask_radius = np.sqrt(np.square(x) + np.square(y)) <= _f(aperture.radius)
            a0, a1 = _f(aperture.angle_start), _f(aperture.angle_stop)
            angle = np.arctan2(y, x)
            angle_positive = angle % (+2 * np.pi)
            angle_negative = angle % (-2 * np.pi)
            mask_positive = (a0 < angle_positive) & (angle_positive < a1)
            mask_negative = (a0 < angle_negative) & (angle_negative < a1)
            mask = mask_radius & (mask_positive | mask_negative)
        elif name == "EllipticalAperture":
            # optika/apertures/_apertures.py:640-663
            mask = (
                np.square(x / _f(aperture.radius.x)) + np.square(y / _f(aperture.radius.y)) <= 1
            )
        elif name == "RectangularAperture":
            # optika/apertures/_apertures.py:941-968
            hx, hy = _half_width(aperture)
            mask = (-hx <= x) & (x <= hx) & (-hy <= y) & (y <= hy)
        elif hasattr(aperture, "vertices"):
            # optika/apertures/_apertures.py:739-778
            if not active:
                return np.ones(x.shape, dtype=bool)
            vx, vy = _polygon_vertices(aperture)
            mask = point_in_polygon(x, y, vx, vy)
        else:
            raise NotImplementedError(f"oracle: aperture {name}")
    if inverted:
        mask = ~mask
    if not active:
        mask = np.ones_like(mask)
    return mask


def aperture_margin(aperture, x, y, z=0.0):
    """
    Distance (same units as the aperture) from each point to the nearest aperture
    edge.  Test infrastructure for the north-star rule "masks bit-exact except for
    rays within tolerance of an aperture edge, which are enumerated".
    """
    name = _name(aperture)
    x, y, z = np.broadcast_arrays(*[np.asarray(a, dtype=float) for a in (x, y, z)])
    x, y, z = _aperture_local(aperture, x, y, z)
    with np.errstate(invalid="ignore", divide="ignore"):
        if name in ("CircularAperture", "CircularSectorAperture"):
            m = np.abs(np.sqrt(x * x + y * y) - _f(aperture.radius))
            if name == "CircularSectorAperture":
                angle = np.arctan2(y, x)
                r = np.sqrt(x * x + y * y)
                for a in (_f(aperture.angle_start), _f(aperture.angle_stop)):
                    da = np.abs((angle - a + np.pi) % (2 * np.pi) - np.pi)
                    m = np.minimum(m, r * da)
            return m
        if name == "EllipticalAperture":
            a, b = _f(aperture.radius.x), _f(aperture.radius.y)
            return np.abs(np.sqrt((x / a) ** 2 + (y / b) ** 2) - 1) * min(abs(a), abs(b))
        if name == "RectangularAperture":
            hx, hy = _half_width(aperture)
            return np.minimum(np.abs(np.abs(x) - hx), np.abs(np.abs(y) - hy))
        vx, vy = _polygon_vertices(aperture)
        m = np.full(x.shape, np.inf)
        nv = len(vx)
        for i in range(nv):
            x0, y0 = vx[i], vy[i]
            x1, y1 = vx[(i + 1) % nv], vy[(i + 1) % nv]
            ex, ey = x1 - x0, y1 - y0
            l2 = ex * ex + ey * ey
            s = np.clip(((x - x0) * ex + (y - y0) * ey) / l2, 0, 1) if l2 > 0 else 0 * x
            m = np.minimum(m, np.sqrt((x - x0 - s * ex) ** 2 + (y - y0 - s * ey) ** 2))
        return m


def is_angular(aperture) -> bool:
    """Stand-in for the unit test of ``optika/apertures/_apertures.py:93-99``."""
    return bool(getattr(aperture, "angular", False))


def aperture_clip(aperture, rays: dict) -> dict:
    """``AbstractAperture.clip_rays``, ``optika/apertures/_apertures.py:82-102``."""
    if is_angular(aperture):
        mask = aperture_mask(aperture, rays["dx"], rays["dy"], rays["dz"])
    else:
        mask = aperture_mask(aperture, rays["px"], rays["py"], rays["pz"])
    result = dict(rays)
    result["unvignetted"] = rays["unvignetted"] & mask
    return result


# ---------------------------------------------------------------------------
# the surface operator and the sequential loop
# ---------------------------------------------------------------------------
def surface_propagate(surface, rays: dict, converge: bool = False, extended: bool = False) -> dict:
    """``AbstractSurface.propagate_rays``, ``optika/surfaces.py:123-198``, step by step."""
    sag = surface.sag
    material = surface.material
    aperture = surface.aperture
    rulings = surface.rulings
    transformation = surface.transformation

    if transformation is not None:  # :141-142
        rays = _rays_transform(transformation, rays, inverse=True)

    rays_1 = sag_propagate(sag, rays, converge=converge, extended=extended)  # :144
    normal = sag_normal(sag, rays_1["px"], rays_1["py"], rays_1["pz"])  # :146-148

    if rulings is not None:  # :150-154
        rays_1 = incident_effective(rulings, rays_1, normal)

    wavelength_1 = rays_1["wavelength"]  # :156-159
    n1 = rays_1["index_refraction"]
    n2 = index_refraction(material, rays_1)  # :162
    with np.errstate(invalid="ignore", divide="ignore"):
        r = n1 / n2  # :163
        wavelength_2 = wavelength_1 / r  # :165
    bx, by, bz = snells_law(  # :167-173
        rays_1["dx"], rays_1["dy"], rays_1["dz"], n1, n2, normal[0], normal[1], normal[2],
        is_mirror(material),
    )
    efficiency = material_efficiency(material, rays_1, normal)  # :175
    if rulings is not None:  # :176-177
        efficiency = efficiency * rulings_efficiency(rulings, rays_1, normal)
    rays_2 = dict(rays_1)  # :182-190
    rays_2["wavelength"] = wavelength_2
    rays_2["dx"], rays_2["dy"], rays_2["dz"] = bx, by, bz
    rays_2["intensity"] = rays_1["intensity"] * efficiency  # :179
    rays_2["attenuation"] = attenuation(material, rays_1)  # :180
    rays_2["index_refraction"] = n2 + 0 * n1

    if aperture is not None:  # :192-193
        rays_2 = aperture_clip(aperture, rays_2)

    if transformation is not None:  # :195-196
        rays_2 = _rays_transform(transformation, rays_2, inverse=False)

    return rays_2


def propagate_rays(surfaces, rays: dict, converge: bool = False, extended: bool = False) -> dict:
    """``optika.propagators.propagate_rays``, ``optika/propagators.py:19-41``."""
    for surface in surfaces:
        rays = surface_propagate(surface, rays, converge=converge, extended=extended)
    return rays


def accumulate_rays(surfaces, rays: dict, converge: bool = False, extended: bool = False) -> dict:
    """``optika.propagators.accumulate_rays``, ``optika/propagators.py:44-73`` (new leading axis)."""
    result = []
    for surface in surfaces:
        rays = surface_propagate(surface, rays, converge=converge, extended=extended)
        result.append(rays)
    return {k: np.stack([r[k] for r in result]) for k in result[0]}


# ---------------------------------------------------------------------------
# direction cosines <-> angles
# ---------------------------------------------------------------------------
def direction(ax, ay):
    """``optika.direction``, ``optika/_util.py:41-73``."""
    return -np.cos(ay) * np.sin(ax), -np.sin(ay), np.cos(ay) * np.cos(ax)


def angles(dx, dy, dz):
    """``optika.angles``, ``optika/_util.py:76-97`` (radians)."""
    length = np.sqrt(dx * dx + dy * dy + dz * dz)
    return -np.arctan2(dx, dz), -np.arcsin(dy / length)
