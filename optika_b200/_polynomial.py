"""
Least-squares polynomial fits over named axes: the stand-in for the third-party
``named_arrays.PolynomialFitFunctionArray.from_degree`` that the reference's distortion and
vignetting models are built on (``optika/distortion/_distortion.py:383-411``,
``optika/radiometry/_vignetting.py:161-172``).  The source of ``named_arrays`` is not part of the
reference checkout, so the behaviour assumed here is stated explicitly ("parity unpinned",
DESIGN.md): every monomial of the (centred) input components up to the total `degree`, fitted
by linear least squares over all logical axes of the inputs at the points selected by `where`;
leading axes that are not axes of the inputs (e.g. configuration axes of the outputs) get
independent fits.  What the reference's own tests pin -- a linear map is reproduced and inverted to
1e-9 deg (``optika/distortion/_distortion_test.py:40-44``) -- holds for any such fit.

Host code: the sums it consumes come from the device (``SequentialSystem.pupil_moments``), the
fit itself is a few hundred points by a dozen coefficients.
"""

from __future__ import annotations
import dataclasses
import itertools
import numpy as np
from . import named as na

__all__ = ["PolynomialFit"]


def _exponents(n_inputs: int, degree: int) -> list[tuple[int, ...]]:
    """All exponent tuples with total degree <= `degree`, constant term first, graded order."""
    out = []
    for total in range(degree + 1):
        for e in itertools.product(range(total + 1), repeat=n_inputs):
            if sum(e) == total:
                out.append(e)
    return out


@dataclasses.dataclass(eq=False)
class PolynomialFit:
    """
    ``outputs[k] ~ sum_e c[k, e] prod_j (inputs[j] - center[j]) ** e[j]``.

    `inputs`: named arrays (the components of the independent variable, e.g. wavelength, field x,
    field y), `outputs`: named arrays fitted independently with the same design matrix,
    `axes`: the logical axes the calibration points are spread over (default: all axes of the
    inputs), `where`: mask of the points that take part.
    """

    inputs: tuple
    outputs: tuple
    degree: int = 1
    center: None | tuple = None
    where: object = True
    axes: None | tuple = None

    def __post_init__(self):
        self.inputs = tuple(na.as_named_array(a) for a in self.inputs)
        self.outputs = tuple(na.as_named_array(a) for a in self.outputs)
        shape_in = na.broadcast_shapes(*[a.shape for a in self.inputs])
        if self.axes is None:
            self.axes = tuple(shape_in)
        full = na.broadcast_shapes(shape_in, na.shape(self.where), *[a.shape for a in self.outputs])
        self._batch = {ax: n for ax, n in full.items() if ax not in self.axes}
        self._points = {ax: full[ax] for ax in self.axes}
        order = dict(self._batch, **self._points)
        dims = tuple(order.values())
        n_batch = int(np.prod(list(self._batch.values()), dtype=np.int64)) if self._batch else 1

        def flat(a, dtype=float):
            return np.broadcast_to(na.aligned(na.as_named_array(a), order), dims).astype(dtype).reshape(n_batch, -1)

        x = [flat(a) for a in self.inputs]
        w = flat(self.where, bool)
        if self.center is None:
            self.center = tuple(
                na.ScalarArray(np.array([xi[b][w[b]].mean() if w[b].any() else 0.0 for b in range(n_batch)]).reshape(
                    tuple(self._batch.values())), tuple(self._batch)) for xi in x
            )
        c = [np.broadcast_to(na.aligned(na.as_named_array(ci), self._batch), tuple(self._batch.values())).reshape(n_batch)
             for ci in self.center]
        # scale every component to O(1) so that the normal equations of a wavelength in millimetres
        # next to an angle in radians stay well conditioned
        self._scale = np.ones((n_batch, len(x)))
        for j, xi in enumerate(x):
            for b in range(n_batch):
                if w[b].any():
                    s = np.max(np.abs(xi[b][w[b]] - c[j][b]))
                    self._scale[b, j] = s if s > 0 else 1.0
        self.exponents = _exponents(len(x), self.degree)
        self.coefficients = np.zeros((len(self.outputs), n_batch, len(self.exponents)))
        for b in range(n_batch):
            if not w[b].any():
                continue
            design = self._design([(xi[b][w[b]] - c[j][b]) / self._scale[b, j] for j, xi in enumerate(x)])
            for k, out in enumerate(self.outputs):
                y = flat(out)[b][w[b]]
                self.coefficients[k, b] = np.linalg.lstsq(design, y, rcond=None)[0]
        self._center_flat = np.stack(c, axis=1) if c else np.zeros((n_batch, 0))

    def _design(self, centred: list) -> np.ndarray:
        cols = []
        for e in self.exponents:
            col = np.ones_like(centred[0]) if centred else np.ones(1)
            for xj, ej in zip(centred, e):
                if ej:
                    col = col * xj ** ej
            cols.append(col)
        return np.stack(cols, axis=-1)

    @property
    def coefficient_names(self) -> list[str]:
        return ["*".join(f"x{j}^{e}" for j, e in enumerate(es) if e) or "1" for es in self.exponents]

    def __call__(self, *inputs) -> tuple:
        """The fitted outputs at new inputs (named arrays; batch axes broadcast by name)."""
        inputs = tuple(na.as_named_array(a) for a in inputs)
        shape_ = na.broadcast_shapes(self._batch, *[a.shape for a in inputs])
        order = dict(self._batch, **{ax: n for ax, n in shape_.items() if ax not in self._batch})
        dims = tuple(order.values())
        n_batch = int(np.prod(list(self._batch.values()), dtype=np.int64)) if self._batch else 1
        x = [np.broadcast_to(na.aligned(a, order), dims).astype(float).reshape(n_batch, -1) for a in inputs]
        results = [np.empty((n_batch, x[0].shape[1])) for _ in self.outputs]
        for b in range(n_batch):
            design = self._design([(xi[b] - self._center_flat[b, j]) / self._scale[b, j] for j, xi in enumerate(x)])
            for k in range(len(self.outputs)):
                results[k][b] = design @ self.coefficients[k, b]
        return tuple(na.ScalarArray(r.reshape(dims), tuple(order)) for r in results)

    @property
    def predictions(self) -> tuple:
        return self(*self.inputs)
