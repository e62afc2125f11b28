"""
Optical constants of chemicals.

Mirrors ``optika.chemicals.Chemical`` (``optika/chemicals/_chemicals.py:67-144``):
the complex index of refraction ``n + ik`` is a linear interpolation of an
IMD/Windt ``.nk`` table (wavelength in Angstrom, ``n``, ``k``; ``;`` comment lines).
The reference re-reads the file on every call; here each table is parsed once
and cached, and the interpolation is done on the host *once per wavelength
grid* before the multilayer kernel is launched.

Tables are searched in ``$OPTIKA_NK_PATH`` (``os.pathsep``-separated), then in
``optika_b200/data/nk`` (Si, SiO2, SiC, Cr, Mo are bundled).
"""

from __future__ import annotations
import dataclasses
import functools
import os
import pathlib
import numpy as np
from . import named as na
from . import units as u

__all__ = ["AbstractChemical", "Chemical"]

_PATH_BUNDLED = pathlib.Path(__file__).parent / "data" / "nk"


def _search_path() -> list[pathlib.Path]:
    result = []
    env = os.environ.get("OPTIKA_NK_PATH")
    if env:
        result += [pathlib.Path(p) for p in env.split(os.pathsep) if p]
    result.append(_PATH_BUNDLED)
    return result


@functools.lru_cache(maxsize=None)
def _load_table(name: str) -> tuple[np.ndarray, np.ndarray]:
    for directory in _search_path():
        file = directory / name
        if file.exists():
            skip = 0
            with open(file, "r") as f:
                for line in f:
                    if line.startswith(";"):
                        skip += 1
                    else:
                        break
            w, n, k = np.loadtxt(file, skiprows=skip, unpack=True)
            return w * u.AA, n + 1j * k
    raise FileNotFoundError(
        f"optical-constant table {name!r} not found in {[str(p) for p in _search_path()]}"
    )


class AbstractChemical:
    pass


@dataclasses.dataclass(eq=False)
class Chemical(AbstractChemical):
    """A chemical identified by its empirical `formula` (``_chemicals.py:176-275``)."""

    formula: str = None
    is_amorphous: bool = False
    table: None | str = None

    @property
    def shape(self) -> dict[str, int]:
        return {}

    @property
    def file_nk(self) -> str:
        # optika/chemicals/_chemicals.py:84-98: "[a-]<formula>[_<table>].nk"
        file = f"{self.formula}"
        if self.table is not None:
            file = f"{file}_{self.table}"
        if self.is_amorphous:
            file = f"a-{file}"
        return f"{file}.nk"

    def n(self, wavelength) -> na.ScalarArray:
        """Complex index of refraction at `wavelength` (engine length units, mm)."""
        w = na.as_named_array(u.length(wavelength))
        wp, fp = _load_table(self.file_nk)
        return na.ScalarArray(np.interp(w.ndarray, wp, fp), w.axes)
