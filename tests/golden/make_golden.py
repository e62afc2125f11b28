"""
Regenerate the golden fixtures in this directory from the read-only reference
checkout (``/root/reference``).  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot be imported here (``named_arrays``/``astropy`` are absent),
so the fixtures are the reference's own *data files*, repacked losslessly:

* ``imd_<name>.npz`` -- the IMD golden reflectivity/transmissivity tables used by
  ``optika/materials/_tests/test_multilayers.py:178-287`` (``_data/*.txt``):
  columns as float64 arrays, the structure description kept as metadata.
* ``optika_b200/data/nk/<formula>.nk`` -- the optical-constant tables
  (wavelength [Angstrom], n, k) that ``optika/chemicals/_chemicals.py:101-144``
  interpolates, copied verbatim for the five chemicals the golden tests and the
  BASELINE multilayer configuration use (Si, SiO2, SiC, Cr, Mo).

Nothing in ``tests/`` or ``bench.py`` reads ``/root/reference`` at run time.
"""

import pathlib
import shutil
import numpy as np

REFERENCE = pathlib.Path("/root/reference/optika")
HERE = pathlib.Path(__file__).parent
NK_OUT = HERE.parent.parent / "optika_b200" / "data" / "nk"

IMD = ["Si", "SiO2", "SiO2_100A", "SiC_Cr", "SiC_Cr_Rough", "SiO2_rough"]
NK = ["Si", "SiO2", "SiC", "Cr", "Mo"]


def read_imd(file: pathlib.Path):
    header = []
    with open(file, "r") as f:
        for line in f:
            if line.startswith(";"):
                header.append(line.rstrip("\n"))
            else:
                break
    data = np.genfromtxt(file, skip_header=len(header), unpack=True)
    return "\n".join(header), data


def main():
    for name in IMD:
        header, data = read_imd(REFERENCE / "materials" / "_tests" / "_data" / f"{name}.txt")
        np.savez_compressed(
            HERE / f"imd_{name}.npz",
            header=np.array(header),
            wavelength_angstrom=data[0],
            columns=data[1:],
        )
        print(name, data.shape)
    NK_OUT.mkdir(parents=True, exist_ok=True)
    for name in NK:
        shutil.copyfile(REFERENCE / "chemicals" / "nk" / f"{name}.nk", NK_OUT / f"{name}.nk")
        (NK_OUT / f"{name}.nk").chmod(0o644)


if __name__ == "__main__":
    main()
