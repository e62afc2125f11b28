"""
Round-2 host behaviour that needs no GPU: the reference arm of ``bench.py`` prints a valid line, the
defaults follow the reference (normalised grids), edited surfaces are re-lowered, digests are stable.
"""

import json
import pathlib
import subprocess
import sys

import numpy as np
import pytest

import configs
from optika_b200 import _lowering, named as na, units as u

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_valid_json_line():
    done = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-rays", "40000"],
        capture_output=True, text=True, timeout=600, cwd=str(ROOT),
    )
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [ln for ln in done.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "ray-surface intercepts/sec" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["unit"] == "intercepts/s" and line["dtype"] == "f64"
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and "numba" in cpu["sample"]
    assert line["e2e"] == dict(value=line["value"], unit="intercepts/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_trace_entry_points_default_to_normalised_grids_like_the_reference():
    import inspect
    from optika_b200.systems import SequentialSystem

    # optika/systems/_sequential.py:843-844, 936-937
    for name in ("raytrace", "rayfunction", "pupil_moments", "image_rays", "distortion", "vignetting", "area_effective"):
        parameters = inspect.signature(getattr(SequentialSystem, name)).parameters
        assert parameters["normalized_field"].default is True, name
        assert parameters["normalized_pupil"].default is True, name
    assert configs.PHYSICAL == dict(normalized_field=False, normalized_pupil=False)


def test_digest_follows_every_edit_and_nothing_else():
    system = configs.misaligned_telescope(num_tilt=3)
    first = _lowering.fingerprint(system.surfaces_all)
    assert first == _lowering.fingerprint(system.surfaces_all) == _lowering.fingerprint(configs.misaligned_telescope(num_tilt=3).surfaces_all)
    system.surfaces[2].sag.focal_length = -201.0 * u.mm
    second = _lowering.fingerprint(system.surfaces_all)
    assert second != first
    angle = system.surfaces[2].transformation.transformations[0].angle
    moved = np.array(angle.ndarray)
    moved[1] += 1e-9  # one element of a named array, deep inside the transformation list
    system.surfaces[2].transformation.transformations[0].angle = na.ScalarArray(moved, angle.axes)
    assert _lowering.fingerprint(system.surfaces_all) != second
    system.surfaces[3].aperture.inverted = True
    assert _lowering.fingerprint(system.surfaces_all) not in (first, second)


def test_edited_surfaces_are_lowered_again_and_unchanged_ones_reuse_the_handle():
    system = configs.newtonian()
    try:
        compiled = system._compiled
    except Exception as e:  # needs liboptk.so, not a GPU
        pytest.skip(f"liboptk.so unavailable: {e}")
    assert system._compiled is compiled and system._compiled_local is system._compiled_local
    assert system._compiled_local is not compiled
    system.surfaces[2].sag.focal_length = -250.0 * u.mm
    again = system._compiled
    assert again is not compiled and again.table[3].sag[0] == -250.0
    system.sensor.num_pixel = na.Cartesian2dVectorArray(64, 32)
    assert system._compiled is not again  # the sensor's aperture follows its pixel grid
    key = _lowering.table_key(system._compiled.table)
    assert key == _lowering.table_key(_lowering.lower_system(system.surfaces_all)[0])


def test_surfaces_without_named_axes_are_lowered_once_per_system():
    system = configs.misaligned_telescope(num_tilt=4)
    table, shape_ = _lowering.lower_system(system.surfaces_all)
    assert shape_ == {"misalign": 4} and len(table) == 4 * 6
    n = 6
    for k in (0, 1, 2, 4, 5):  # every surface but the tilted primary is the same record in all configurations
        assert all(bytes(table[c * n + k]) == bytes(table[k]) for c in range(4))
    assert len({bytes(table[c * n + 3]) for c in range(4)}) == 4
