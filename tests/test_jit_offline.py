"""
The translation unit that ``csrc/jit.cu`` hands to NVRTC at run time, compiled here with nvcc and
without a GPU (``tools/jit_offline.py`` generates it from the lowered surface table the same way):
a guard against breaking the compile-time specialised paths of ``trace_impl.cuh`` (FixedKinds,
the inlined rarer element kinds, the input-layout and group-accumulator macros), which otherwise
only a GPU box would notice.
"""

import pathlib
import re
import shutil
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


# static size of the kernel body (SASS instructions for the two rays of a thread) after round 2's diets, with 10 %
# of slack: round 1 had 2623 for cfg 3 dense and 2723 for cfg 2 grid
SASS_BUDGET = {("cfg3", "dense"): 2000, ("cfg1", "image"): 2400, ("cfg2", "grid"): 1820, ("cfg1", "groups"): 2670}


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
@pytest.mark.parametrize("config,mode", [("cfg3", "dense"), ("cfg1", "image"), ("cfg2", "grid"), ("cfg1", "groups")])
def test_specialised_translation_unit_compiles_for_sm_100a(config, mode):
    out = subprocess.run(
        [sys.executable, str(ROOT / "tools" / "jit_offline.py"), config, mode], capture_output=True, text=True, timeout=600
    )
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-2000:]
    assert " error" not in text and "error:" not in text, text[-2000:]
    m = re.search(r"Used (\d+) registers", text)
    assert m, text[-2000:]
    assert int(m.group(1)) <= 80  # three 256-thread CTAs per SM
    spills = re.search(r"(\d+) bytes spill stores", text)
    assert spills and int(spills.group(1)) <= 64
    assert "FixedKinds<" in text and "SASS instructions:" in text
    size = int(re.search(r"SASS instructions: (\d+)", text).group(1))
    assert size <= SASS_BUDGET[(config, mode)], size
