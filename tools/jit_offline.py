#!/usr/bin/env python
"""
Compile, with nvcc and WITHOUT a GPU, the translation unit that jit.cu would hand to NVRTC for
one of the BASELINE systems, and print ptxas' register / spill report and the SASS size:

    python tools/jit_offline.py cfg3 [dense|image|grid|groups]

The walk is generated here from the lowered surface table the same way jit.cu::jit_source does
(keep the two in step); flags that only the library computes (OPTK_F_TRANSLATION_ONLY is set by
optk_system_create) are reproduced below.
"""
import pathlib
import subprocess
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

from optika_b200 import _lib as L, _lowering
import configs

POLYGON, POLYNOMIAL = L.APERTURE_POLYGON, L.RULING_POLYNOMIAL


def convex_flags(vx, vy):
    """OPTK_F_APERTURE_CONVEX / _CLOCKWISE as api.cu::classify_polygon sets them"""
    n = len(vx)
    if n < 3:
        return 0
    bound = max(np.abs(vx).max(), np.abs(vy).max())
    area2 = sum(vx[i] * vy[(i + 1) % n] - vx[(i + 1) % n] * vy[i] for i in range(n))
    if not bound > 0 or area2 == 0:
        return 0
    o = 1.0 if area2 > 0 else -1.0
    for i in range(n):
        j = (i + 1) % n
        ex, ey = vx[j] - vx[i], vy[j] - vy[i]
        for k in range(n):
            if k in (i, j):
                continue
            if not o * (ex * (vy[k] - vy[i]) - ey * (vx[k] - vx[i])) > 1e-9 * bound * bound:
                return 0
    return 0x4000 | (0x8000 if o < 0 else 0)


def walk_source(table, n_surface):
    lines = []
    for k in range(n_surface):
        S = table[k]
        flags = S.flags
        r = np.array(list(S.transform.r)).reshape(3, 3)
        if (flags & L.F_TRANSFORM) and np.array_equal(r, np.eye(3)):
            flags |= 0x400  # OPTK_F_TRANSLATION_ONLY
        if S.aperture_kind == POLYGON:
            flags |= convex_flags(np.array(S.vertices_x[: S.n_vertices]), np.array(S.vertices_y[: S.n_vertices]))
        eff = S.material_efficiency != 0 or S.ruling_profile != 0
        nv = S.n_vertices if S.aperture_kind == POLYGON else 0
        nc = S.n_coeff if S.ruling_kind == POLYNOMIAL else 0
        pw = [S.ruling_power[j] if j < nc else 0 for j in range(8)]
        args = [S.sag_kind, S.material_kind, S.ruling_kind, S.aperture_kind, flags, nv, nc, *pw]
        lines.append(
            f"    surface_full<2, {'true' if eff else 'false'}, FixedKinds<{', '.join(map(str, args))}>>"
            f"(P.surf[{k}], r, newton_iterations, state);"
        )
    return "\n".join(lines)


def main():
    name = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "dense"
    system = {
        "cfg1": lambda: configs.newtonian(100, 100, 128),
        "cfg2": lambda: configs.spherical_grating(100, 100, 1, 2048),
        "cfg3": lambda: configs.toroidal_vls(100, 100, 1),
    }[name]()
    surfaces = system.surfaces_all
    table, _ = _lowering.lower_system(surfaces, stages=L.STAGE_ALL)
    dense, vec, image, grid = {
        "dense": (1, 1, 0, 0), "image": (0, 0, 1, 0), "grid": (0, 0, 1, 1), "groups": (0, 0, 1, 0),
    }[mode]
    minb = 3
    b = lambda v: "true" if v else "false"  # noqa: E731
    layout = ""
    if mode in ("image", "groups"):
        # a separable grid as image_rays passes it: axes (field_x, field_y | pupil_x, pupil_y); wavelength and
        # direction vary along the leading (per-CTA) axes, position x / y along one trailing axis each, no mask
        lo = (1 << (1 * 8 + 2)) | (1 << (2 * 8 + 3)) | (1 << (4 * 8 + 0)) | (1 << (4 * 8 + 1)) | (1 << (5 * 8 + 0)) \
            | (1 << (5 * 8 + 1)) | (1 << (6 * 8 + 0)) | (1 << (6 * 8 + 1))
        layout = (
            "#define OPTK_JIT_LAYOUT 1\n#define OPTK_JIT_N_AXES 4\n#define OPTK_JIT_FIRST 2\n#define OPTK_JIT_HAS_MASK 0\n"
            f"#define OPTK_JIT_VARIES(f, a) ((((f) < 8 ? {hex(lo)}ULL >> (((f) & 7) * 8 + (a)) : 0x0ULL >> (((f) & 7) * 8 + (a))) & 1) != 0)\n"
            "#define OPTK_JIT_VARIES_OUTER(f) (((0x70u >> (f)) & 1) != 0)\n"
        )
    if mode == "groups":  # per-group accumulators instead of detector pixels (optk_image_t.group_size)
        layout = "#define OPTK_JIT_GROUPS 1\n" + layout
    if mode in ("image", "grid") and "--no-flags" not in sys.argv:
        # what launch_trace compiles in for a sensor image with a known range, linspace pixel edges, one spectral
        # bin and the planes flux / moment_real / counts; for the grid: object at infinity, packed angular cells,
        # jitter, both weights, no frame
        layout = "#define OPTK_JIT_IMAGE_FLAGS 0x7f\n" + layout
        if mode == "grid":
            layout = "#define OPTK_JIT_GRID_FLAGS 0x37\n" + layout
    src = layout + f"""#define OPTK_JIT_WALK 1
#include "trace_impl.cuh"
namespace optk {{
__device__ __forceinline__ void optk_jit_walk(const TraceParams& P, Ray (&r)[2], unsigned& newton_iterations,
                                              WalkState& state) {{
{walk_source(table, len(surfaces))}
}}
}}  // namespace optk
extern "C" __global__ void __launch_bounds__(256, {minb}) optk_jit_kernel(const __grid_constant__ optk::TraceParams P) {{
    optk::trace_body<2, true, {b(dense)}, {b(vec)}, false, {b(image)}, {grid}, false, 1>(P);
}}
"""
    print(src)
    with tempfile.TemporaryDirectory() as d:
        cu = pathlib.Path(d) / "optk_jit.cu"
        cu.write_text(src)
        out = pathlib.Path(d) / "optk_jit.cubin"
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-cubin", "-lineinfo",
               "-Xptxas", "-v", "--expt-relaxed-constexpr", f"-I{ROOT / 'optika_b200' / 'csrc'}", f"-I{ROOT / 'include'}",
               str(cu), "-o", str(out)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(r.stdout, r.stderr)
        if r.returncode == 0:
            sass = subprocess.run(["cuobjdump", "-sass", str(out)], capture_output=True, text=True).stdout
            ops = [ln.split()[1].rstrip(";") for ln in sass.splitlines() if ln.strip().startswith("/*") and len(ln.split()) > 2 and ln.split()[1][0].isalpha()]
            import collections
            c = collections.Counter(o.split(".")[0] for o in ops)
            print("SASS instructions:", len(ops), dict(c.most_common(14)))
            if "--sass" in sys.argv:
                print("=== SASS ===")
                print(sass)


if __name__ == "__main__":
    main()
