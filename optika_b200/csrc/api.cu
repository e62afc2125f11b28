// C ABI of liboptk (see include/optk.h).  Host-side glue only: argument
// checking, parameter packing, launch, and the host-pointer streaming path.
#include "common.cuh"
#include "bin.cuh"
#include "params.cuh"
#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in Nsight when a tool is attached, cost nothing otherwise

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace optk {

// ---- errors -------------------------------------------------------------------
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return OPTK_ERR_CUDA;
}

}  // namespace optk

using namespace optk;

struct optk_system {
    int32_t n_surface;
    int32_t n_config;
    std::vector<optk_surface_t> table;
    double* lut_device = nullptr;  // every efficiency table of the system, one allocation
    int lut_ordinal = -1;          // the device that allocation lives on
};

// Launches go to the device that owns the caller's stream, whatever device happens to be current in
// the calling thread (a process that drives several GPUs passes `device=` per call): the current
// device is switched for the duration of the entry point and restored.  NULL (the legacy default
// stream) keeps the current device.
struct DeviceScope {
    int previous = -1;
    bool switched = false;
    bool ranged = false;
    // an NVTX range per entry point (the C-ABI boundary is where a profile of the reference's Python would
    // want its markers: SURVEY.md section 5, tracing / profiling)
    DeviceScope(void* stream, const char* name) : DeviceScope(stream) {
        nvtxRangePushA(name);
        ranged = true;
    }
    explicit DeviceScope(void* stream) {
        int wanted = -1;
        if (!stream) return;
        if (cudaStreamGetDevice((cudaStream_t)stream, &wanted) != cudaSuccess || cudaGetDevice(&previous) != cudaSuccess) {
            cudaGetLastError();  // not a stream of this process: let the launch report it
            return;
        }
        if (wanted != previous && cudaSetDevice(wanted) == cudaSuccess) switched = true;
    }
    ~DeviceScope() {
        if (ranged) nvtxRangePop();
        if (switched) cudaSetDevice(previous);
    }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};

// ---- scratch cache for the host-pointer path ------------------------------------
namespace {

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    int ensure(size_t n) {
        if (n <= bytes) return OPTK_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&ptr, n);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
            return OPTK_ERR_NOMEM;
        }
        bytes = n;
        return OPTK_OK;
    }
};

const int kSlots = 3;

struct HostPipeline {
    std::mutex mutex;
    cudaStream_t stream[kSlots] = {nullptr, nullptr, nullptr};
    DeviceBuffer in[kSlots], out[kSlots];
    DeviceBuffer small;  // broadcast (non-slabbed) inputs, edges, image planes, stats
    bool init = false;
    int ensure_streams() {
        if (init) return OPTK_OK;
        for (int k = 0; k < kSlots; ++k) OPTK_CUDA(cudaStreamCreateWithFlags(&stream[k], cudaStreamNonBlocking));
        init = true;
        return OPTK_OK;
    }
};

HostPipeline g_pipeline;

int validate_surface(const optk_surface_t& s, int index) {
    if (s.sag_kind < OPTK_SAG_FLAT || s.sag_kind > OPTK_SAG_TOROIDAL) {
        set_error("surface %d: unsupported sag kind %d", index, s.sag_kind);
        return OPTK_ERR_UNSUPPORTED;
    }
    if (s.material_kind < OPTK_MAT_VACUUM || s.material_kind > OPTK_MAT_PASS) {
        set_error("surface %d: unsupported material kind %d", index, s.material_kind);
        return OPTK_ERR_UNSUPPORTED;
    }
    if (s.ruling_kind < OPTK_RULING_NONE || s.ruling_kind > OPTK_RULING_HOLOGRAPHIC) {
        set_error("surface %d: unsupported ruling kind %d", index, s.ruling_kind);
        return OPTK_ERR_UNSUPPORTED;
    }
    if (s.aperture_kind < OPTK_APERTURE_NONE || s.aperture_kind > OPTK_APERTURE_SECTOR) {
        set_error("surface %d: unsupported aperture kind %d", index, s.aperture_kind);
        return OPTK_ERR_UNSUPPORTED;
    }
    if (s.aperture_kind == OPTK_APERTURE_POLYGON && (s.n_vertices < 3 || s.n_vertices > OPTK_MAX_VERTICES)) {
        set_error("surface %d: polygon aperture needs 3..%d vertices, got %d", index, OPTK_MAX_VERTICES,
                  s.n_vertices);
        return OPTK_ERR_INVALID;
    }
    if (s.ruling_kind == OPTK_RULING_POLYNOMIAL && (s.n_coeff < 1 || s.n_coeff > OPTK_MAX_COEFF)) {
        set_error("surface %d: polynomial ruling needs 1..%d coefficients, got %d", index, OPTK_MAX_COEFF,
                  s.n_coeff);
        return OPTK_ERR_INVALID;
    }
    if (s.material_efficiency < OPTK_EFF_UNIT || s.material_efficiency > OPTK_EFF_TABLE2D) {
        set_error("surface %d: unsupported material efficiency kind %d", index, s.material_efficiency);
        return OPTK_ERR_UNSUPPORTED;
    }
    if (s.material_efficiency == OPTK_EFF_TABLE2D) {
        if (s.material_kind != OPTK_MAT_MIRROR && s.material_kind != OPTK_MAT_PASS) {
            set_error("surface %d: a 2-D efficiency table needs a mirror or pass-through material", index);
            return OPTK_ERR_INVALID;
        }
        if (s.material_lut_n < 2 || !s.material_lut_x || !s.material_lut_y || !(s.material[1] > 0.0) ||
            !(s.material[2] >= 4.0) || s.material[2] != (double)(int)s.material[2]) {
            set_error("surface %d: a 2-D efficiency table needs >= 2 wavelength nodes, >= 4 cosine nodes with a "
                      "positive spacing, and device pointers to both arrays", index);
            return OPTK_ERR_INVALID;
        }
        if (s.ruling_profile == OPTK_PROFILE_MEASURED) {
            set_error("surface %d: a 2-D efficiency table cannot be combined with measured rulings", index);
            return OPTK_ERR_UNSUPPORTED;
        }
    }
    if (s.ruling_profile < OPTK_PROFILE_IDEAL || s.ruling_profile > OPTK_PROFILE_MEASURED) {
        set_error("surface %d: unsupported ruling profile %d", index, s.ruling_profile);
        return OPTK_ERR_UNSUPPORTED;
    }
    if (s.ruling_profile != OPTK_PROFILE_IDEAL && s.ruling_kind == OPTK_RULING_NONE) {
        set_error("surface %d: a ruling profile needs a ruling spacing", index);
        return OPTK_ERR_INVALID;
    }
    for (int which = 0; which < 2; ++which) {
        const bool used = which == 0 ? s.material_efficiency == OPTK_EFF_LUT : s.ruling_profile == OPTK_PROFILE_MEASURED;
        if (!used) continue;
        const int n = which == 0 ? s.material_lut_n : s.ruling_lut_n;
        const double* x = which == 0 ? s.material_lut_x : s.ruling_lut_x;
        const double* y = which == 0 ? s.material_lut_y : s.ruling_lut_y;
        if (n < 1 || !x || !y) {
            set_error("surface %d: measured efficiency needs a table of at least one (wavelength, value) pair", index);
            return OPTK_ERR_INVALID;
        }
        for (int i = 1; i < n; ++i)
            if (!(x[i] > x[i - 1])) {
                set_error("surface %d: measured efficiency wavelengths must be strictly ascending", index);
                return OPTK_ERR_INVALID;
            }
    }
    return OPTK_OK;
}

// Number of rays and a check of the grid description.
int grid_size(const optk_rays_in_t* in, long long* n_out) {
    if (!in) {
        set_error("rays_in is NULL");
        return OPTK_ERR_INVALID;
    }
    if (in->n_axes < 0 || in->n_axes > OPTK_MAX_AXES) {
        set_error("n_axes must be in 0..%d, got %d", OPTK_MAX_AXES, in->n_axes);
        return OPTK_ERR_INVALID;
    }
    long long n = 1;
    for (int a = 0; a < in->n_axes; ++a) {
        if (in->dims[a] < 0) {
            set_error("dims[%d] is negative", a);
            return OPTK_ERR_INVALID;
        }
        n *= in->dims[a];
    }
    for (int f = 0; f < OPTK_NUM_FIELDS; ++f) {
        if (!in->field[f]) {
            set_error("input field %d is NULL", f);
            return OPTK_ERR_INVALID;
        }
    }
    *n_out = n;
    return OPTK_OK;
}

bool stride_is_dense(const int64_t* stride, const int64_t* dims, int n_axes) {
    long long expect = 1;
    for (int a = n_axes - 1; a >= 0; --a) {
        if (dims[a] != 1 && stride[a] != expect) return false;
        expect *= dims[a];
    }
    return true;
}

bool all_dense(const optk_rays_in_t& in) {
    for (int f = 0; f < OPTK_NUM_FIELDS; ++f)
        if (!stride_is_dense(in.stride[f], in.dims, in.n_axes)) return false;
    if (in.normal[0])
        for (int k = 0; k < 3; ++k)
            if (!in.normal[k] || !stride_is_dense(in.normal_stride[k], in.dims, in.n_axes)) return false;
    if (in.unvignetted && !stride_is_dense(in.mask_stride, in.dims, in.n_axes)) return false;
    return true;
}

// 1 + the largest element offset a strided view touches
long long view_extent(const int64_t* stride, const int64_t* dims, int n_axes) {
    long long e = 0;
    for (int a = 0; a < n_axes; ++a) {
        if (dims[a] == 0) return 0;
        e += (dims[a] - 1) * (stride[a] < 0 ? 0 : stride[a]);
    }
    return e + 1;
}

int fill_image(const optk_image_t* image, ImageDev* dev, bool groups_allowed = true) {
    if (image->group_size != 0) {
        // per-group accumulators instead of a detector (optk_image_t::group_size)
        if (!groups_allowed || image->group_size < 0 || image->group_size > 0x7fffffffLL || image->n_x < 1 ||
            image->n_wavelength != 1 || image->n_y != 1) {
            set_error("image.group_size needs optk_trace / optk_trace_grid, 0 < group_size < 2^31, n_x >= 1 groups and "
                      "n_wavelength = n_y = 1");
            return OPTK_ERR_INVALID;
        }
        memset(dev, 0, sizeof(*dev));
        dev->n_w = 1;
        dev->n_x = image->n_x;
        dev->n_y = 1;
        dev->group = 1;
        dev->flux = image->flux;
        dev->moment_real = image->moment_real;
        dev->moment_imag = image->moment_imag;
        dev->counts = image->counts;
        dev->has_range = 1;  // nothing to fetch: the edges are not used
        dev->range[1] = dev->range[3] = dev->range[5] = 1.0;
        dev->div_group = make_fastdiv((uint32_t)image->group_size);
        return OPTK_OK;
    }
    if (image->n_wavelength < 1 || image->n_x < 1 || image->n_y < 1) {
        set_error("image needs at least one bin along every axis");
        return OPTK_ERR_INVALID;
    }
    if (!image->edges_wavelength || !image->edges_x || !image->edges_y) {
        set_error("image bin edges are NULL");
        return OPTK_ERR_INVALID;
    }
    if ((long long)image->n_wavelength * image->n_x * image->n_y > 0x7fffffffLL) {
        set_error("an image plane set holds at most 2^31 - 1 bins; split the wavelength axis");
        return OPTK_ERR_INVALID;
    }
    dev->n_w = image->n_wavelength;
    dev->n_x = image->n_x;
    dev->n_y = image->n_y;
    dev->group = 0;
    dev->div_group = make_fastdiv(1u);
    dev->e_w = image->edges_wavelength;
    dev->e_x = image->edges_x;
    dev->e_y = image->edges_y;
    dev->flux = image->flux;
    dev->moment_real = image->moment_real;
    dev->moment_imag = image->moment_imag;
    dev->counts = image->counts;
    dev->has_range = image->has_range ? 1 : 0;
    dev->uniform = image->has_range ? (image->uniform_edges & 3) : 0;
    for (int k = 0; k < 6; ++k) dev->range[k] = image->range[k];
    // the same expression the kernels evaluate when they have to fetch the range themselves (image_guess_fill)
    dev->inv_dx = image->has_range ? (double)image->n_x / (image->range[3] - image->range[2]) : 0.0;
    dev->inv_dy = image->has_range ? (double)image->n_y / (image->range[5] - image->range[4]) : 0.0;
    return OPTK_OK;
}

// Common packing of everything except ray pointers.
int pack_trace(const optk_system_t* sys, int32_t config, int32_t surf_begin, int32_t surf_count, int32_t surf_step,
               int32_t accumulate, TraceParams* P) {
    if (!sys) {
        set_error("system handle is NULL");
        return OPTK_ERR_INVALID;
    }
    if (config < 0 || config >= sys->n_config) {
        set_error("config %d out of range [0, %d)", config, sys->n_config);
        return OPTK_ERR_INVALID;
    }
    if (surf_count < 0 || surf_count > OPTK_MAX_SURFACES) {
        set_error("surf_count %d out of range [0, %d]; chain longer systems", surf_count, OPTK_MAX_SURFACES);
        return OPTK_ERR_INVALID;
    }
    if (surf_step != 1 && surf_step != -1) {
        set_error("surf_step must be +1 or -1");
        return OPTK_ERR_INVALID;
    }
    if (sys->lut_device) {
        int current = -1;
        cudaGetDevice(&current);
        if (current != sys->lut_ordinal) {
            set_error("this system handle owns efficiency tables on device %d but the launch is on device %d; "
                      "create one handle per device", sys->lut_ordinal, current);
            return OPTK_ERR_INVALID;
        }
    }
    const optk_surface_t* row = sys->table.data() + (size_t)config * sys->n_surface;
    for (int k = 0; k < surf_count; ++k) {
        const int s = surf_begin + k * surf_step;
        if (s < 0 || s >= sys->n_surface) {
            set_error("surface index %d out of range [0, %d)", s, sys->n_surface);
            return OPTK_ERR_INVALID;
        }
        P->surf[k] = row[s];
    }
    P->n_surf = surf_count;
    P->accumulate = accumulate ? 1 : 0;
    P->relative_done = 0;  // fresh copies of the surfaces: launch_trace may compose their frames again
    return OPTK_OK;
}

// A polygon whose vertices are in strictly convex position (every vertex clearly on the inner side of every
// edge it does not belong to, which also rules out star polygons, repeated and collinear vertices) is the
// intersection of its edges' half-planes: see OPTK_F_APERTURE_CONVEX in optk.h and aperture_test.
void classify_polygon(optk_surface_t& s) {
    const int n = s.n_vertices;
    if (n < 3 || n > OPTK_MAX_VERTICES) return;
    double bound = 0.0, area2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        if (!std::isfinite(s.vertices_x[i]) || !std::isfinite(s.vertices_y[i])) return;
        bound = std::fmax(bound, std::fmax(std::fabs(s.vertices_x[i]), std::fabs(s.vertices_y[i])));
        area2 += s.vertices_x[i] * s.vertices_y[j] - s.vertices_x[j] * s.vertices_y[i];
    }
    if (!(bound > 0.0) || !(bound < 1e150) || area2 == 0.0) return;
    const double orientation = area2 > 0.0 ? 1.0 : -1.0;
    const double margin = 1e-9 * bound * bound;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        const double ex = s.vertices_x[j] - s.vertices_x[i], ey = s.vertices_y[j] - s.vertices_y[i];
        for (int k = 0; k < n; ++k) {
            if (k == i || k == j) continue;
            const double c = ex * (s.vertices_y[k] - s.vertices_y[i]) - ey * (s.vertices_x[k] - s.vertices_x[i]);
            if (!(orientation * c > margin)) return;
        }
    }
    s.flags |= OPTK_F_APERTURE_CONVEX | (orientation < 0.0 ? OPTK_F_APERTURE_CLOCKWISE : 0);
    s.aperture[0] = bound;
    s.aperture[1] = 1e-12 * bound * bound;
}

}  // namespace

extern "C" {

OPTK_API int optk_abi_version(void) { return OPTK_ABI_VERSION; }

OPTK_API const char* optk_last_error(void) { return g_error; }

OPTK_API int optk_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

OPTK_API int optk_system_create(const optk_surface_t* table, int32_t n_surface, int32_t n_config, optk_system_t** out) {
    if (!table || !out || n_surface < 1 || n_config < 1) {
        set_error("optk_system_create: bad arguments");
        return OPTK_ERR_INVALID;
    }
    for (int c = 0; c < n_config; ++c)
        for (int s = 0; s < n_surface; ++s) {
            int rc = validate_surface(table[(size_t)c * n_surface + s], s);
            if (rc) return rc;
        }
    optk_system* sys = new (std::nothrow) optk_system;
    if (!sys) return OPTK_ERR_NOMEM;
    sys->n_surface = n_surface;
    sys->n_config = n_config;
    sys->table.assign(table, table + (size_t)n_surface * n_config);
    for (optk_surface_t& s : sys->table) {
        // per-surface constants every ray would otherwise recompute
        s.sag[3] = 1.0 / s.sag[0];
        s.flags &= ~OPTK_F_TRANSLATION_ONLY;
        const double* r = s.transform.r;
        const bool identity = r[0] == 1.0 && r[4] == 1.0 && r[8] == 1.0 && r[1] == 0.0 && r[2] == 0.0 &&
                              r[3] == 0.0 && r[5] == 0.0 && r[6] == 0.0 && r[7] == 0.0;
        if ((s.flags & OPTK_F_TRANSFORM) && identity) s.flags |= OPTK_F_TRANSLATION_ONLY;
        s.flags &= ~(OPTK_F_APERTURE_CONVEX | OPTK_F_APERTURE_CLOCKWISE);
        if (s.aperture_kind == OPTK_APERTURE_POLYGON) classify_polygon(s);
        if (s.aperture_kind == OPTK_APERTURE_CIRCULAR || s.aperture_kind == OPTK_APERTURE_SECTOR) {
            // T = max{v : sqrt(v) <= radius} (correctly rounded sqrt): "sqrt(x^2 + y^2) <= radius"
            // (optika/apertures/_apertures.py:309) becomes "x^2 + y^2 <= T" with identical results
            const double radius = s.aperture[0];
            double t = radius * radius;
            if (radius >= 0.0 && t < INFINITY) {
                while (std::sqrt(t) > radius) t = std::nextafter(t, 0.0);
                while (std::sqrt(std::nextafter(t, INFINITY)) <= radius) t = std::nextafter(t, INFINITY);
            } else if (!(radius >= 0.0)) {
                t = -1.0;  // negative or NaN radius: nothing is inside
            }
            s.aperture[3] = t;
        }
    }
    // measured-efficiency tables: host arrays -> one device allocation owned by the handle
    size_t lut_total = 0;
    for (const optk_surface_t& s : sys->table) {
        if (s.material_efficiency == OPTK_EFF_LUT) lut_total += 2 * (size_t)s.material_lut_n;
        if (s.ruling_profile == OPTK_PROFILE_MEASURED) lut_total += 2 * (size_t)s.ruling_lut_n;
    }
    if (lut_total) {
        std::vector<double> host;
        host.reserve(lut_total);
        cudaGetDevice(&sys->lut_ordinal);
        cudaError_t e = cudaMalloc((void**)&sys->lut_device, lut_total * sizeof(double));
        if (e != cudaSuccess) {
            delete sys;
            return cuda_fail(e, "cudaMalloc (efficiency tables)");
        }
        for (optk_surface_t& s : sys->table) {
            for (int which = 0; which < 2; ++which) {
                const bool used = which == 0 ? s.material_efficiency == OPTK_EFF_LUT : s.ruling_profile == OPTK_PROFILE_MEASURED;
                if (!used) continue;
                const int n = which == 0 ? s.material_lut_n : s.ruling_lut_n;
                const double*& x = which == 0 ? s.material_lut_x : s.ruling_lut_x;
                const double*& y = which == 0 ? s.material_lut_y : s.ruling_lut_y;
                const size_t at = host.size();
                host.insert(host.end(), x, x + n);
                host.insert(host.end(), y, y + n);
                x = sys->lut_device + at;
                y = sys->lut_device + at + n;
            }
        }
        e = cudaMemcpy(sys->lut_device, host.data(), lut_total * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(sys->lut_device);
            delete sys;
            return cuda_fail(e, "cudaMemcpy (efficiency tables)");
        }
    }
    *out = sys;
    return OPTK_OK;
}

OPTK_API int optk_system_destroy(optk_system_t* sys) {
    if (sys && sys->lut_device) cudaFree(sys->lut_device);
    delete sys;
    return OPTK_OK;
}

OPTK_API int optk_system_size(const optk_system_t* sys, int32_t* n_surface, int32_t* n_config) {
    if (!sys) {
        set_error("system handle is NULL");
        return OPTK_ERR_INVALID;
    }
    if (n_surface) *n_surface = sys->n_surface;
    if (n_config) *n_config = sys->n_config;
    return OPTK_OK;
}

OPTK_API int optk_system_surface(const optk_system_t* sys, int32_t config, int32_t index, optk_surface_t* out) {
    if (!sys || !out) {
        set_error("optk_system_surface: NULL argument");
        return OPTK_ERR_INVALID;
    }
    if (config < 0 || config >= sys->n_config || index < 0 || index >= sys->n_surface) {
        set_error("optk_system_surface: configuration %d / surface %d out of range [0, %d) x [0, %d)", config, index,
                  sys->n_config, sys->n_surface);
        return OPTK_ERR_INVALID;
    }
    *out = sys->table[(size_t)config * sys->n_surface + index];
    return OPTK_OK;
}

OPTK_API int optk_trace(const optk_system_t* sys, int32_t config, const optk_rays_in_t* in, const optk_rays_out_t* out,
               int32_t surf_begin, int32_t surf_count, int32_t surf_step, int32_t accumulate,
               int64_t accumulate_stride, const optk_image_t* image, const optk_affine_t* image_frame,
               optk_trace_stats_t* stats_device, void* stream) {
    DeviceScope device_scope(stream, "optk_trace");
    static thread_local TraceParams P;
    long long n = 0;
    int rc = grid_size(in, &n);
    if (rc) return rc;
    rc = pack_trace(sys, config, surf_begin, surf_count, surf_step, accumulate, &P);
    if (rc) return rc;
    if (n > 0x7fffffffLL) {
        set_error("optk_trace: %lld rays exceed one launch (2^31 - 1); split the outermost axis", n);
        return OPTK_ERR_INVALID;
    }
    P.in = *in;
    if (out) P.out = *out; else memset(&P.out, 0, sizeof(P.out));
    for (int a = 0; a < OPTK_MAX_AXES; ++a)
        P.div[a] = make_fastdiv(a < in->n_axes ? (uint32_t)in->dims[a] : 1u);
    P.n_rays = n;
    P.index_offset = 0;
    P.flat_index = 0;
    P.accumulate_stride = accumulate ? accumulate_stride : 0;
    if (accumulate && accumulate_stride < n) {
        set_error("accumulate_stride %lld is smaller than the number of rays %lld", (long long)accumulate_stride, n);
        return OPTK_ERR_INVALID;
    }
    P.dense_in = all_dense(*in) ? 1 : 0;
    P.has_image = image ? 1 : 0;
    P.has_frame = image_frame ? 1 : 0;
    if (image) {
        rc = fill_image(image, &P.image);
        if (rc) return rc;
    }
    if (image_frame) P.frame = *image_frame;
    P.stats = stats_device;
    return launch_trace(P, (cudaStream_t)stream);
}

OPTK_API int optk_trace_grid(const optk_system_t* sys, int32_t config, const optk_grid_t* grid, const optk_rays_out_t* out,
                    int32_t surf_begin, int32_t surf_count, int32_t surf_step, int32_t accumulate,
                    int64_t accumulate_stride, const optk_image_t* image, const optk_affine_t* image_frame,
                    optk_trace_stats_t* stats_device, void* stream) {
    DeviceScope device_scope(stream, "optk_trace_grid");
    static thread_local TraceParams P;
    if (!grid) {
        set_error("optk_trace_grid: grid is NULL");
        return OPTK_ERR_INVALID;
    }
    long long n = 1;
    for (int a = 0; a < 5; ++a) {
        if (grid->n[a] < 1 || grid->begin[a] < 0 || grid->count[a] < 0 ||
            (long long)grid->begin[a] + grid->count[a] > grid->n[a]) {
            set_error("optk_trace_grid: axis %d: sub-box [%d, %d + %d) does not fit %d cells", a, grid->begin[a],
                      grid->begin[a], grid->count[a], grid->n[a]);
            return OPTK_ERR_INVALID;
        }
        if (!grid->vertices[a]) {
            set_error("optk_trace_grid: vertices[%d] is NULL", a);
            return OPTK_ERR_INVALID;
        }
        n *= grid->count[a];
        if (n > 0x7fffffffLL) {
            set_error("optk_trace_grid: sub-box exceeds one launch (2^31 - 1 rays); split it");
            return OPTK_ERR_INVALID;
        }
    }
    int rc = pack_trace(sys, config, surf_begin, surf_count, surf_step, accumulate, &P);
    if (rc) return rc;
    if (accumulate && accumulate_stride < n) {
        set_error("accumulate_stride %lld is smaller than the number of rays %lld", (long long)accumulate_stride, n);
        return OPTK_ERR_INVALID;
    }
    memset(&P.in, 0, sizeof(P.in));
    P.in.n_axes = 5;
    for (int a = 0; a < OPTK_MAX_AXES; ++a) {
        if (a < 5) P.in.dims[a] = grid->count[a];
        P.div[a] = make_fastdiv(a < 5 ? (uint32_t)(grid->count[a] > 0 ? grid->count[a] : 1) : 1u);
    }
    if (out) P.out = *out; else memset(&P.out, 0, sizeof(P.out));
    P.n_rays = n;
    P.index_offset = 0;
    P.flat_index = 0;
    P.accumulate_stride = accumulate ? accumulate_stride : 0;
    P.dense_in = 0;
    P.from_grid = 1;
    P.grid = *grid;
    P.cell_stride[4] = 1;
    for (int a = 3; a >= 0; --a) P.cell_stride[a] = P.cell_stride[a + 1] * (unsigned long long)grid->n[a + 1];
    if ((long long)grid->n[0] * grid->n[1] > 0x7fffffffLL) {
        set_error("optk_trace_grid: n[0] * n[1] exceeds 2^31 - 1");
        return OPTK_ERR_INVALID;
    }
    P.has_image = image ? 1 : 0;
    P.has_frame = image_frame ? 1 : 0;
    if (image) {
        rc = fill_image(image, &P.image);
        if (rc) return rc;
    }
    if (image_frame) P.frame = *image_frame;
    P.stats = stats_device;
    rc = launch_trace(P, (cudaStream_t)stream);
    P.from_grid = 0;
    return rc;
}

OPTK_API int optk_solve_stops(const optk_system_t* sys, int32_t config, const optk_stop_problem_t* problem, int64_t n,
                     const double* wavelength, const double* fixed_x, const double* fixed_y, const double* fixed_z,
                     const double* target_x, const double* target_y, double* x, double* y, double* z,
                     uint32_t* n_unconverged, void* stream) {
    DeviceScope device_scope(stream, "optk_solve_stops");
    static thread_local TraceParams P;
    if (!problem || n < 0) {
        set_error("optk_solve_stops: problem is NULL or n is negative");
        return OPTK_ERR_INVALID;
    }
    if (n > 0 && (!wavelength || !fixed_x || !fixed_y || !fixed_z || !target_x || !target_y || !x || !y || !z ||
                  !n_unconverged)) {
        set_error("optk_solve_stops: NULL array");
        return OPTK_ERR_INVALID;
    }
    if ((problem->variable != OPTK_STOP_DIRECTION && problem->variable != OPTK_STOP_POSITION) ||
        (problem->target != OPTK_STOP_DIRECTION && problem->target != OPTK_STOP_POSITION)) {
        set_error("optk_solve_stops: variable / target must be OPTK_STOP_DIRECTION or OPTK_STOP_POSITION");
        return OPTK_ERR_INVALID;
    }
    const int count = problem->surf_last - problem->surf_first;
    if (count < 1 || count > OPTK_MAX_SURFACES - 1) {
        set_error("optk_solve_stops: surf_last - surf_first = %d out of range [1, %d]", count, OPTK_MAX_SURFACES - 1);
        return OPTK_ERR_INVALID;
    }
    if (!(problem->step > 0.0) || !(problem->max_abs_error >= 0.0) || problem->max_iterations < 1) {
        set_error("optk_solve_stops: step must be positive, max_abs_error non-negative, max_iterations >= 1");
        return OPTK_ERR_INVALID;
    }
    // surfaces surf_first + 1 ... surf_last are walked; the first stop surface rides along in the
    // next slot for its sag
    int rc = pack_trace(sys, config, problem->surf_first + 1, count, 1, 0, &P);
    if (rc) return rc;
    if (problem->surf_first < 0) {
        set_error("optk_solve_stops: surf_first is negative");
        return OPTK_ERR_INVALID;
    }
    const optk_surface_t& first = sys->table[(size_t)config * sys->n_surface + problem->surf_first];
    int sag_slot = -1;
    if (problem->variable == OPTK_STOP_POSITION && first.sag_kind != OPTK_SAG_FLAT) {
        if (first.flags & OPTK_F_SAG_TRANSFORM) {
            set_error("optk_solve_stops: a first stop surface whose sag carries a transformation is not supported");
            return OPTK_ERR_UNSUPPORTED;
        }
        sag_slot = count;
        P.surf[sag_slot] = first;
    }
    for (int k = 0; k < count; ++k) {
        if (P.surf[k].stages != OPTK_STAGE_ALL) {
            set_error("optk_solve_stops: the system must be compiled with the full surface operator");
            return OPTK_ERR_INVALID;
        }
    }
    P.surf[count - 1].flags |= OPTK_F_LOCAL_OUT;  // residuals live in the last surface's local frame
    memset(&P.in, 0, sizeof(P.in));
    memset(&P.out, 0, sizeof(P.out));
    P.n_rays = n;
    P.stats = nullptr;
    P.has_image = 0;
    P.has_frame = 0;
    const double* fixed[3] = {fixed_x, fixed_y, fixed_z};
    const double* target[2] = {target_x, target_y};
    return launch_stop_newton(P, *problem, sag_slot, n, wavelength, fixed, target, x, y, z, n_unconverged,
                              (cudaStream_t)stream);
}

OPTK_API int optk_reduce_groups(int64_t n_groups, int64_t n_inner, const double* x, const double* y,
                       const double* intensity, const uint8_t* unvignetted, double* sum_intensity, double* sum_x,
                       double* sum_y, uint64_t* count, double* sum_x_all, double* sum_y_all, void* stream) {
    DeviceScope device_scope(stream, "optk_reduce_groups");
    if (n_groups < 0 || n_inner < 1 || n_groups > 0x7fffffffLL || n_inner > 0x7fffffffLL ||
        n_groups * n_inner > 0x7fffffffLL) {
        set_error("optk_reduce_groups: n_groups >= 0, n_inner >= 1 and n_groups * n_inner <= 2^31 - 1 are required");
        return OPTK_ERR_INVALID;
    }
    if (n_groups > 0 && (!x || !y)) {
        set_error("optk_reduce_groups: x or y is NULL");
        return OPTK_ERR_INVALID;
    }
    return launch_reduce_groups(n_groups, n_inner, x, y, intensity, unvignetted, sum_intensity, sum_x, sum_y,
                                (unsigned long long*)count, sum_x_all, sum_y_all, (cudaStream_t)stream);
}

OPTK_API int optk_jit_mode(int32_t mode) {
    jit_set_mode(mode);
    return OPTK_OK;
}

OPTK_API int64_t optk_jit_compiled(void) { return jit_compiled_count(); }

OPTK_API int optk_interp(int64_t n, const double* x, int32_t m, const double* xp, const double* fp_re, const double* fp_im,
                double* out_re, double* out_im, void* stream) {
    DeviceScope device_scope(stream, "optk_interp");
    if (n < 0 || m < 1 || !x || !xp || !fp_re || !out_re || (fp_im && !out_im)) {
        set_error("optk_interp: bad arguments");
        return OPTK_ERR_INVALID;
    }
    return launch_interp(n, x, m, xp, fp_re, fp_im, out_re, out_im, (cudaStream_t)stream);
}

OPTK_API int optk_apply_efficiency(int64_t n, double* intensity, const double* e_s, const double* e_p, void* stream) {
    DeviceScope device_scope(stream, "optk_apply_efficiency");
    if (n < 0 || !intensity || !e_s || !e_p) {
        set_error("optk_apply_efficiency: bad arguments");
        return OPTK_ERR_INVALID;
    }
    return launch_apply_efficiency(n, intensity, e_s, e_p, (cudaStream_t)stream);
}

OPTK_API int optk_debug_math(int32_t op, int64_t n, const double* a, const double* b, double* out, void* stream) {
    DeviceScope device_scope(stream, "optk_debug_math");
    if (op < 0 || op > 7 || n < 0 || !a || !out || ((op == 0 || op >= 6) && !b)) {
        set_error("optk_debug_math: bad arguments");
        return OPTK_ERR_INVALID;
    }
    return launch_debug_math(op, n, a, b, out, (cudaStream_t)stream);
}

OPTK_API int optk_bin(int64_t n_rays, const double* wavelength, const double* x, const double* y, const double* dz,
             const double* intensity, const uint8_t* unvignetted, const optk_image_t* image, void* stream) {
    DeviceScope device_scope(stream, "optk_bin");
    if (!wavelength || !x || !y || !image) {
        set_error("optk_bin: NULL argument");
        return OPTK_ERR_INVALID;
    }
    ImageDev dev;
    int rc = fill_image(image, &dev, false);
    if (rc) return rc;
    return launch_bin(n_rays, wavelength, x, y, dz, intensity, unvignetted, dev, (cudaStream_t)stream);
}

OPTK_API int optk_trace_host(const optk_system_t* sys, int32_t config, const optk_rays_in_t* in, const optk_rays_out_t* out,
                    int32_t surf_begin, int32_t surf_count, int32_t surf_step, int32_t accumulate,
                    int64_t accumulate_stride, const optk_image_t* image, const optk_affine_t* image_frame,
                    optk_trace_stats_t* stats_host, int64_t slab_rays, int32_t pinned) {
    (void)pinned;
    static thread_local TraceParams P;
    long long n = 0;
    int rc = grid_size(in, &n);
    if (rc) return rc;
    rc = pack_trace(sys, config, surf_begin, surf_count, surf_step, accumulate, &P);
    if (rc) return rc;
    if (n > 0xffffffffLL) {
        set_error("optk_trace_host: %lld rays exceed 2^32 - 1; split the outermost axis", n);
        return OPTK_ERR_INVALID;
    }
    if (accumulate && accumulate_stride < n) {
        set_error("accumulate_stride %lld is smaller than the number of rays %lld", (long long)accumulate_stride, n);
        return OPTK_ERR_INVALID;
    }
    if (n == 0) return OPTK_OK;
    if (in->normal[0]) {
        set_error("optk_trace_host: caller-supplied normals are only supported with device pointers");
        return OPTK_ERR_UNSUPPORTED;
    }

    std::lock_guard<std::mutex> lock(g_pipeline.mutex);
    HostPipeline& pl = g_pipeline;
    rc = pl.ensure_streams();
    if (rc) return rc;

    if (slab_rays <= 0) slab_rays = 1 << 22;
    if (slab_rays > n) slab_rays = n;
    const int n_slabs = (int)((n + slab_rays - 1) / slab_rays);
    const int n_out_states = accumulate ? surf_count : 1;

    // --- classify inputs: dense fields are streamed in slabs, the rest are
    //     broadcast views whose (small) backing arrays are uploaded once.
    bool dense[OPTK_NUM_FIELDS + 1];
    long long extent[OPTK_NUM_FIELDS + 1];
    size_t small_bytes = 0;
    for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) {
        const int64_t* st = f < OPTK_NUM_FIELDS ? in->stride[f] : in->mask_stride;
        const bool present = f < OPTK_NUM_FIELDS ? true : in->unvignetted != nullptr;
        dense[f] = present && stride_is_dense(st, in->dims, in->n_axes);
        extent[f] = present ? view_extent(st, in->dims, in->n_axes) : 0;
        if (present && !dense[f]) {
            for (int a = 0; a < in->n_axes; ++a)
                if (st[a] < 0) {
                    set_error("negative strides are not supported");
                    return OPTK_ERR_INVALID;
                }
            small_bytes += ((size_t)extent[f] * (f < OPTK_NUM_FIELDS ? 8 : 1) + 255) & ~(size_t)255;
        }
    }
    size_t image_bytes = 0, plane_elems = 0;
    if (image) {
        plane_elems = (size_t)image->n_wavelength * image->n_x * image->n_y;
        image_bytes += (((size_t)image->n_wavelength + 1) * 8 + 255) & ~(size_t)255;
        image_bytes += (((size_t)image->n_x + 1) * 8 + 255) & ~(size_t)255;
        image_bytes += (((size_t)image->n_y + 1) * 8 + 255) & ~(size_t)255;
        image_bytes += 4 * ((plane_elems * 8 + 255) & ~(size_t)255);
    }
    const size_t stats_bytes = 256;
    rc = pl.small.ensure(small_bytes + image_bytes + stats_bytes);
    if (rc) return rc;

    cudaStream_t s0 = pl.stream[0];
    char* cursor = (char*)pl.small.ptr;
    auto take = [&](size_t bytes) {
        char* p = cursor;
        cursor += (bytes + 255) & ~(size_t)255;
        return (void*)p;
    };

    P.in = *in;
    for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) {
        const bool present = f < OPTK_NUM_FIELDS ? true : in->unvignetted != nullptr;
        if (!present || dense[f]) continue;
        const size_t bytes = (size_t)extent[f] * (f < OPTK_NUM_FIELDS ? 8 : 1);
        void* d = take(bytes);
        const void* h = f < OPTK_NUM_FIELDS ? (const void*)in->field[f] : (const void*)in->unvignetted;
        OPTK_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s0));
        if (f < OPTK_NUM_FIELDS) P.in.field[f] = (const double*)d; else P.in.unvignetted = (const uint8_t*)d;
    }
    optk_image_t image_dev;
    if (image) {
        image_dev = *image;
        double* ew = (double*)take(((size_t)image->n_wavelength + 1) * 8);
        double* ex = (double*)take(((size_t)image->n_x + 1) * 8);
        double* ey = (double*)take(((size_t)image->n_y + 1) * 8);
        OPTK_CUDA(cudaMemcpyAsync(ew, image->edges_wavelength, ((size_t)image->n_wavelength + 1) * 8,
                                  cudaMemcpyHostToDevice, s0));
        OPTK_CUDA(cudaMemcpyAsync(ex, image->edges_x, ((size_t)image->n_x + 1) * 8, cudaMemcpyHostToDevice, s0));
        OPTK_CUDA(cudaMemcpyAsync(ey, image->edges_y, ((size_t)image->n_y + 1) * 8, cudaMemcpyHostToDevice, s0));
        image_dev.edges_wavelength = ew;
        image_dev.edges_x = ex;
        image_dev.edges_y = ey;
        image_dev.has_range = 1;
        image_dev.range[0] = image->edges_wavelength[0];
        image_dev.range[1] = image->edges_wavelength[image->n_wavelength];
        image_dev.range[2] = image->edges_x[0];
        image_dev.range[3] = image->edges_x[image->n_x];
        image_dev.range[4] = image->edges_y[0];
        image_dev.range[5] = image->edges_y[image->n_y];
        void* host_planes[4] = {image->flux, image->moment_real, image->moment_imag, image->counts};
        void* dev_planes[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int k = 0; k < 4; ++k) {
            if (!host_planes[k]) continue;
            dev_planes[k] = take(plane_elems * 8);
            OPTK_CUDA(cudaMemcpyAsync(dev_planes[k], host_planes[k], plane_elems * 8, cudaMemcpyHostToDevice, s0));
        }
        image_dev.flux = (double*)dev_planes[0];
        image_dev.moment_real = (double*)dev_planes[1];
        image_dev.moment_imag = (double*)dev_planes[2];
        image_dev.counts = (unsigned long long*)dev_planes[3];
        rc = fill_image(&image_dev, &P.image, false);
        if (rc) return rc;
    }
    optk_trace_stats_t* stats_dev = nullptr;
    if (stats_host) {
        stats_dev = (optk_trace_stats_t*)take(sizeof(optk_trace_stats_t));
        OPTK_CUDA(cudaMemsetAsync(stats_dev, 0, sizeof(optk_trace_stats_t), s0));
    }
    OPTK_CUDA(cudaStreamSynchronize(s0));

    // --- per-slot slab buffers
    size_t in_bytes = 0, out_bytes = 0;
    for (int f = 0; f < OPTK_NUM_FIELDS; ++f)
        if (dense[f]) in_bytes += (size_t)slab_rays * 8;
    if (dense[OPTK_NUM_FIELDS]) in_bytes += ((size_t)slab_rays + 255) & ~(size_t)255;
    if (out) {
        for (int f = 0; f < OPTK_NUM_FIELDS; ++f)
            if (out->field[f]) out_bytes += (size_t)slab_rays * 8 * n_out_states;
        if (out->unvignetted) out_bytes += (((size_t)slab_rays + 255) & ~(size_t)255) * n_out_states;
    }
    const int slots = n_slabs < kSlots ? n_slabs : kSlots;
    for (int k = 0; k < slots; ++k) {
        rc = pl.in[k].ensure(in_bytes ? in_bytes : 256);
        if (rc) return rc;
        rc = pl.out[k].ensure(out_bytes ? out_bytes : 256);
        if (rc) return rc;
    }

    for (int a = 0; a < OPTK_MAX_AXES; ++a)
        P.div[a] = make_fastdiv(a < in->n_axes ? (uint32_t)in->dims[a] : 1u);
    P.has_image = image ? 1 : 0;
    P.has_frame = image_frame ? 1 : 0;
    if (image_frame) P.frame = *image_frame;
    P.stats = stats_dev;
    bool every_dense = true;
    for (int f = 0; f < OPTK_NUM_FIELDS; ++f) every_dense = every_dense && dense[f];
    if (in->unvignetted) every_dense = every_dense && dense[OPTK_NUM_FIELDS];
    P.dense_in = every_dense ? 1 : 0;

    const optk_rays_in_t base_in = P.in;
    for (int slab = 0; slab < n_slabs; ++slab) {
        const int k = slab % kSlots;
        cudaStream_t st = pl.stream[k];
        const long long i0 = (long long)slab * slab_rays;
        const long long m = (i0 + slab_rays <= n) ? slab_rays : n - i0;
        // the stream is in order: reusing slot k waits for its previous D2H

        // H2D of the dense inputs of this slab
        char* cin = (char*)pl.in[k].ptr;
        P.in = base_in;
        for (int f = 0; f < OPTK_NUM_FIELDS; ++f) {
            if (!dense[f]) continue;
            OPTK_CUDA(cudaMemcpyAsync(cin, in->field[f] + i0, (size_t)m * 8, cudaMemcpyHostToDevice, st));
            // dense view: element offset == global ray index; bias so that offset i0 + j lands on cin[j]
            P.in.field[f] = every_dense ? (const double*)cin : (const double*)cin - i0;
            cin += (size_t)slab_rays * 8;
        }
        if (dense[OPTK_NUM_FIELDS]) {
            OPTK_CUDA(cudaMemcpyAsync(cin, in->unvignetted + i0, (size_t)m, cudaMemcpyHostToDevice, st));
            P.in.unvignetted = every_dense ? (const uint8_t*)cin : (const uint8_t*)cin - i0;
        }
        // outputs of this slab
        char* cout = (char*)pl.out[k].ptr;
        memset(&P.out, 0, sizeof(P.out));
        if (out) {
            for (int f = 0; f < OPTK_NUM_FIELDS; ++f) {
                if (!out->field[f]) continue;
                P.out.field[f] = (double*)cout;
                cout += (size_t)slab_rays * 8 * n_out_states;
            }
            if (out->unvignetted) P.out.unvignetted = (uint8_t*)cout;
        }
        P.n_rays = m;
        P.flat_index = 1;
        P.index_offset = every_dense ? 0 : i0;
        P.accumulate_stride = accumulate ? slab_rays : 0;
        rc = launch_trace(P, st);
        if (rc) return rc;
        // D2H
        if (out) {
            for (int f = 0; f < OPTK_NUM_FIELDS; ++f) {
                if (!out->field[f]) continue;
                for (int s = 0; s < n_out_states; ++s)
                    OPTK_CUDA(cudaMemcpyAsync(out->field[f] + (long long)s * accumulate_stride + i0,
                                              P.out.field[f] + (long long)s * slab_rays, (size_t)m * 8,
                                              cudaMemcpyDeviceToHost, st));
            }
            if (out->unvignetted)
                for (int s = 0; s < n_out_states; ++s)
                    OPTK_CUDA(cudaMemcpyAsync(out->unvignetted + (long long)s * accumulate_stride + i0,
                                              P.out.unvignetted + (long long)s * slab_rays, (size_t)m,
                                              cudaMemcpyDeviceToHost, st));
        }
    }
    for (int k = 0; k < slots; ++k) OPTK_CUDA(cudaStreamSynchronize(pl.stream[k]));

    if (image) {
        void* host_planes[4] = {image->flux, image->moment_real, image->moment_imag, image->counts};
        void* dev_planes[4] = {image_dev.flux, image_dev.moment_real, image_dev.moment_imag, image_dev.counts};
        for (int k = 0; k < 4; ++k)
            if (host_planes[k])
                OPTK_CUDA(cudaMemcpyAsync(host_planes[k], dev_planes[k], plane_elems * 8, cudaMemcpyDeviceToHost, s0));
    }
    if (stats_host)
        OPTK_CUDA(cudaMemcpyAsync(stats_host, stats_dev, sizeof(optk_trace_stats_t), cudaMemcpyDeviceToHost, s0));
    OPTK_CUDA(cudaStreamSynchronize(s0));
    return OPTK_OK;
}

OPTK_API int optk_multilayer(const optk_ml_input_t* input, int32_t n_layers, const optk_ml_layer_t* layers,
                    int32_t n_segments, const optk_ml_segment_t* segments, double* reflectivity_s,
                    double* reflectivity_p, double* transmissivity_s, double* transmissivity_p, void* stream) {
    DeviceScope device_scope(stream, "optk_multilayer");
    if (!input || !layers || n_layers < 1 || n_layers > OPTK_ML_MAX_LAYERS) {
        set_error("optk_multilayer: need 1..%d layers (the last one is the substrate)", OPTK_ML_MAX_LAYERS);
        return OPTK_ERR_INVALID;
    }
    if (n_segments < 0 || n_segments > 32 || (n_segments > 0 && !segments)) {
        set_error("optk_multilayer: n_segments must be in 0..32");
        return OPTK_ERR_INVALID;
    }
    if (input->n_axes < 0 || input->n_axes > OPTK_ML_MAX_AXES) {
        set_error("optk_multilayer: n_axes must be in 0..%d", OPTK_ML_MAX_AXES);
        return OPTK_ERR_INVALID;
    }
    if (!input->wavelength || !input->direction_re || !input->n_re) {
        set_error("optk_multilayer: wavelength, direction_re and n_re are required");
        return OPTK_ERR_INVALID;
    }
    static thread_local MultilayerParams P;
    P.in = *input;
    long long n = 1;
    for (int a = 0; a < input->n_axes; ++a) n *= input->dims[a];
    if (n > 0xffffffffLL) {
        set_error("optk_multilayer: %lld evaluations exceed 2^32 - 1; split an axis", n);
        return OPTK_ERR_INVALID;
    }
    for (int a = 0; a < OPTK_ML_MAX_AXES; ++a)
        P.div[a] = make_fastdiv(a < input->n_axes ? (uint32_t)input->dims[a] : 1u);
    P.n_eval = n;
    P.n_layers = n_layers;
    P.n_segments = n_segments;
    for (int g = 0; g < n_segments; ++g) {
        const optk_ml_segment_t& s = segments[g];
        if (s.first < 0 || s.count < 0 || s.first + s.count > n_layers - 1 || s.repeat < 0) {
            set_error("optk_multilayer: segment %d is out of range", g);
            return OPTK_ERR_INVALID;
        }
        P.segments[g] = s;
    }
    for (int j = 0; j < n_layers; ++j)
        if (!layers[j].n_re) {
            set_error("optk_multilayer: layer %d has no index of refraction", j);
            return OPTK_ERR_INVALID;
        }

    // layer table -> device (stream ordered; the staging buffer is kept per thread)
    static thread_local DeviceBuffer table_dev;
    static thread_local std::vector<LayerDev> table_host;
    static thread_local LayerDev* table_pinned = nullptr;
    static thread_local size_t table_pinned_count = 0;
    if (table_pinned_count < (size_t)n_layers) {
        if (table_pinned) cudaFreeHost(table_pinned);
        OPTK_CUDA(cudaMallocHost((void**)&table_pinned, sizeof(LayerDev) * OPTK_ML_MAX_LAYERS));
        table_pinned_count = OPTK_ML_MAX_LAYERS;
    }
    int rc = table_dev.ensure(sizeof(LayerDev) * OPTK_ML_MAX_LAYERS);
    if (rc) return rc;
    // the pinned staging table may still be in flight from a previous call on this stream
    OPTK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int j = 0; j < n_layers; ++j) {
        LayerDev& d = table_pinned[j];
        const optk_ml_layer_t& l = layers[j];
        d.n_re = l.n_re;
        d.n_im = l.n_im;
        d.thickness = l.thickness;
        d.width = l.width;
        for (int a = 0; a < OPTK_ML_MAX_AXES; ++a) {
            d.n_stride[a] = l.n_stride[a];
            d.t_stride[a] = l.thickness ? l.thickness_stride[a] : 0;
            d.w_stride[a] = l.width ? l.width_stride[a] : 0;
        }
        d.profile_kind = l.width ? l.profile_kind : 0;
        d.pad = 0;
        auto single_axis = [&](const long long* st) -> int8_t {
            int axis = -1;
            for (int a = 0; a < input->n_axes; ++a) {
                if (st[a] != 0 && input->dims[a] > 1) {
                    if (axis >= 0) return (int8_t)-2;
                    axis = a;
                }
            }
            return (int8_t)axis;
        };
        d.n_axis = single_axis(d.n_stride);
        d.t_axis = single_axis(d.t_stride);
        d.w_axis = single_axis(d.w_stride);
    }
    OPTK_CUDA(cudaMemcpyAsync(table_dev.ptr, table_pinned, sizeof(LayerDev) * n_layers, cudaMemcpyHostToDevice,
                              (cudaStream_t)stream));
    P.layers = (const LayerDev*)table_dev.ptr;
    // cross-configuration reuse: an axis that only the thicknesses depend on
    P.reuse_axis = -1;
    for (int a = input->n_axes - 1; a >= 0 && P.reuse_axis < 0; --a) {
        if (input->dims[a] < 2) continue;
        bool only_thickness = input->wavelength_stride[a] == 0 && input->direction_stride[a] == 0 &&
                              input->n_stride[a] == 0;
        bool any_thickness = false;
        for (int j = 0; j < n_layers && only_thickness; ++j) {
            only_thickness = table_pinned[j].n_stride[a] == 0 && table_pinned[j].w_stride[a] == 0;
            any_thickness = any_thickness || table_pinned[j].t_stride[a] != 0;
        }
        if (only_thickness && any_thickness) P.reuse_axis = a;
    }
    P.r_s = reflectivity_s;
    P.r_p = reflectivity_p;
    P.t_s = transmissivity_s;
    P.t_p = transmissivity_p;
    return launch_multilayer(P, (cudaStream_t)stream);
}

OPTK_API int optk_measure_fp64_peak(double* flops_per_second, void* stream) {
    DeviceScope device_scope(stream);
    if (!flops_per_second) {
        set_error("optk_measure_fp64_peak: NULL argument");
        return OPTK_ERR_INVALID;
    }
    return measure_fp64_peak(flops_per_second, (cudaStream_t)stream);
}

OPTK_API int optk_measure_soa_copy(int64_t n_rays, double* gbytes_per_second, void* stream) {
    DeviceScope device_scope(stream);
    if (!gbytes_per_second || n_rays < 2) {
        set_error("optk_measure_soa_copy: bad arguments");
        return OPTK_ERR_INVALID;
    }
    return measure_soa_copy(n_rays, gbytes_per_second, (cudaStream_t)stream);
}

OPTK_API int optk_electrons_measured(int32_t n_plane, int32_t n_x, int32_t n_y, const optk_ccd_plane_t* planes,
                                     const int64_t* photons, uint64_t* electrons, int32_t wrap, uint64_t seed,
                                     void* stream) {
    DeviceScope device_scope(stream, "optk_electrons_measured");
    if (n_plane < 0 || n_x < 0 || n_y < 0 || (!planes && n_plane) || !photons || !electrons) {
        set_error("optk_electrons_measured: bad arguments");
        return OPTK_ERR_INVALID;
    }
    if ((long long)n_plane * n_x * n_y == 0) return OPTK_OK;
    for (int i = 0; i < n_plane; ++i) {
        const optk_ccd_plane_t& P = planes[i];
        if (P.energy <= 50.0 && (P.n_pmf < 1 || !P.cmf || !P.n_values)) {
            set_error("optk_electrons_measured: plane %d (%.3g eV) needs the pair-number distribution", i, P.energy);
            return OPTK_ERR_INVALID;
        }
        if (!(P.thickness_substrate >= 0) || !(P.absorption >= 0) || !(P.energy_pair_inf > 0)) {
            set_error("optk_electrons_measured: plane %d has a negative thickness / absorption or no pair energy", i);
            return OPTK_ERR_INVALID;
        }
    }
    optk_ccd_plane_t* device_planes = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    OPTK_CUDA(cudaMallocAsync((void**)&device_planes, sizeof(optk_ccd_plane_t) * n_plane, s));
    cudaError_t e = cudaMemcpyAsync(device_planes, planes, sizeof(optk_ccd_plane_t) * n_plane, cudaMemcpyHostToDevice, s);
    int rc = e == cudaSuccess ? OPTK_OK : cuda_fail(e, "cudaMemcpyAsync (planes)");
    // the record array is pageable host memory: the copy above has consumed it when it returns
    if (rc == OPTK_OK)
        rc = launch_electrons(n_plane, n_x, n_y, device_planes, (const long long*)photons, (unsigned long long*)electrons,
                              wrap ? 1 : 0, seed, s);
    cudaFreeAsync(device_planes, s);
    return rc;
}

OPTK_API int optk_host_register(void* data, int64_t n_bytes) {
    if (!data || n_bytes <= 0) {
        set_error("optk_host_register: bad arguments");
        return OPTK_ERR_INVALID;
    }
    OPTK_CUDA(cudaHostRegister(data, (size_t)n_bytes, cudaHostRegisterPortable));
    return OPTK_OK;
}

OPTK_API int optk_host_unregister(void* data) {
    if (!data) {
        set_error("optk_host_unregister: NULL argument");
        return OPTK_ERR_INVALID;
    }
    OPTK_CUDA(cudaHostUnregister(data));
    return OPTK_OK;
}

OPTK_API int optk_memcpy_async(void* dst, const void* src, int64_t n_bytes, void* stream) {
    DeviceScope device_scope(stream);
    if (!dst || !src || n_bytes < 0) {
        set_error("optk_memcpy_async: bad arguments");
        return OPTK_ERR_INVALID;
    }
    if (n_bytes == 0) return OPTK_OK;
    OPTK_CUDA(cudaMemcpyAsync(dst, src, (size_t)n_bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return OPTK_OK;
}

OPTK_API int optk_enable_peer_access(int32_t peer_device) {
    int current = -1;
    OPTK_CUDA(cudaGetDevice(&current));
    if (peer_device == current) return OPTK_OK;
    int can = 0;
    OPTK_CUDA(cudaDeviceCanAccessPeer(&can, current, peer_device));
    if (!can) {
        set_error("device %d cannot access device %d as a peer", current, peer_device);
        return OPTK_ERR_UNSUPPORTED;
    }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return OPTK_OK;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    return OPTK_OK;
}

}  // extern "C"
