// Run-time specialisation of the streamlined trace kernels for one system.
//
// The table-driven kernels decide per surface and per ray pair which sag, material, ruling
// and aperture code to run and read every surface parameter through an indexed constant load.
// For a long job that is worth compiling away: NVRTC builds the same trace_body with the walk
// written out as surface_full<2, EFF, FixedKinds<...>>(P.surf[k], ...) for compile-time k --
// no kind or flag tests, parameters as direct constant-bank operands, no spills (measured on
// cfg 2: fused image 7.8e10 -> 1.0e11 intercepts/s).  The device headers are embedded in the
// library (embedded_sources.inc); libnvrtc and libcuda are loaded lazily with dlopen so that
// the library still loads on a machine without a driver.  Any failure falls back to the
// table-driven CUDA kernels (never to a CPU path).
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "params.cuh"
#include "embedded_sources.inc"

namespace optk {

namespace {

typedef int (*nvrtcCreateProgram_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
typedef int (*nvrtcCompileProgram_t)(void*, int, const char* const*);
typedef int (*nvrtcGetSize_t)(void*, size_t*);
typedef int (*nvrtcGetData_t)(void*, char*);
typedef int (*nvrtcDestroyProgram_t)(void**);
typedef int (*cuModuleLoadData_t)(void**, const void*);
typedef int (*cuModuleGetFunction_t)(void**, void*, const char*);
typedef int (*cuLaunchKernel_t)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*,
                                void**, void**);

struct Api {
    bool tried = false, ok = false;
    nvrtcCreateProgram_t create = nullptr;
    nvrtcCompileProgram_t compile = nullptr;
    nvrtcGetSize_t cubin_size = nullptr, log_size = nullptr;
    nvrtcGetData_t cubin = nullptr, log = nullptr;
    nvrtcDestroyProgram_t destroy = nullptr;
    cuModuleLoadData_t module_load = nullptr;
    cuModuleGetFunction_t get_function = nullptr;
    cuLaunchKernel_t launch = nullptr;
};

Api g_api;
std::mutex g_mutex;
std::map<std::string, void*> g_kernels;  // signature -> CUfunction (nullptr: compilation failed, do not retry)
int g_not_on_disk_tag;
void* const kNotOnDisk = &g_not_on_disk_tag;  // looked in the disk cache, nothing there; not compiled yet
int g_mode = -2;                         // -2: read OPTK_JIT on first use; -1 auto; 0 off; 1 always
long long g_compiled = 0, g_loaded = 0;  // kernels compiled / loaded from the disk cache

void* open_first(const char* const* names) {
    for (; *names; ++names)
        if (void* h = dlopen(*names, RTLD_NOW | RTLD_GLOBAL)) return h;
    return nullptr;
}

bool load_api() {
    if (g_api.tried) return g_api.ok;
    g_api.tried = true;
    static const char* const rtc_names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so", nullptr};
    static const char* const cuda_names[] = {"libcuda.so.1", "libcuda.so", nullptr};
    void* rtc = open_first(rtc_names);
    void* cu = open_first(cuda_names);
    if (!rtc || !cu) return false;
#define OPTK_SYM(handle, field, name)                   \
    g_api.field = (decltype(g_api.field))dlsym(handle, name); \
    if (!g_api.field) return false;
    OPTK_SYM(rtc, create, "nvrtcCreateProgram")
    OPTK_SYM(rtc, compile, "nvrtcCompileProgram")
    OPTK_SYM(rtc, cubin_size, "nvrtcGetCUBINSize")
    OPTK_SYM(rtc, cubin, "nvrtcGetCUBIN")
    OPTK_SYM(rtc, log_size, "nvrtcGetProgramLogSize")
    OPTK_SYM(rtc, log, "nvrtcGetProgramLog")
    OPTK_SYM(rtc, destroy, "nvrtcDestroyProgram")
    OPTK_SYM(cu, module_load, "cuModuleLoadData")
    OPTK_SYM(cu, get_function, "cuModuleGetFunction")
    OPTK_SYM(cu, launch, "cuLaunchKernel")
#undef OPTK_SYM
    g_api.ok = true;
    return true;
}

bool verbose() {
    static const bool v = getenv("OPTK_JIT_VERBOSE") != nullptr;
    return v;
}


// ---- disk cache of compiled kernels: compilation costs ~1.5 s, more than a 1e10-ray job saves, so
// the cubin is kept across processes in $OPTK_JIT_CACHE (default ~/.cache/optika_b200/jit; set it
// to the empty string to switch the cache off), keyed by a hash of the generated source AND of
// the embedded device headers.
uint64_t fnv1a(const char* data, size_t n, uint64_t h = 1469598103934665603ULL) {
    for (size_t i = 0; i < n; ++i) {
        h ^= (unsigned char)data[i];
        h *= 1099511628211ULL;
    }
    return h;
}

std::string cache_path(const std::string& src) {
    const char* dir = getenv("OPTK_JIT_CACHE");
    std::string base;
    if (dir) {
        if (!*dir) return std::string();
        base = dir;
    } else {
        const char* home = getenv("HOME");
        if (!home || !*home) return std::string();
        base = std::string(home) + "/.cache/optika_b200/jit";
    }
    uint64_t h = fnv1a(src.data(), src.size());
    for (int k = 0; k < kHeaderCount; ++k) h = fnv1a(kHeaderSources[k], strlen(kHeaderSources[k]), h);
    char name[64];
    snprintf(name, sizeof(name), "/optk_sm100a_abi%d_%016llx.cubin", OPTK_ABI_VERSION, (unsigned long long)h);
    // mkdir -p
    for (size_t i = 1; i <= base.size(); ++i)
        if (i == base.size() || base[i] == '/') mkdir(base.substr(0, i).c_str(), 0755);
    return base + name;
}

bool read_file(const std::string& path, std::vector<char>& data) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    bool ok = n > 0;
    if (ok) {
        data.resize((size_t)n);
        ok = fread(data.data(), 1, (size_t)n, f) == (size_t)n;
    }
    fclose(f);
    return ok;
}

void write_file_atomically(const std::string& path, const std::vector<char>& data) {
    char tmp[32];
    snprintf(tmp, sizeof(tmp), ".%d.tmp", (int)getpid());
    const std::string partial = path + tmp;
    FILE* f = fopen(partial.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (!ok || rename(partial.c_str(), path.c_str()) != 0) remove(partial.c_str());
}

}  // namespace

void jit_set_mode(int mode) {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_mode = mode < -1 ? -1 : (mode > 1 ? 1 : mode);
}

long long jit_compiled_count() { return g_compiled + g_loaded; }

// The generated translation unit for one (system signature, kernel variant).
std::string jit_source(const TraceParams& P, const JitVariant& v, std::string* key_out) {
    char line[1024];
    std::string walk;
    for (int k = 0; k < P.n_surf; ++k) {
        const optk_surface_t& S = P.surf[k];
        const bool eff = S.material_efficiency != OPTK_EFF_UNIT || S.ruling_profile != OPTK_PROFILE_IDEAL;
        // loop lengths and integer exponents that shape the code are compile-time constants as well:
        // polygon vertices (the edge loop unrolls) and the polynomial ruling's powers (x^p by
        // repeated multiplication folds to the few multiplications it is)
        const int nv = S.aperture_kind == OPTK_APERTURE_POLYGON ? S.n_vertices : 0;
        const int nc = S.ruling_kind == OPTK_RULING_POLYNOMIAL ? S.n_coeff : 0;
        int pw[OPTK_MAX_COEFF] = {0};
        for (int j = 0; j < nc && j < OPTK_MAX_COEFF; ++j) pw[j] = S.ruling_power[j];
        snprintf(line, sizeof(line),
                 "    surface_full<2, %s, FixedKinds<%d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d>>"
                 "(P.surf[%d], r, newton_iterations, state);\n",
                 eff ? "true" : "false", S.sag_kind, S.material_kind, S.ruling_kind, S.aperture_kind, S.flags, nv, nc,
                 pw[0], pw[1], pw[2], pw[3], pw[4], pw[5], pw[6], pw[7], k);
        walk += line;
    }
    snprintf(line, sizeof(line), "// variant: dense %d vec %d image %d grid %d minb %d\n", v.dense, v.vec, v.image, v.grid,
             v.minb);
    std::string src = line;
    if (v.groups) src += "#define OPTK_JIT_GROUPS 1\n";
    static const int bin_direct = [] {
        const char* e = getenv("OPTK_BIN_DIRECT");
        return e ? atoi(e) : 1;
    }();
    if (!bin_direct) src += "#define OPTK_JIT_BIN_DIRECT 0\n";
    if (v.image_flags & 0x100) {
        snprintf(line, sizeof(line), "#define OPTK_JIT_IMAGE_FLAGS 0x%x\n", v.image_flags & 0xff);
        src += line;
    }
    if (v.grid_flags & 0x100) {
        snprintf(line, sizeof(line), "#define OPTK_JIT_GRID_FLAGS 0x%x\n", v.grid_flags & 0xff);
        src += line;
    }
    // strided ("broadcast") input: the layout is part of the kernel -- see load_rays
    if (!v.dense && !v.grid && P.offsets32 && !P.in.normal[0]) {
        unsigned long long lo = 0, hi = 0;
        unsigned outer = 0;
        const int n_axes = P.in.n_axes, first = n_axes - P.n_inner_axes;
        const bool has_mask = P.in.unvignetted != nullptr;
        for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) {
            if (f == OPTK_NUM_FIELDS && !has_mask) continue;
            for (int a = 0; a < n_axes && a < OPTK_MAX_AXES; ++a) {
                if (P.stride32[f][a] == 0) continue;
                if (a < first) outer |= 1u << f;
                if (f < 8) lo |= 1ULL << (f * 8 + a);
                else hi |= 1ULL << ((f - 8) * 8 + a);
            }
        }
        snprintf(line, sizeof(line),
                 "#define OPTK_JIT_LAYOUT 1\n#define OPTK_JIT_N_AXES %d\n#define OPTK_JIT_FIRST %d\n"
                 "#define OPTK_JIT_HAS_MASK %d\n"
                 "#define OPTK_JIT_VARIES(f, a) ((((f) < 8 ? 0x%llxULL >> (((f) & 7) * 8 + (a)) : 0x%llxULL >> (((f) & 7) * 8 + (a))) & 1) != 0)\n"
                 "#define OPTK_JIT_VARIES_OUTER(f) (((0x%xu >> (f)) & 1) != 0)\n",
                 n_axes, first, has_mask ? 1 : 0, lo, hi, outer);
        src += line;
    }
    src +=
        "#define OPTK_JIT_WALK 1\n"
        "#include \"trace_impl.cuh\"\n"
        "namespace optk {\n"
        "__device__ __forceinline__ void optk_jit_walk(const TraceParams& P, Ray (&r)[2], unsigned& newton_iterations,\n"
        "                                              WalkState& state) {\n";
    src += walk;
    src += "}\n}  // namespace optk\n";
    snprintf(line, sizeof(line),
             "extern \"C\" __global__ void __launch_bounds__(256, %d) optk_jit_kernel(const __grid_constant__ optk::TraceParams P) {\n"
             "    optk::trace_body<2, true, %s, %s, false, %s, %d, false, 1>(P);\n}\n",
             v.minb, v.dense ? "true" : "false", v.vec ? "true" : "false", v.image ? "true" : "false", v.grid);
    src += line;
    if (key_out) *key_out = src;  // the source is its own cache key
    return src;
}

// CUfunction for this launch, compiling it on first use; nullptr = use the table-driven kernel.
void* jit_kernel(const TraceParams& P, const JitVariant& v) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_mode == -2) {
        const char* e = getenv("OPTK_JIT");
        g_mode = e ? (atoi(e) > 0 ? 1 : (atoi(e) == 0 ? 0 : -1)) : -1;
    }
    if (g_mode == 0) return nullptr;
    // automatic: compile (~1.5 s) only for launches long enough that a production run of them
    // pays for it; a kernel that is already in the disk cache costs a file read and is taken for
    // launches from 2^17 rays on
    const bool may_compile = g_mode == 1 || P.n_rays >= (1LL << 25) || v.groups;  // group launches have no other kernel
    if (!may_compile && P.n_rays < (1LL << 17)) return nullptr;  // below ~1e5 rays the launch is latency either way
    std::string key;
    const std::string src = jit_source(P, v, &key);
    // a CUfunction belongs to the context of ONE device: the in-process cache is per device ordinal
    // (the cubin on disk is not)
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return nullptr;
    key = "device " + std::to_string(device) + "\n" + key;
    auto it = g_kernels.find(key);
    if (it != g_kernels.end() && it->second != kNotOnDisk) return it->second;
    if (it != g_kernels.end() && !may_compile) return nullptr;  // looked before: not cached, too short to compile
    const bool looked_on_disk = it != g_kernels.end();
    void* function = nullptr;
    g_kernels[key] = nullptr;
    if (!load_api()) {
        if (verbose()) fprintf(stderr, "optk jit: libnvrtc / libcuda not available, using the table-driven kernels\n");
        return nullptr;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (major != 10 || minor != 0) return nullptr;  // this library is sm_100a only
    cudaFree(nullptr);  // make sure the primary context is current for the driver API
    const std::string cached = cache_path(src);
    if (!cached.empty() && !looked_on_disk) {
        std::vector<char> cubin;
        void* module = nullptr;
        if (read_file(cached, cubin) && g_api.module_load(&module, cubin.data()) == 0 &&
            g_api.get_function(&function, module, "optk_jit_kernel") == 0) {
            if (verbose()) fprintf(stderr, "optk jit: loaded %s\n", cached.c_str());
            ++g_loaded;
            g_kernels[key] = function;
            return function;
        }
        function = nullptr;
    }
    if (!may_compile) {
        g_kernels[key] = kNotOnDisk;
        return nullptr;
    }
    void* prog = nullptr;
    if (g_api.create(&prog, src.c_str(), "optk_jit.cu", kHeaderCount, kHeaderSources, kHeaderNames) != 0) return nullptr;
    const char* options[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo"};
    const int rc = g_api.compile(prog, 4, options);
    if (rc != 0) {
        size_t n = 0;
        g_api.log_size(prog, &n);
        std::vector<char> log(n + 1, 0);
        g_api.log(prog, log.data());
        fprintf(stderr, "optk jit: compilation failed (%d), using the table-driven kernels\n%s\n", rc, log.data());
        g_api.destroy(&prog);
        return nullptr;
    }
    size_t n = 0;
    g_api.cubin_size(prog, &n);
    std::vector<char> cubin(n);
    g_api.cubin(prog, cubin.data());
    g_api.destroy(&prog);
    cudaFree(nullptr);  // make sure the primary context is current for the driver API
    void* module = nullptr;
    if (g_api.module_load(&module, cubin.data()) != 0 || g_api.get_function(&function, module, "optk_jit_kernel") != 0) {
        fprintf(stderr, "optk jit: loading the compiled kernel failed, using the table-driven kernels\n");
        return nullptr;
    }
    ++g_compiled;
    if (!cached.empty()) write_file_atomically(cached, cubin);
    if (verbose()) fprintf(stderr, "optk jit: compiled a kernel for %d surfaces (%zu bytes)\n", P.n_surf, n);
    g_kernels[key] = function;
    return function;
}

int jit_launch(void* function, const TraceParams& P, unsigned grid, cudaStream_t stream) {
    void* args[] = {(void*)&P};
    const int rc = g_api.launch(function, grid, 1, 1, 256, 1, 1, 0, (void*)stream, args, nullptr);
    if (rc != 0) {
        set_error("optk jit: cuLaunchKernel failed (%d)", rc);
        return OPTK_ERR_CUDA;
    }
    return OPTK_OK;
}

}  // namespace optk
