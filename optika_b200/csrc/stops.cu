// Stop solver on the device: which ray leaves a given point (or
// direction) on the first stop surface and arrives at a given point (or direction) on the
// other one.
//
// The reference solves this with na.optimize.root_newton and a finite-difference na.jacobian
// around a function that traces a small batch through the sub-system between the two stops
// (SequentialSystem._calc_rayfunction_stops_only, optika/systems/_sequential.py:551-606;
// residual function _ray_error, :363-394).  On the host that is three traces and a handful
// of NumPy operations per iteration, i.e. launch- and interpreter-bound (30 ms per solve with
// every trace on the device).  Here one thread owns one unknown ray and runs the whole 2-D
// Newton iteration -- residual, two forward differences, 2x2 solve -- around the same surface
// walk the trace kernels use; one launch per configuration replaces ~15 traces.
//
// Differences to the host iteration (optika_b200/_stops.py::_newton), both below the solver's
// own tolerance: a ray stops iterating when ITS residual is below max_abs_error (the host
// keeps updating every ray until the last one has converged).
#include "trace_impl.cuh"

namespace optk {

struct StopParams {
    const double* wavelength;
    const double* fixed[3];   // the vector of the ray that is NOT solved for (global frame)
    const double* target[2];  // where the traced ray must arrive (local frame of the last surface)
    double* x;                // in: initial guess; out: solution (global frame components)
    double* y;
    double* z;                // out: third component, z(x, y)
    unsigned int* n_unconverged;
    long long n;
    double step;
    double max_abs_error;
    int variable;  // OPTK_STOP_DIRECTION / OPTK_STOP_POSITION
    int target_kind;
    int max_iterations;
    int sag_slot;  // index into TraceParams::surf of the first stop surface (its sag gives z), or -1
};

namespace {

// third component of the free vector (optika/systems/_sequential.py:436-437, 468-469)
__device__ __forceinline__ double free_z(const TraceParams& P, const StopParams& Q, double x, double y) {
    if (Q.variable == OPTK_STOP_DIRECTION) return sqrt(1.0 - (x * x + y * y));
    return Q.sag_slot >= 0 ? sag_value(P.surf[Q.sag_slot], x, y) : 0.0;
}

// _ray_error: trace the trial ray through the sub-system and compare with the target
__device__ __forceinline__ void residual(const TraceParams& P, const StopParams& Q, const Ray& base, double x, double y,
                                         double tx, double ty, double& fx, double& fy) {
    Ray r = base;
    const double z = free_z(P, Q, x, y);
    if (Q.variable == OPTK_STOP_DIRECTION) {
        r.dx = x; r.dy = y; r.dz = z;
    } else {
        r.px = x; r.py = y; r.pz = z;
    }
    unsigned iterations = 0;
    double cos_incidence = 0.0;
    for (int s = 0; s < P.n_surf; ++s) surface_generic(P.surf[s], r, iterations, false, 0.0, 0.0, -1.0, cos_incidence);
    // the walk leaves the ray in the local frame of the last surface (OPTK_F_LOCAL_OUT is set on
    // the copy of that surface): transformation_last.inverse(rays), :383-384
    if (Q.target_kind == OPTK_STOP_DIRECTION) {
        fx = r.dx - tx;
        fy = r.dy - ty;
    } else {
        fx = r.px - tx;
        fy = r.py - ty;
    }
}

__global__ void __launch_bounds__(128) stop_newton_kernel(const __grid_constant__ TraceParams P,
                                                          const __grid_constant__ StopParams Q) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.n) return;
    Ray base{Q.wavelength[i], 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0, true};
    if (Q.variable == OPTK_STOP_DIRECTION) {
        base.px = Q.fixed[0][i]; base.py = Q.fixed[1][i]; base.pz = Q.fixed[2][i];
    } else {
        base.dx = Q.fixed[0][i]; base.dy = Q.fixed[1][i]; base.dz = Q.fixed[2][i];
    }
    const double tx = Q.target[0][i], ty = Q.target[1][i];
    double x = Q.x[i], y = Q.y[i];
    const double h = Q.step;
    bool converged = false;
    for (int it = 0; it < Q.max_iterations; ++it) {
        double fx, fy;
        residual(P, Q, base, x, y, tx, ty, fx, fy);
        if (fabs(fx) <= Q.max_abs_error && fabs(fy) <= Q.max_abs_error) {
            converged = true;
            break;
        }
        double ax, ay, bx, by;
        residual(P, Q, base, x + h, y, tx, ty, ax, ay);
        residual(P, Q, base, x, y + h, tx, ty, bx, by);
        const double j11 = (ax - fx) / h, j21 = (ay - fy) / h;
        const double j12 = (bx - fx) / h, j22 = (by - fy) / h;
        const double det = j11 * j22 - j12 * j21;
        x = x - (j22 * fx - j12 * fy) / det;
        y = y - (-j21 * fx + j11 * fy) / det;
    }
    Q.x[i] = x;
    Q.y[i] = y;
    Q.z[i] = free_z(P, Q, x, y);
    if (!converged) atomicAdd(Q.n_unconverged, 1u);
}

}  // namespace

int launch_stop_newton(const TraceParams& P, const optk_stop_problem_t& problem, int sag_slot, long long n,
                       const double* wavelength, const double* const fixed[3], const double* const target[2], double* x,
                       double* y, double* z, unsigned int* n_unconverged, cudaStream_t stream) {
    if (n == 0) return OPTK_OK;
    StopParams Q;
    Q.wavelength = wavelength;
    for (int k = 0; k < 3; ++k) Q.fixed[k] = fixed[k];
    for (int k = 0; k < 2; ++k) Q.target[k] = target[k];
    Q.x = x;
    Q.y = y;
    Q.z = z;
    Q.n_unconverged = n_unconverged;
    Q.n = n;
    Q.step = problem.step;
    Q.max_abs_error = problem.max_abs_error;
    Q.variable = problem.variable;
    Q.target_kind = problem.target;
    Q.max_iterations = problem.max_iterations;
    Q.sag_slot = sag_slot;
    const int block = 128;
    const long long grid = (n + block - 1) / block;
    stop_newton_kernel<<<(unsigned)grid, block, 0, stream>>>(P, Q);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

}  // namespace optk
