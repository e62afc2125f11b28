// Kernel 1: the fused sequential raytrace (sm_100a, fp64).
//
// One thread owns one ray and walks the whole surface list, which arrives as a
// kernel parameter (constant bank, broadcast to the warp).  Ray state stays in
// registers; per surface the thread runs AbstractSurface.propagate_rays
// (optika/surfaces.py:123-198): global->local, sag intercept + Beer-Lambert,
// sag normal, grating equation, index / wavelength rescale, vector Snell,
// aperture clip, local->global.  Rays touch HBM on entry and exit only (plus one
// store per surface when accumulating), and optionally never leave the chip:
// the final rays can be binned straight into the detector image.
#pragma once
// resident CTAs per SM the streamlined (two rays per thread) kernels are compiled for: 3 -> 80
// registers with ~190 B of spills, 24 warps; 2 -> up to 128 registers, no spills, 16 warps
#ifndef OPTK_FULL_MINB
#define OPTK_FULL_MINB 3
#endif
#include "common.cuh"
#include "bin.cuh"
#include "params.cuh"
#ifndef __CUDACC_RTC__
#include <cstdlib>
#endif

namespace optk {

struct Ray {
    double w, px, py, pz, dx, dy, dz, intensity, att, n;
    bool unv;
};

// Where the element kinds and flags of a surface come from: the table (run time) or template
// arguments (a kernel compiled at run time for one system, jit.cu: every kind test folds away,
// polygon and polynomial loops unroll, and the rarer element kinds are inlined for both rays).
struct TableKinds {
    static constexpr bool fixed = false;
    static __device__ __forceinline__ int sag(const optk_surface_t& S) { return S.sag_kind; }
    static __device__ __forceinline__ int material(const optk_surface_t& S) { return S.material_kind; }
    static __device__ __forceinline__ int ruling(const optk_surface_t& S) { return S.ruling_kind; }
    static __device__ __forceinline__ int aperture(const optk_surface_t& S) { return S.aperture_kind; }
    static __device__ __forceinline__ int flags(const optk_surface_t& S) { return S.flags; }
    static __device__ __forceinline__ int n_vertices(const optk_surface_t& S) { return S.n_vertices; }
    static __device__ __forceinline__ int n_coeff(const optk_surface_t& S) { return S.n_coeff; }
    static __device__ __forceinline__ int power(const optk_surface_t& S, int k) { return S.ruling_power[k]; }
};
template <int SAG, int MATERIAL, int RULING, int APERTURE, int FLAGS, int NV = 0, int NC = 0, int P0 = 0, int P1 = 0,
          int P2 = 0, int P3 = 0, int P4 = 0, int P5 = 0, int P6 = 0, int P7 = 0>
struct FixedKinds {
    static constexpr bool fixed = true;
    static __device__ __forceinline__ constexpr int sag(const optk_surface_t&) { return SAG; }
    static __device__ __forceinline__ constexpr int material(const optk_surface_t&) { return MATERIAL; }
    static __device__ __forceinline__ constexpr int ruling(const optk_surface_t&) { return RULING; }
    static __device__ __forceinline__ constexpr int aperture(const optk_surface_t&) { return APERTURE; }
    static __device__ __forceinline__ constexpr int flags(const optk_surface_t&) { return FLAGS; }
    static __device__ __forceinline__ constexpr int n_vertices(const optk_surface_t&) { return NV; }
    static __device__ __forceinline__ constexpr int n_coeff(const optk_surface_t&) { return NC; }
    static __device__ __forceinline__ constexpr int power(const optk_surface_t&, int k) {
        return k == 0 ? P0 : k == 1 ? P1 : k == 2 ? P2 : k == 3 ? P3 : k == 4 ? P4 : k == 5 ? P5 : k == 6 ? P6 : P7;
    }
};

// ---------------------------------------------------------------------------
// sag profiles, evaluated in the sag's own frame
// ---------------------------------------------------------------------------

// sign0(v) * root for root >= 0, in 3 instructions: copysign unless v == 0 (NaN v gives NaN upstream anyway)
__device__ __forceinline__ double signed_root(double v, double root) {
    return (v == 0.0) ? v * root : copysign(root, v);
}

// optika/sags/_parabolic.py:142-151: root of a t^2 + 2 h t + c = 0 with
//   a = ux^2 + uy^2,  h = ox ux + oy uy - 2 f uz,  c = ox^2 + oy^2 - 4 f oz;
// the reference takes t = (-h - s sqrt(h^2 - a c)) / a with s = sign(f uz), and for
// a <= 1e-10 its paraxial form t = c / (4 f uz) (which drops ox ux + oy uy; kept verbatim).
// For near-axial rays -h and s sqrt() nearly cancel (the reference loses ~eps |2f| / a to
// rounding there, see DESIGN.md "conditioning"); the same root is evaluated in its
// cancellation-free form c / (-h + s sqrt()).  One division, no branch.
__device__ __forceinline__ double parabola_intercept(double f, double ox, double oy, double oz, double ux, double uy,
                                                     double uz) {
    const double a = norm2_2(ux, uy);
    const double fuz = f * uz;
    const double h = fma(ox, ux, fma(oy, uy, -2.0 * fuz));
    const double c = fma(ox, ox, fma(oy, oy, -4.0 * (f * oz)));
    const double root = signed_root(fuz, fsqrt(fma(h, h, -(a * c))));
    const bool general = a > 1e-10;
    const bool stable = -h * fuz >= 0.0;  // -h and root have the same sign (or one of them is zero)
    const double num = (general && !stable) ? (-h - root) : c;
    const double den = general ? (stable ? (-h + root) : a) : 4.0 * fuz;
    return fdiv_finite(num, den);
}

// Path length t to the surface for a ray o + t u (closed forms), or NaN/inf on a miss.
template <class K = TableKinds>
__device__ __forceinline__ double sag_intercept_closed(const optk_surface_t& S, double ox, double oy, double oz,
                                                       double ux, double uy, double uz) {
    switch (K::sag(S)) {
        case OPTK_SAG_FLAT:
            // optika/sags/_flat.py:58: d = -o.z / u.z
            return fdiv_finite(-oz, uz);
        case OPTK_SAG_SPHERICAL: {
            // optika/sags/_spherical.py:176-184
            const double r = S.sag[0];
            const double pz = oz - r;
            const double up = dot3(ux, uy, uz, ox, oy, pz);
            const double disc = fma(up, up, -(norm2_3(ox, oy, pz) - r * r));
            return -up - sign0(r * uz) * fsqrt(disc);
        }
        case OPTK_SAG_PARABOLIC:
            return parabola_intercept(S.sag[0], ox, oy, oz, ux, uy, uz);
        case OPTK_SAG_CONIC: {
            // optika/sags/_conic.py:126-162: A t^2 + B t + C = 0, both roots tested for the
            // vertex sheet, the smaller |t| wins.  root(-1) = (-B - sqrt)/(2A) and
            // root(+1) = (-B + sqrt)/(2A) are formed from q = -(B + sign(B) sqrt)/2 as q/A and
            // C/q so that neither suffers the cancellation of the textbook formula (A -> 0
            // for k -> -1 and near-axial rays).
            const double c = S.sag[3];
            const double kp1 = 1.0 + S.sag[1];
            const double a = c * (ux * ux + uy * uy + kp1 * uz * uz);
            const double b = 2.0 * (c * (ox * ux + oy * uy + kp1 * oz * uz) - uz);
            const double cc = c * (ox * ox + oy * oy + kp1 * oz * oz) - 2.0 * oz;
            const double disc = b * b - 4.0 * a * cc;
            const bool real = disc >= 0;
            const double sq_disc = fsqrt(real ? disc : 0.0);
            const bool degenerate = fabs(a) < 1e-12;
            double t_minus, t_plus;  // root(-1), root(+1)
            if (degenerate) {
                t_minus = t_plus = fdiv(-cc, b);
            } else {
                const double q = -0.5 * (b + copysign(sq_disc, b));
                const double t_q = fdiv(q, a), t_c = (q != 0.0) ? fdiv(cc, q) : t_q;
                // b >= 0: q/a = (-b - sqrt)/(2a) = root(-1);  b < 0: q/a = root(+1)
                t_minus = (b >= 0.0) ? t_q : t_c;
                t_plus = (b >= 0.0) ? t_c : t_q;
            }
            double t_root[2] = {t_minus, t_plus};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double t = t_root[k];
                const double x = ox + ux * t, y = oy + uy * t, z = oz + uz * t;
                const double r2 = x * x + y * y;
                const bool on_vertex_sheet = (z * (c * r2 - z)) >= 0;
                t_root[k] = (real && on_vertex_sheet) ? t : INFINITY;
            }
            return fabs(t_root[0]) <= fabs(t_root[1]) ? t_root[0] : t_root[1];
        }
        case OPTK_SAG_CYLINDRICAL: {
            // optika/sags/_cylindrical.py:126-152, cross products with a = y-hat written out
            const double r = S.sag[0];
            const double bx = -ox, bz = r - oz;
            const double ncx = -uz, ncz = ux;
            const double nca2 = ncx * ncx + ncz * ncz;
            const double negative_b = ncx * (-bz) + ncz * bx;
            const double dot = bx * ncx + bz * ncz;
            const double disc = nca2 * (r * r) - dot * dot;
            if (disc > 0) return fdiv(negative_b - sign0(r * uz) * fsqrt(disc), nca2);
            return fdiv_finite(-oz, uz);
        }
    }
    return NAN;
}

// Toroid: z and its gradient (optika/sags/_toroidal.py:38-88).
__device__ __forceinline__ void toroid_eval(double c, double r, double x, double y, double& z, double& dzdx,
                                            double& dzdy) {
    // g = sqrt(a) and f = sqrt(b) are needed together with 1 / g and 1 / f: one reciprocal square
    // root each (g = a / sqrt(a)) and one reciprocal for 1 / (1 + g), instead of two square roots,
    // two divisions and a reciprocal.  Outside the domain (a < 0 or b < 0) everything is NaN as in
    // the reference; exactly on its boundary (a = 0 or b = 0, a set of measure zero) the reference's
    // infinite slope becomes NaN (so the unrepaired reciprocal roots do: 0 * inf either way).
    const double y2 = y * y;
    const double a = fma(-(c * c), y2, 1.0);
    const double inv_g = frsqrt_raw(a);
    const double g = a * inv_g;
    const double zy = (c * y2) * frcp_raw(1.0 + g);  // 1 + g >= 1, or NaN
    const double rz = r - zy;
    const double b = fma(rz, rz, -(x * x));
    const double inv_f = frsqrt_raw(b);
    z = fma(-b, inv_f, r);
    dzdx = x * inv_f;
    dzdy = (rz * ((c * y) * inv_g)) * inv_f;
}

// Unit normal at a point of the sag frame (not rotated back, as in the reference).
template <class K = TableKinds>
__device__ __forceinline__ void sag_normal(const optk_surface_t& S, double x, double y, double& nx, double& ny,
                                           double& nz) {
    switch (K::sag(S)) {
        case OPTK_SAG_FLAT:
            nx = 0.0; ny = 0.0; nz = -1.0;  // optika/sags/_flat.py:43-47
            return;
        case OPTK_SAG_SPHERICAL: {
            // optika/sags/_spherical.py:141-146
            const double c = S.sag[3];
            nx = c * x;
            ny = c * y;
            nz = -fsqrt(1.0 - nx * nx - ny * ny);
            return;
        }
        case OPTK_SAG_CYLINDRICAL: {
            // optika/sags/_cylindrical.py:107-113
            nx = x * S.sag[3];
            ny = 0.0;
            nz = -fsqrt(1.0 - nx * nx);
            return;
        }
        case OPTK_SAG_PARABOLIC: {
            // optika/sags/_parabolic.py:56-63: (x, y, -R) / sqrt((x/R)^2 + (y/R)^2 + 1) / R
            const double ir = 0.5 * S.sag[3];  // 1 / (2 f)
            const double xr = x * ir, yr = y * ir;
            const double inv = frsqrt_raw(fma(xr, xr, fma(yr, yr, 1.0)));
            nx = xr * inv;
            ny = yr * inv;
            nz = -inv;
            return;
        }
        case OPTK_SAG_CONIC: {
            // optika/sags/_conic.py:69-81
            const double c = S.sag[3];
            const double ig = frsqrt_raw(1.0 - (1.0 + S.sag[1]) * c * c * (x * x + y * y));
            const double dzdx = c * x * ig, dzdy = c * y * ig;
            const double inv = frsqrt_raw(fma(dzdx, dzdx, fma(dzdy, dzdy, 1.0)));
            nx = dzdx * inv;
            ny = dzdy * inv;
            nz = -inv;
            return;
        }
        case OPTK_SAG_TOROIDAL: {
            // optika/sags/_toroidal.py:71-88
            double z, dzdx, dzdy;
            toroid_eval(S.sag[3], S.sag[2], x, y, z, dzdx, dzdy);
            const double inv = frsqrt_raw(fma(dzdx, dzdx, fma(dzdy, dzdy, 1.0)));
            nx = dzdx * inv;
            ny = dzdy * inv;
            nz = -inv;
            return;
        }
    }
    nx = ny = nz = NAN;
}

// z(x, y) in the sag frame, for OPTK_STAGE_SAG_OUT.
__device__ __forceinline__ double sag_value(const optk_surface_t& S, double x, double y) {
    switch (S.sag_kind) {
        case OPTK_SAG_FLAT:
            return 0.0;
        case OPTK_SAG_SPHERICAL: {
            const double c = S.sag[3];
            const double r2 = x * x + y * y;
            return c * r2 / (1.0 + sqrt(1.0 - c * c * r2));
        }
        case OPTK_SAG_CYLINDRICAL: {
            const double c = S.sag[3];
            const double r2 = x * x;
            return c * r2 / (1.0 + sqrt(1.0 - c * c * r2));
        }
        case OPTK_SAG_PARABOLIC:
        case OPTK_SAG_CONIC: {
            const double radius = S.sag_kind == OPTK_SAG_PARABOLIC ? 2.0 * S.sag[0] : S.sag[0];
            const double conic = S.sag_kind == OPTK_SAG_PARABOLIC ? -1.0 : S.sag[1];
            const double c = 1.0 / radius;
            const double r2 = x * x + y * y;
            return c * r2 / (1.0 + sqrt(1.0 - (1.0 + conic) * c * c * r2));
        }
        case OPTK_SAG_TOROIDAL: {
            double z, dzdx, dzdy;
            toroid_eval(S.sag[3], S.sag[2], x, y, z, dzdx, dzdy);
            return z;
        }
    }
    return NAN;
}

// ---------------------------------------------------------------------------
// rulings: kappa = spacing_(position, normal)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double ipow(double x, int p) {
    bool neg = p < 0;
    unsigned e = neg ? (unsigned)(-p) : (unsigned)p;
    double r = 1.0, b = x;
    while (e) {
        if (e & 1u) r *= b;
        b *= b;
        e >>= 1;
    }
    return neg ? 1.0 / r : r;
}

template <class K = TableKinds>
__device__ __forceinline__ void ruling_vector(const optk_surface_t& S, double px, double py, double pz, double nx,
                                              double ny, double nz, double& kx, double& ky, double& kz) {
    switch (K::ruling(S)) {
        case OPTK_RULING_CONSTANT: {
            // optika/rulings/_spacing.py:69-74
            const double c = S.ruling_coeff[0];
            kx = c * S.ruling_normal[0];
            ky = c * S.ruling_normal[1];
            kz = c * S.ruling_normal[2];
            return;
        }
        case OPTK_RULING_POLYNOMIAL: {
            // optika/rulings/_spacing.py:109-128
            if (K::flags(S) & OPTK_F_RULING_TRANSFORM) affine_forward(S.ruling_transform, px, py, pz, false);
            const double x = dot3(px, py, pz, S.ruling_normal[0], S.ruling_normal[1], S.ruling_normal[2]);
            double d = 0.0;
            if (K::fixed) {
#pragma unroll
                for (int k = 0; k < K::n_coeff(S); ++k) d = fma(S.ruling_coeff[k], ipow(x, K::power(S, k)), d);
            } else {
                for (int k = 0; k < K::n_coeff(S); ++k) d = fma(S.ruling_coeff[k], ipow(x, K::power(S, k)), d);
            }
            kx = d * S.ruling_normal[0];
            ky = d * S.ruling_normal[1];
            kz = d * S.ruling_normal[2];
            return;
        }
        case OPTK_RULING_HOLOGRAPHIC: {
            // optika/rulings/_spacing.py:295-328
            const double d1 = (K::flags(S) & OPTK_F_HOLO_DIVERGING_1) ? 1.0 : -1.0;
            const double d2 = (K::flags(S) & OPTK_F_HOLO_DIVERGING_2) ? 1.0 : -1.0;
            double ax = px - S.holo_x1[0], ay = py - S.holo_x1[1], az = pz - S.holo_x1[2];
            double bx = px - S.holo_x2[0], by = py - S.holo_x2[1], bz = pz - S.holo_x2[2];
            const double ia = d1 * frsqrt(ax * ax + ay * ay + az * az);
            const double ib = d2 * frsqrt(bx * bx + by * by + bz * bz);
            const double rx = ax * ia - bx * ib, ry = ay * ia - by * ib, rz = az * ia - bz * ib;
            // aq = n x dr
            const double qx = ny * rz - nz * ry, qy = nz * rx - nx * rz, qz = nx * ry - ny * rx;
            // spacing * (q/a) x n  with spacing = w / a   =>  (w / a^2) (aq x n)
            const double s = fdiv(S.holo_wavelength, qx * qx + qy * qy + qz * qz);
            kx = s * (qy * nz - qz * ny);
            ky = s * (qz * nx - qx * nz);
            kz = s * (qx * ny - qy * nx);
            return;
        }
    }
    kx = ky = kz = NAN;
}

// ---------------------------------------------------------------------------
// apertures (edge-sensitive: products and sums are not contracted)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double py_mod(double a, double b) {
    // numpy's float remainder: result takes the sign of the divisor
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

// One edge (x0, y0) -> (x1, y1) of the even-odd crossing test (see aperture_test).
__device__ __forceinline__ void polygon_edge(double x, double y, double x0, double y0, double x1, double y1,
                                             bool& inside, bool& on_edge) {
    const double ex = sub_rn(x1, x0), ey = sub_rn(y1, y0);
    const double cross = sub_rn(mul_rn(ex, sub_rn(y, y0)), mul_rn(ey, sub_rn(x, x0)));
    if (cross == 0.0) {  // on the supporting line (rare): is it on the segment?
        on_edge |= (fmin(x0, x1) <= x) && (x <= fmax(x0, x1)) && (fmin(y0, y1) <= y) && (y <= fmax(y0, y1));
    }
    // left of a straddling edge  <=>  the cross product has the sign of the edge's dy
    const bool straddles = (y0 > y) != (y1 > y);
    inside ^= straddles && ((cross > 0.0) == (ey > 0.0));
}

template <class K = TableKinds>
__device__ __forceinline__ bool polygon_even_odd(const optk_surface_t& S, double x, double y) {
    bool inside = false, on_edge = false;
    const int nv = K::n_vertices(S);
    double x0 = S.vertices_x[nv - 1], y0 = S.vertices_y[nv - 1];
    if constexpr (K::fixed) {
#pragma unroll
        for (int i = 0; i < nv; ++i) {
            polygon_edge(x, y, x0, y0, S.vertices_x[i], S.vertices_y[i], inside, on_edge);
            x0 = S.vertices_x[i];
            y0 = S.vertices_y[i];
        }
    } else {
        for (int i = 0; i < nv; ++i) {
            const double x1 = S.vertices_x[i], y1 = S.vertices_y[i];
            polygon_edge(x, y, x0, y0, x1, y1, inside, on_edge);
            x0 = x1;
            y0 = y1;
        }
    }
    return inside || on_edge;
}

// the same, out of line, for the rare points next to an edge line of a convex polygon
static __device__ __noinline__ bool polygon_exact(const optk_surface_t& S, double x, double y) {
    return polygon_even_odd<TableKinds>(S, x, y);
}

template <class K = TableKinds>
__device__ __forceinline__ bool aperture_test(const optk_surface_t& S, double x, double y, double z) {
    if (K::flags(S) & OPTK_F_APERTURE_TRANSFORM) affine_inverse(S.aperture_transform, x, y, z, false);
    bool mask = false;
    switch (K::aperture(S)) {
        case OPTK_APERTURE_CIRCULAR:
            // optika/apertures/_apertures.py:309: position.xy.length <= radius
            // sqrt is monotone and correctly rounded in the reference, so "sqrt(r2) <= radius" is
            // exactly "r2 <= T" with T = max{v : sqrt(v) <= radius}, which optk_system_create
            // stores in aperture[3]: same decision bit for bit, no square root per ray
            mask = add_rn(mul_rn(x, x), mul_rn(y, y)) <= S.aperture[3];
            break;
        case OPTK_APERTURE_RECTANGULAR:
            // optika/apertures/_apertures.py:962-963
            mask = (-S.aperture[0] <= x) && (x <= S.aperture[0]) && (-S.aperture[1] <= y) && (y <= S.aperture[1]);
            break;
        case OPTK_APERTURE_ELLIPTICAL: {
            // optika/apertures/_apertures.py:657
            const double a = x / S.aperture[0], b = y / S.aperture[1];
            mask = add_rn(mul_rn(a, a), mul_rn(b, b)) <= 1.0;
            break;
        }
        case OPTK_APERTURE_SECTOR: {
            // optika/apertures/_apertures.py:466-476
            const bool mask_radius = add_rn(mul_rn(x, x), mul_rn(y, y)) <= S.aperture[3];
            const double a0 = S.aperture[1], a1 = S.aperture[2];
            const double angle = atan2(y, x);
            const double two_pi = 6.283185307179586;
            const double ap = py_mod(angle, two_pi);
            const double an = py_mod(angle, -two_pi);
            mask = mask_radius && (((a0 < ap) && (ap < a1)) || ((a0 < an) && (an < a1)));
            break;
        }
        case OPTK_APERTURE_POLYGON: {
            // na.geometry.point_in_polygon (third party): even-odd crossing, boundary inside
            if (!(K::flags(S) & OPTK_F_APERTURE_ACTIVE)) return true;  // _apertures.py:751, 775-776
            if (K::flags(S) & OPTK_F_APERTURE_CONVEX) {
                // Strictly convex vertices (checked by optk_system_create): inside <=> on the inner side of
                // every edge line.  c_i = e_i x (p - v_i) in contracted arithmetic is within 1e-14 B^2 of the
                // reference's operation-by-operation value for |x|, |y| <= B; beyond +-1e-12 B^2 both have
                // the same, non-zero sign and the even-odd count is the geometric one.  Only points inside that
                // band of an edge line (a 1e-12 th of the rays) take the exact arithmetic, out of line.
                // 2 FMA + 2 compares per edge instead of 7 + 5, the edge constants are per thread.
                const double bound = S.aperture[0], band = S.aperture[1];
                const bool clockwise = K::flags(S) & OPTK_F_APERTURE_CLOCKWISE;
                const int nv = K::n_vertices(S);
                // a NaN coordinate is neither: it takes the exact arithmetic, whose comparisons decide as the
                // reference's do (a NaN x with y inside the polygon's range comes out INSIDE there)
                bool all_in = (fabs(x) <= bound) && (fabs(y) <= bound);
                bool any_out = (fabs(x) > bound) || (fabs(y) > bound);
                double x0 = S.vertices_x[nv - 1], y0 = S.vertices_y[nv - 1];
                auto edge = [&](int i) {
                    const double x1 = S.vertices_x[i], y1 = S.vertices_y[i];
                    const double ex = x1 - x0, ey = y1 - y0;
                    const double c0 = fma(ex, y0, -ey * x0);
                    const double cross = fma(ex, y, fma(-ey, x, -c0));
                    const double c = clockwise ? -cross : cross;
                    all_in = all_in && (c > band);
                    any_out = any_out || (c < -band);
                    x0 = x1;
                    y0 = y1;
                };
                if constexpr (K::fixed) {
#pragma unroll
                    for (int i = 0; i < nv; ++i) edge(i);
                } else {
                    for (int i = 0; i < nv; ++i) edge(i);
                }
                mask = all_in || (!any_out && polygon_exact(S, x, y));
                break;
            }
            mask = polygon_even_odd<K>(S, x, y);
            break;
        }
        default:
            return true;
    }
    if (K::flags(S) & OPTK_F_APERTURE_INVERTED) mask = !mask;
    if (!(K::flags(S) & OPTK_F_APERTURE_ACTIVE)) mask = true;
    return mask;
}

// ---------------------------------------------------------------------------
// Out-of-line versions of the rarer element kinds for the streamlined kernel.  The kernel
// body is instruction-cache bound once the generator, the surface walk and the binning
// network are all live (ncu: `no_instruction` is the top stall of the fused kernels), so
// only the kinds the BASELINE systems spend their time in (flat / sphere / parabola,
// constant rulings, rectangle / circle) are inlined twice (two rays per thread); the rest
// costs one call per ray and keeps the hot loop short.
// ---------------------------------------------------------------------------
// Results come back BY VALUE: a reference parameter would pin the caller's ray state to
// local memory for the whole kernel.
struct SagHit {
    double t, nx, ny, nz;
    unsigned iterations;
};
struct Vec3 {
    double x, y, z;
};

// The toroid intercept of sag_cold below for the R rays of a thread at once (run-time compiled
// kernels): the Newton chains of the rays interleave instead of running one after the other
// behind a call.  Each ray takes exactly the steps it takes in sag_cold (a ray that has converged
// keeps its value while the other one finishes), so the results are bit-identical.
template <int R>
__device__ __forceinline__ void toroid_intercept(const optk_surface_t& S, const Ray (&r)[R], double (&t)[R],
                                                 unsigned& iterations) {
    const double c = S.sag[3], rr = S.sag[2];
    const double curvature = 4.0 * fmax(fabs(c), fabs(frcp(rr)));
    bool active[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        t[k] = fdiv_newton(-r[k].pz, r[k].dz);
        if (!(fabs(t[k]) < OPTK_INF)) t[k] = 0.0;
        active[k] = true;
    }
    for (int it = 0; it < 64; ++it) {
        bool any = false;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            double z, dzdx, dzdy;
            toroid_eval(c, rr, fma(r[k].dx, t[k], r[k].px), fma(r[k].dy, t[k], r[k].py), z, dzdx, dzdy);
            const double f = fma(r[k].dz, t[k], r[k].pz) - z;
            const double df = r[k].dz - fma(dzdx, r[k].dx, dzdy * r[k].dy);
            const double step = fdiv_newton(f, df);
            const double tt = t[k] - step;
            const double tolerance = 1e-13 * (1.0 + fabs(tt));
            const bool done = !(fabs(step) > tolerance) || (step * step * curvature <= tolerance * fabs(df));
            t[k] = active[k] ? tt : t[k];
            iterations += active[k] ? 1u : 0u;
            active[k] = active[k] && !done;
            any = any || active[k];
        }
        if (!any) break;
    }
}

static __device__ __noinline__ SagHit sag_cold(const optk_surface_t& S, double qx, double qy, double qz, double vx,
                                               double vy, double vz) {
    SagHit hit;
    hit.iterations = 0;
    if (S.sag_kind == OPTK_SAG_TOROIDAL) {
        // AbstractSag.intercept, optika/sags/_abc.py:76-107: root of
        // f(t) = (o + t u).z - sag(o + t u) from t = 0.  Newton with the analytic
        // gradient, iterated to convergence (the reference's secant stops at
        // |step| < 1e-6 mm; see DESIGN.md "toroid intercept").
        const double c = S.sag[3], rr = S.sag[2];
        // Newton converges quadratically: e' ~ |f'' / (2 f')| e^2 with |f''| bounded by a few
        // times the larger curvature.  Once the step just taken predicts a next step below the
        // tolerance, that next evaluation (two square roots, two divisions) is skipped.
        const double curvature = 4.0 * fmax(fabs(c), fabs(frcp(rr)));
        // The reference starts at t = 0, and its first step lands next to the vertex plane.  Start
        // there directly (one division instead of a toroid evaluation); same root.
        double tt = fdiv_newton(-qz, vz);
        if (!(fabs(tt) < OPTK_INF)) tt = 0.0;
        for (int it = 0; it < 64; ++it) {
            double z, dzdx, dzdy;
            toroid_eval(c, rr, fma(vx, tt, qx), fma(vy, tt, qy), z, dzdx, dzdy);
            const double f = fma(vz, tt, qz) - z;
            const double df = vz - fma(dzdx, vx, dzdy * vy);
            const double step = fdiv_newton(f, df);
            tt -= step;
            ++hit.iterations;
            const double tolerance = 1e-13 * (1.0 + fabs(tt));
            if (!(fabs(step) > tolerance)) break;
            if (step * step * curvature <= tolerance * fabs(df)) break;
        }
        hit.t = tt;
    } else {
        hit.t = sag_intercept_closed(S, qx, qy, qz, vx, vy, vz);
    }
    double nx, ny, nz;
    sag_normal(S, qx + vx * hit.t, qy + vy * hit.t, nx, ny, nz);
    hit.nx = nx;
    hit.ny = ny;
    hit.nz = nz;
    return hit;
}

static __device__ __noinline__ Vec3 ruling_cold(const optk_surface_t& S, double px, double py, double pz, double nx,
                                                double ny, double nz) {
    Vec3 k;
    double kx, ky, kz;
    ruling_vector(S, px, py, pz, nx, ny, nz, kx, ky, kz);
    k.x = kx;
    k.y = ky;
    k.z = kz;
    return k;
}

static __device__ __noinline__ bool aperture_cold(const optk_surface_t& S, double x, double y, double z) {
    return aperture_test(S, x, y, z);
}

__device__ __forceinline__ double glass_index_inline(const optk_surface_t& S, double w) {
    // optika/materials/_materials.py:428-438
    const double w2 = w * w;
    return fsqrt(1.0 + (S.material[0] * fdiv(w2, w2 - S.material[3]) + S.material[1] * fdiv(w2, w2 - S.material[4]) +
                        S.material[2] * fdiv(w2, w2 - S.material[5])));
}
static __device__ __noinline__ double glass_index(const optk_surface_t& S, double w) { return glass_index_inline(S, w); }

static __device__ __noinline__ double surface_efficiency(const optk_surface_t& S, double w, double px, double py,
                                                         double pz, double dx, double dy, double dz, double nx,
                                                         double ny, double nz);

// ---------------------------------------------------------------------------
// one surface, FULL operator (every stage, sag normal, no sag transformation): the
// streamlined path that SequentialSystem.raytrace / propagate_rays / accumulate_rays
// take.  R rays per thread walk the surface together: every decision that depends
// only on the surface (transform kind, sag kind, rulings, material, aperture kind) is
// taken once for the R rays, whose arithmetic then interleaves (instruction-level
// parallelism for the long fp64 dependency chains).
// AbstractSurface.propagate_rays, optika/surfaces.py:123-198.
// ---------------------------------------------------------------------------
// What a thread knows about its rays between the surfaces of a walk.
// `attenuating`: some ray of the thread still carries a non-zero attenuation (set when the rays are loaded,
// cleared by every non-mirror material, which zeroes the attenuation).
// `unit`: the directions are unit vectors to rounding -- true behind every surface: either the surface
// verified it (the "straight" test below) or Snell's law renormalised them.  Only the first surface of a
// walk has to look.
// `deferred`: the walk started from finite positions (generated rays) and only its last state is reported.  The
// reference turns the intensity into NaN at the surface whose displacement is not finite (sags/_abc.py:116-120
// with attenuation 0); a position that is not finite stays so through every later surface, so looking once, at
// the end of the walk, marks the same rays (trace_body does; 3 instructions per ray instead of per surface).
struct WalkState {
    bool attenuating;
    bool unit;
    bool deferred;
};

// 0 * (finite) = 0, 0 * (inf or NaN) = NaN: adds the reference's NaN without a select
__device__ __forceinline__ void poison_intensity(Ray& r) { r.intensity = fma(0.0, r.px + r.py + r.pz, r.intensity); }

template <int R, bool EFF = false, class K = TableKinds>
__device__ __forceinline__ void surface_full(const optk_surface_t& S, Ray (&r)[R], unsigned& newton_iterations,
                                             WalkState& state) {
    const int flags = K::flags(S);
    bool& attenuating = state.attenuating;

    // 1. global -> surface-local (surfaces.py:141-142).  Inside a walk the rays come straight from the local
    // frame of the previous surface (launch_trace composes "previous local -> global -> this local" into one
    // map per transition: half the affine arithmetic of a system whose surfaces each have a frame).
    if (flags & OPTK_F_RELATIVE_IN) {
        if (flags & OPTK_F_RELATIVE_IDENTITY) {
            // consecutive surfaces in one frame (an obscuration in front of the mirror that casts it)
        } else if (flags & OPTK_F_RELATIVE_TRANSLATION) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                r[k].px += S.sag_transform.t[0];
                r[k].py += S.sag_transform.t[1];
                r[k].pz += S.sag_transform.t[2];
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                affine_forward(S.sag_transform, r[k].px, r[k].py, r[k].pz, false);
                affine_forward(S.sag_transform, r[k].dx, r[k].dy, r[k].dz, true);
            }
        }
    } else if (flags & OPTK_F_TRANSLATION_ONLY) {  // R == identity: R^T (p - t) = p - t exactly
#pragma unroll
        for (int k = 0; k < R; ++k) {
            r[k].px -= S.transform.t[0];
            r[k].py -= S.transform.t[1];
            r[k].pz -= S.transform.t[2];
        }
    } else if (flags & OPTK_F_TRANSFORM) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            affine_inverse(S.transform, r[k].px, r[k].py, r[k].pz, false);
            affine_inverse(S.transform, r[k].dx, r[k].dy, r[k].dz, true);
        }
    }

    // 2 + 3. path length to the sag and the unit normal at the hit point (surfaces.py:144-148)
    double t[R], nx[R], ny[R], nz[R];
    switch (K::sag(S)) {
        case OPTK_SAG_FLAT:
#pragma unroll
            for (int k = 0; k < R; ++k) {
                t[k] = fdiv_finite(-r[k].pz, r[k].dz);  // optika/sags/_flat.py:58
                nx[k] = 0.0; ny[k] = 0.0; nz[k] = -1.0;  // :43-47
            }
            break;
        case OPTK_SAG_SPHERICAL: {
            // optika/sags/_spherical.py:176-184, 141-146
            const double rad = S.sag[0], c = S.sag[3];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const double qx = r[k].px, qy = r[k].py, pz = r[k].pz - rad;
                const double vx = r[k].dx, vy = r[k].dy, vz = r[k].dz;
                const double up = dot3(vx, vy, vz, qx, qy, pz);
                const double disc = fma(up, up, -(norm2_3(qx, qy, pz) - rad * rad));
                t[k] = -up - signed_root(rad * vz, fsqrt(disc));
                nx[k] = c * fma(vx, t[k], qx);
                ny[k] = c * fma(vy, t[k], qy);
                nz[k] = -fsqrt(fma(-ny[k], ny[k], fma(-nx[k], nx[k], 1.0)));
            }
            break;
        }
        case OPTK_SAG_PARABOLIC: {
            // optika/sags/_parabolic.py:142-151, 56-63
            const double f = S.sag[0], ir = 0.5 * S.sag[3];  // 1 / (2 f)
#pragma unroll
            for (int k = 0; k < R; ++k) {
                t[k] = parabola_intercept(f, r[k].px, r[k].py, r[k].pz, r[k].dx, r[k].dy, r[k].dz);
                const double xr = fma(r[k].dx, t[k], r[k].px) * ir, yr = fma(r[k].dy, t[k], r[k].py) * ir;
                const double inv = frsqrt_raw(fma(xr, xr, fma(yr, yr, 1.0)));
                nx[k] = xr * inv;
                ny[k] = yr * inv;
                nz[k] = -inv;
            }
            break;
        }
        default:
            if constexpr (K::fixed) {
                // a kernel compiled for this system: the rarer kinds inline, for all rays together
                if (K::sag(S) == OPTK_SAG_TOROIDAL) {
                    toroid_intercept<R>(S, r, t, newton_iterations);
                } else {
#pragma unroll
                    for (int k = 0; k < R; ++k)
                        t[k] = sag_intercept_closed<K>(S, r[k].px, r[k].py, r[k].pz, r[k].dx, r[k].dy, r[k].dz);
                }
#pragma unroll
                for (int k = 0; k < R; ++k)
                    sag_normal<K>(S, r[k].px + r[k].dx * t[k], r[k].py + r[k].dy * t[k], nx[k], ny[k], nz[k]);
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const SagHit hit = sag_cold(S, r[k].px, r[k].py, r[k].pz, r[k].dx, r[k].dy, r[k].dz);
                    t[k] = hit.t;
                    nx[k] = hit.nx;
                    ny[k] = hit.ny;
                    nz[k] = hit.nz;
                    newton_iterations += hit.iterations;
                }
            }
            break;
    }
    {
        // optika/sags/_abc.py:116-120: intensity *= exp(-attenuation * |displacement|).
        // With attenuation == 0 the factor is exactly 1 unless the displacement is not finite
        // (exp(-0 * inf) = exp(-0 * nan) = nan in the reference).
        if (attenuating) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const double ex = r[k].dx * t[k], ey = r[k].dy * t[k], ez = r[k].dz * t[k];
                r[k].intensity = exp(-r[k].att * fsqrt(ex * ex + ey * ey + ez * ez)) * r[k].intensity;
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const double hx = fma(r[k].dx, t[k], r[k].px), hy = fma(r[k].dy, t[k], r[k].py), hz = fma(r[k].dz, t[k], r[k].pz);
            r[k].px = hx; r[k].py = hy; r[k].pz = hz;
            if (!state.deferred) poison_intensity(r[k]);
        }
    }

    // 4. rulings.incident_effective  (surfaces.py:150-154, rulings/_rulings.py:107-128, 187-204)
    if (K::ruling(S) != OPTK_RULING_NONE) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            double kx, ky, kz;
            if (K::ruling(S) == OPTK_RULING_CONSTANT) {
                // optika/rulings/_spacing.py:69-74
                const double c = S.ruling_coeff[0];
                kx = c * S.ruling_normal[0];
                ky = c * S.ruling_normal[1];
                kz = c * S.ruling_normal[2];
            } else if constexpr (K::fixed) {
                ruling_vector<K>(S, r[k].px, r[k].py, r[k].pz, nx[k], ny[k], nz[k], kx, ky, kz);
            } else {
                const Vec3 kappa = ruling_cold(S, r[k].px, r[k].py, r[k].pz, nx[k], ny[k], nz[k]);
                kx = kappa.x;
                ky = kappa.y;
                kz = kappa.z;
            }
            // a + sign(a.n) m w g / (n d), g = kappa / d, d = |kappa|  ==  a + sign(a.n) m w kappa / (n d^2)
            const double k2 = norm2_3(kx, ky, kz);
            const double an = dot3(r[k].dx, r[k].dy, r[k].dz, nx[k], ny[k], nz[k]);
            const double sg = (an == 0.0) ? an : copysign(1.0, an);  // numpy.sign
            const double f = fdiv_finite(sg * S.ruling_order * r[k].w, r[k].n * k2);
            r[k].dx = fma(f, kx, r[k].dx);
            r[k].dy = fma(f, ky, r[k].dy);
            r[k].dz = fma(f, kz, r[k].dz);
        }
    }

    // efficiency = material.efficiency(rays_1, normal) [* rulings.efficiency(rays_1, normal)] on the
    // effective direction and the incoming wavelength (surfaces.py:175-179); a call, and only for
    // measured mirrors / rulings and groove profiles
    // measured mirrors / rulings and groove profiles (kernels instantiated with EFF, trace_eff.cu:
    // the call site costs the unit-efficiency kernels 2-3 % even when it is never taken)
    if (EFF && (S.material_efficiency != OPTK_EFF_UNIT || S.ruling_profile != OPTK_PROFILE_IDEAL)) {
#pragma unroll
        for (int k = 0; k < R; ++k)
            r[k].intensity *= surface_efficiency(S, r[k].w, r[k].px, r[k].py, r[k].pz, r[k].dx, r[k].dy, r[k].dz,
                                                 nx[k], ny[k], nz[k]);
    }

    // 5-8. material: index, wavelength, Snell, attenuation  (surfaces.py:156-190)
    {
        const bool mirror = K::material(S) == OPTK_MAT_MIRROR;
        const bool glass = K::material(S) == OPTK_MAT_GLASS;
        double n2[R];
        bool same_medium = true;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (mirror) {
                n2[k] = r[k].n;  // _materials.py:135-139
            } else if (glass) {
                n2[k] = K::fixed ? glass_index_inline(S, r[k].w) : glass_index(S, r[k].w);
            } else {
                n2[k] = 1.0;  // _materials.py:95-99
            }
            same_medium = same_medium && (r[k].n == n2[k]);
        }
        // A unit, undiffracted direction that stays in its medium leaves Snell's law unchanged:
        // b = a + d n with d = -a.n + sign(a.n) sqrt(1 + (a.n)^2 - |a|^2) = O(|a|^2 - 1).  The
        // plain surfaces of a system (object, stops, baffles, the sensor) are all of this kind;
        // skipping the root there changes directions by < 1e-14 / |a.n|.  Directions that are not
        // unit to 1e-14 (user input) take the full formula, which renormalises them as the
        // reference does.
        bool straight = same_medium && !mirror && K::ruling(S) == OPTK_RULING_NONE;
        if (straight && !state.unit) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const double a2 = norm2_3(r[k].dx, r[k].dy, r[k].dz);
                straight = straight && (fabs(a2 - 1.0) <= 1e-14);
            }
        }
        if (!mirror) attenuating = false;  // every non-mirror material zeroes the attenuation below
        state.unit = true;  // verified just now, or renormalised by the full formula below
        if (straight) {
#pragma unroll
            for (int k = 0; k < R; ++k) r[k].att = 0.0;  // _materials.py:101-105, 440-444; index unchanged
        } else {
        // n1 == n2: r = 1 and 1 / r^2 = 1 exactly (no divisions, wavelength unchanged)
        double ratio[R], inv_r2[R];
#pragma unroll
        for (int k = 0; k < R; ++k) ratio[k] = inv_r2[k] = 1.0;
        if (!same_medium) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                ratio[k] = fdiv(r[k].n, n2[k]);
                inv_r2[k] = frcp(ratio[k] * ratio[k]);
                r[k].w = fdiv(r[k].w, ratio[k]);  // surfaces.py:165
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            // optika/materials/_snells_law.py:341-366
            const double a2 = norm2_3(r[k].dx, r[k].dy, r[k].dz);
            const double au = dot3(r[k].dx, r[k].dy, r[k].dz, nx[k], ny[k], nz[k]);
            const double root = fsqrt(fma(au, au, inv_r2[k] - a2));
            // d = -au + sgn (2 mirror - 1) root with sgn = -copysign(1, au)
            const double d = -au + copysign(root, mirror ? -au : au);
            r[k].dx = ratio[k] * fma(d, nx[k], r[k].dx);
            r[k].dy = ratio[k] * fma(d, ny[k], r[k].dy);
            r[k].dz = ratio[k] * fma(d, nz[k], r[k].dz);
            if (!mirror) r[k].att = 0.0;  // _materials.py:101-105, 141-145, 440-444
            r[k].n = n2[k];
        }
        }
    }

    // 9. aperture.clip_rays on the outgoing ray, local coordinates  (surfaces.py:192-193)
    if (K::aperture(S) != OPTK_APERTURE_NONE) {
        if (K::aperture(S) == OPTK_APERTURE_RECTANGULAR && !(flags & (OPTK_F_APERTURE_TRANSFORM | OPTK_F_APERTURE_ANGULAR))) {
            // optika/apertures/_apertures.py:962-963 (the common case, inlined)
            const double hx = S.aperture[0], hy = S.aperture[1];
            const bool inverted = flags & OPTK_F_APERTURE_INVERTED, active = flags & OPTK_F_APERTURE_ACTIVE;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                bool m = (-hx <= r[k].px) && (r[k].px <= hx) && (-hy <= r[k].py) && (r[k].py <= hy);
                m = (m != inverted) || !active;
                r[k].unv = r[k].unv && m;
            }
        } else if (K::aperture(S) == OPTK_APERTURE_CIRCULAR && !(flags & (OPTK_F_APERTURE_TRANSFORM | OPTK_F_APERTURE_ANGULAR))) {
            // optika/apertures/_apertures.py:309 as "x^2 + y^2 <= T" (see aperture_test), inlined
            const double threshold = S.aperture[3];
            const bool inverted = flags & OPTK_F_APERTURE_INVERTED, active = flags & OPTK_F_APERTURE_ACTIVE;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                bool m = add_rn(mul_rn(r[k].px, r[k].px), mul_rn(r[k].py, r[k].py)) <= threshold;
                m = (m != inverted) || !active;
                r[k].unv = r[k].unv && m;
            }
        } else {
            const bool angular = flags & OPTK_F_APERTURE_ANGULAR;  // dimensionless aperture: test the direction
#pragma unroll
            for (int k = 0; k < R; ++k) {
                bool m;
                if constexpr (K::fixed) {
                    m = angular ? aperture_test<K>(S, r[k].dx, r[k].dy, r[k].dz)
                                : aperture_test<K>(S, r[k].px, r[k].py, r[k].pz);
                } else {
                    m = angular ? aperture_cold(S, r[k].dx, r[k].dy, r[k].dz)
                                : aperture_cold(S, r[k].px, r[k].py, r[k].pz);
                }
                r[k].unv = r[k].unv && m;
            }
        }
    }

    // 10. local -> global  (surfaces.py:195-196)
    if (!(flags & OPTK_F_LOCAL_OUT)) {
        if (flags & OPTK_F_TRANSLATION_ONLY) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                r[k].px += S.transform.t[0];
                r[k].py += S.transform.t[1];
                r[k].pz += S.transform.t[2];
            }
        } else if (flags & OPTK_F_TRANSFORM) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                affine_forward(S.transform, r[k].px, r[k].py, r[k].pz, false);
                affine_forward(S.transform, r[k].dx, r[k].dy, r[k].dz, true);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// efficiencies (optika/surfaces.py:175-179): cold code, only systems with measured
// mirrors or groove profiles reach it (they run the generic kernels)
// ---------------------------------------------------------------------------

// numpy.interp(x, xp, fp): linear, ends clamped, NaN -> NaN
// (MeasuredMirror / MeasuredRulings, optika/materials/_materials.py:301-305, rulings/_rulings.py:309-313)
static __device__ __noinline__ double lut_interp(const double* __restrict__ xp, const double* __restrict__ fp, int n,
                                                 double x) {
    if (x != x) return x;
    if (n == 1 || x <= __ldg(xp)) return __ldg(fp);
    if (x >= __ldg(xp + n - 1)) return __ldg(fp + n - 1);
    int lo = 0, hi = n - 1;  // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= __ldg(xp + mid)) lo = mid; else hi = mid;
    }
    const double x0 = __ldg(xp + lo), x1 = __ldg(xp + lo + 1), f0 = __ldg(fp + lo), f1 = __ldg(fp + lo + 1);
    const double slope = (f1 - f0) / (x1 - x0);
    return slope * (x - x0) + f0;
}

// OPTK_EFF_TABLE2D: efficiency(wavelength, cosine of incidence) of a multilayer coating from a table that
// optk_multilayer filled once (include/optk.h, optk_surface_t.material_efficiency).  Wavelength nodes are
// explicit (the distinct wavelengths of a ray grid: exact there; or a refinement with a node on every
// kink of the optical constants): located with a guess from the mean spacing corrected against the node
// values, linear in between.  Cosine nodes are uniform: cubic Lagrange through the four around c.
static __device__ __noinline__ double table2d_lookup(const optk_surface_t& S, double w, double c) {
    const double* __restrict__ xw = S.material_lut_x;
    const double* __restrict__ table = S.material_lut_y;
    const int n_w = S.material_lut_n, n_c = (int)S.material[2];
    if (w != w || c != c) return OPTK_NAN;
    const double w_first = __ldg(xw), w_last = __ldg(xw + n_w - 1);
    bool outside = !(w >= w_first && w <= w_last);
    w = fmin(fmax(w, w_first), w_last);
    // interval i with xw[i] <= w <= xw[i + 1]
    int i = (int)((w - w_first) * ((double)(n_w - 1) / (w_last - w_first)));
    i = i < 0 ? 0 : (i > n_w - 2 ? n_w - 2 : i);
    double lo = __ldg(xw + i), hi = __ldg(xw + i + 1);
    while (w < lo) {
        --i;
        hi = lo;
        lo = __ldg(xw + i);
    }
    while (w > hi && i < n_w - 2) {
        ++i;
        lo = hi;
        hi = __ldg(xw + i + 1);
    }
    const double tw = (w - lo) / (hi - lo);
    // cosine: u in node units, stencil k - 1 .. k + 2 around the cell [k, k + 1]
    double u = (c - S.material[0]) * S.material[1];
    outside = outside || !(u >= 1.0 && u <= (double)(n_c - 2));
    u = fmin(fmax(u, 1.0), (double)(n_c - 2));
    int k = (int)u;
    k = k > n_c - 3 ? n_c - 3 : k;
    const double t = u - (double)k;
    const double tm = t - 1.0, tp = t + 1.0, t2 = t - 2.0;
    const double l0 = -t * tm * t2 * (1.0 / 6.0), l1 = tp * tm * t2 * 0.5, l2 = -tp * t * t2 * 0.5, l3 = tp * t * tm * (1.0 / 6.0);
    const double* row = table + (long long)i * n_c + (k - 1);
    const double a = l0 * __ldg(row) + l1 * __ldg(row + 1) + l2 * __ldg(row + 2) + l3 * __ldg(row + 3);
    double value = a;
    if (tw != 0.0) {  // exactly on a wavelength node (a ray grid's own wavelengths): that row alone
        row += n_c;
        const double b = l0 * __ldg(row) + l1 * __ldg(row + 1) + l2 * __ldg(row + 2) + l3 * __ldg(row + 3);
        value = fma(tw, b - a, a);
    }
    if (outside && S.ruling_lut_x && S.ruling_profile != OPTK_PROFILE_MEASURED)
        atomicAdd(reinterpret_cast<unsigned long long*>(const_cast<double*>(S.ruling_lut_x)), 1ULL);
    return value;
}

// Bessel function of the first kind of integer order (scipy.special.jv in
// SinusoidalRulings.efficiency, optika/rulings/_rulings.py:455): Miller's backward recurrence
// J_{k-1} = (2 k / x) J_k - J_{k+1} from far above max(n, x), normalised with
// 1 = J_0 + 2 (J_2 + J_4 + ...).  Absolute error ~1e-16 for |x| up to a few thousand.
static __device__ __noinline__ double bessel_jn(int n, double x) {
    if (x != x) return x;
    double sign = 1.0;
    if (n < 0) {
        n = -n;
        if (n & 1) sign = -sign;  // J_{-n} = (-1)^n J_n
    }
    if (x < 0.0) {
        x = -x;
        if (n & 1) sign = -sign;  // J_n(-x) = (-1)^n J_n(x)
    }
    if (x == 0.0) return n == 0 ? 1.0 : 0.0;
    if (!(x < 1.0e6)) return 0.0 * x;  // inf -> nan like jv; huge arguments are outside the thin-grating model
    if (x < 1e-8) {  // leading term of the series, exact to 1e-17 relative here
        double term = 1.0;
        for (int k = 1; k <= n; ++k) term *= 0.5 * x / k;
        return sign * term;
    }
    const int top = (n > (int)x ? n : (int)x);
    int m = top + 24 + (int)(8.0 * cbrt(x));
    m += m & 1;  // even start
    double jp = 0.0, j = 1e-300, sum = 0.0, result = 0.0;
    const double two_over_x = 2.0 / x;
    for (int k = m; k > 0; --k) {
        const double jm = (double)k * two_over_x * j - jp;  // J_{k-1}
        jp = j;
        j = jm;
        if (fabs(j) > 1e250) {  // rescale; every quantity scales together
            j *= 1e-250;
            jp *= 1e-250;
            sum *= 1e-250;
            result *= 1e-250;
        }
        if (((k - 1) & 1) == 0 && k - 1 > 0) sum += j;  // even orders >= 2
        if (k - 1 == n) result = j;
    }
    sum = 2.0 * sum + j;  // j is J_0 now
    return sign * result / sum;
}

// material.efficiency(rays_1, normal) * rulings.efficiency(rays_1, normal): `r` carries the
// effective direction (after incident_effective) and the wavelength before the rescale.
static __device__ __noinline__ double surface_efficiency(const optk_surface_t& S, double w_, double px_, double py_,
                                                         double pz_, double dx_, double dy_, double dz_, double nx,
                                                         double ny, double nz) {
    // (arguments by value: a reference to the ray would pin the caller's state to local memory)
    const struct { double w, px, py, pz, dx, dy, dz; } r = {w_, px_, py_, pz_, dx_, dy_, dz_};
    double eff = 1.0;
    if (S.material_efficiency == OPTK_EFF_LUT) eff = lut_interp(S.material_lut_x, S.material_lut_y, S.material_lut_n, r.w);
    if (S.material_efficiency == OPTK_EFF_TABLE2D)
        eff = table2d_lookup(S, r.w, -(r.dx * nx + r.dy * ny + r.dz * nz));  // -direction @ normal, _multilayers.py:858, 927
    const int profile = S.ruling_profile;
    if (profile == OPTK_PROFILE_IDEAL) return eff;
    if (profile == OPTK_PROFILE_MEASURED)
        return eff * lut_interp(S.ruling_lut_x, S.ruling_lut_y, S.ruling_lut_n, r.w);
    // normal_rulings = spacing_(position, normal).normalized; parallel = (normal x normal_rulings).normalized
    double gx, gy, gz;
    ruling_vector(S, r.px, r.py, r.pz, nx, ny, nz, gx, gy, gz);
    const double gi = 1.0 / sqrt(gx * gx + gy * gy + gz * gz);
    gx *= gi; gy *= gi; gz *= gi;
    double px = ny * gz - nz * gy, py = nz * gx - nx * gz, pz = nx * gy - ny * gx;
    const double pi_ = 1.0 / sqrt(px * px + py * py + pz * pz);
    px *= pi_; py *= pi_; pz *= pi_;
    // direction - direction @ parallel: the reference subtracts the SCALAR from every component
    const double dp = r.dx * px + r.dy * py + r.dz * pz;
    const double cos_theta = -((r.dx - dp) * nx + (r.dy - dp) * ny + (r.dz - dp) * nz);
    const double PI = 3.141592653589793;
    const double m = S.ruling_order;
    const bool even = fmod(m, 2.0) == 0.0;
    switch (profile) {
        case OPTK_PROFILE_SINUSOIDAL: {
            const double gamma = PI * S.ruling_depth / (r.w * cos_theta);
            return eff * bessel_jn((int)m, 2.0 * gamma);
        }
        case OPTK_PROFILE_SQUARE: {
            const double gamma = PI * (S.ruling_depth / (PI / 4)) / (r.w * cos_theta);
            double e;
            if (m == 0.0) {
                const double c = cos(PI * gamma / 2);
                e = c * c;
            } else if (even) {
                e = 0.0;
            } else {
                const double q = 2 * sin(PI * gamma / 2) / (m * PI);
                e = q * q;
            }
            return eff * e;
        }
        case OPTK_PROFILE_SAWTOOTH: {
            const double gamma = PI * (S.ruling_depth / (PI / 2)) / (r.w * cos_theta);
            const double q = sin(PI * gamma) / (PI * (gamma + m));
            return eff * q * q;
        }
        case OPTK_PROFILE_TRIANGULAR: {
            const double gamma = PI * (S.ruling_depth / (PI * PI / 8)) / (r.w * cos_theta);
            const double h = PI * gamma / 2;
            const double a = gamma / (h * h + m * m);  // "+" as written in the reference (:901)
            const double q = a * (even ? sin(PI * PI * gamma / 4) : cos(PI * PI * gamma / 4));
            return eff * q * q;
        }
        default: {  // OPTK_PROFILE_RECTANGULAR
            const double a = 2 * PI * S.ruling_duty;
            const double root = sqrt(2 * (1 - cos(a)));
            const double gamma = PI * (S.ruling_depth / (PI / (2 * root))) / (r.w * cos_theta);
            double b = sin(PI * gamma / root);
            b = b * b;
            const double e = (m == 0.0) ? 1 - ((2 * a / PI) - (a / PI) * (a / PI)) * b
                                        : (2 / ((m * PI) * (m * PI))) * (1 - cos(m * a)) * b;
            return eff * e;
        }
    }
}

// ---------------------------------------------------------------------------
// one surface, GENERIC path (partial stage masks, caller-supplied normals: the unit
// operations of the reference API): AbstractSurface.propagate_rays, optika/surfaces.py:123-198
// ---------------------------------------------------------------------------
static __device__ __noinline__ void surface_generic(const optk_surface_t& S, Ray& r, unsigned& newton_iterations,
                                                  bool normal_given, double gnx, double gny, double gnz,
                                                  double& cos_incidence) {
    const int stages = S.stages;
    const int flags = S.flags;

    // 1. global -> surface-local (surfaces.py:141-142)
    if (flags & OPTK_F_TRANSLATION_ONLY) {  // R == identity: R^T (p - t) = p - t exactly
        r.px -= S.transform.t[0];
        r.py -= S.transform.t[1];
        r.pz -= S.transform.t[2];
    } else if (flags & OPTK_F_TRANSFORM) {
        affine_inverse(S.transform, r.px, r.py, r.pz, false);
        affine_inverse(S.transform, r.dx, r.dy, r.dz, true);
    }

    // sag frame copy of the ray (sag.transformation, e.g. optika/sags/_spherical.py:153-154)
    double qx = r.px, qy = r.py, qz = r.pz;
    double vx = r.dx, vy = r.dy, vz = r.dz;
    const bool sag_t = flags & OPTK_F_SAG_TRANSFORM;
    if (sag_t) {
        affine_inverse(S.sag_transform, qx, qy, qz, false);
        affine_inverse(S.sag_transform, vx, vy, vz, true);
    }

    // 2. sag.propagate_rays: intercept (+ Beer-Lambert)  (surfaces.py:144)
    if (stages & OPTK_STAGE_INTERCEPT) {
        double t;
        if (S.sag_kind == OPTK_SAG_TOROIDAL) {
            // AbstractSag.intercept, optika/sags/_abc.py:76-107: root of
            // f(t) = (o + t u).z - sag(T^-1 (o + t u)) from t = 0.  Newton with the
            // analytic gradient, iterated to convergence (the reference's secant
            // stops at |step| < 1e-6 mm; see DESIGN.md "toroid intercept").
            const double c = S.sag[3], rr = S.sag[2];
            const double curvature = 4.0 * fmax(fabs(c), fabs(frcp(rr)));  // see sag_cold
            t = fdiv_newton(-r.pz, r.dz);
            if (!(fabs(t) < OPTK_INF)) t = 0.0;
            for (int it = 0; it < 64; ++it) {
                double z, dzdx, dzdy;
                toroid_eval(c, rr, qx + vx * t, qy + vy * t, z, dzdx, dzdy);
                const double f = (r.pz + r.dz * t) - z;
                const double df = r.dz - (dzdx * vx + dzdy * vy);
                const double step = fdiv_newton(f, df);
                t -= step;
                ++newton_iterations;
                const double tolerance = 1e-13 * (1.0 + fabs(t));
                if (!(fabs(step) > tolerance)) break;
                if (step * step * curvature <= tolerance * fabs(df)) break;
            }
        } else {
            t = sag_intercept_closed(S, qx, qy, qz, vx, vy, vz);
        }
        const double nx_ = r.px + r.dx * t, ny_ = r.py + r.dy * t, nz_ = r.pz + r.dz * t;
        if (stages & OPTK_STAGE_ATTENUATE) {
            // optika/sags/_abc.py:116-120: intensity *= exp(-attenuation * |displacement|)
            const double ex = nx_ - r.px, ey = ny_ - r.py, ez = nz_ - r.pz;
            const double len2 = ex * ex + ey * ey + ez * ez;
            if (r.att != 0.0) {
                r.intensity = exp(-r.att * fsqrt(len2)) * r.intensity;
            } else if (!(len2 <= 1.7976931348623157e308)) {
                r.intensity = NAN;  // exp(-0 * inf) = exp(-0 * nan) = nan in the reference
            }
        }
        r.px = nx_; r.py = ny_; r.pz = nz_;
        qx += vx * t; qy += vy * t; qz += vz * t;
        if (sag_t) {  // keep the sag-frame position consistent with the surface-frame one
            qx = r.px; qy = r.py; qz = r.pz;
            affine_inverse(S.sag_transform, qx, qy, qz, false);
        }
    }

    // 3. normal = sag.normal(position_1)  (surfaces.py:146-148)
    double nx, ny, nz;
    sag_normal(S, qx, qy, nx, ny, nz);
    if (normal_given) {  // unit operations with a caller-supplied normal
        nx = gnx; ny = gny; nz = gnz;
    }

    if (stages & OPTK_STAGE_SAG_OUT) {
        double z = sag_value(S, qx, qy);
        if (S.sag_kind == OPTK_SAG_FLAT && sag_t) {
            // optika/sags/_flat.py:31-41: z of the transformed (x, y, 0)
            double tx = qx, ty = qy, tz = 0.0;
            affine_forward(S.sag_transform, tx, ty, tz, false);
            z = tz;
        }
        r.pz = z;
    }
    if (stages & OPTK_STAGE_NORMAL_OUT) {
        r.dx = nx; r.dy = ny; r.dz = nz;
    }

    // 4. rulings.incident_effective  (surfaces.py:150-154, rulings/_rulings.py:107-128, 187-204)
    if ((stages & (OPTK_STAGE_RULINGS | OPTK_STAGE_KAPPA_OUT)) && S.ruling_kind != OPTK_RULING_NONE) {
        double kx, ky, kz;
        ruling_vector(S, r.px, r.py, r.pz, nx, ny, nz, kx, ky, kz);
        if (stages & OPTK_STAGE_KAPPA_OUT) {
            r.dx = kx; r.dy = ky; r.dz = kz;
        } else {
            // a + sign(a.n) m w g / (n d), g = kappa / d, d = |kappa|  ==  a + sign(a.n) m w kappa / (n d^2)
            const double k2 = kx * kx + ky * ky + kz * kz;
            const double s = sign0(r.dx * nx + r.dy * ny + r.dz * nz);
            const double f = fdiv_finite(s * S.ruling_order * r.w, r.n * k2);
            r.dx += f * kx;
            r.dy += f * ky;
            r.dz += f * kz;
        }
    }

    // efficiency = material.efficiency(rays_1, normal) [* rulings.efficiency(rays_1, normal)]
    // on the effective direction and the incoming wavelength (surfaces.py:175-179)
    const bool has_efficiency = S.material_efficiency != OPTK_EFF_UNIT || S.ruling_profile != OPTK_PROFILE_IDEAL;
    double efficiency = 1.0;
    if (has_efficiency && (stages & (OPTK_STAGE_REFRACT | OPTK_STAGE_EFFICIENCY_OUT)))
        efficiency = surface_efficiency(S, r.w, r.px, r.py, r.pz, r.dx, r.dy, r.dz, nx, ny, nz);
    if (stages & OPTK_STAGE_EFFICIENCY_OUT) r.intensity = efficiency;

    // 5-8. material: index, wavelength, Snell, attenuation  (surfaces.py:156-190)
    if (stages & OPTK_STAGE_REFRACT) {
        if (has_efficiency) r.intensity *= efficiency;
        const double n1 = r.n;
        double n2;
        const bool mirror = S.material_kind == OPTK_MAT_MIRROR || S.material_kind == OPTK_MAT_INDEX_MIRROR;
        const bool pass = S.material_kind == OPTK_MAT_PASS;  // multilayer film: index and attenuation unchanged
        if (S.material_kind == OPTK_MAT_INDEX || S.material_kind == OPTK_MAT_INDEX_MIRROR) {
            n2 = S.material[0];  // snells_law(direction, n1, n2, normal, is_mirror) unit operation
        } else if (pass) {
            n2 = n1;  // _multilayers.py:795-799
        } else if (S.material_kind == OPTK_MAT_GLASS) {
            // optika/materials/_materials.py:428-438
            const double w2 = r.w * r.w;
            n2 = fsqrt(1.0 + (S.material[0] * fdiv(w2, w2 - S.material[3]) + S.material[1] * fdiv(w2, w2 - S.material[4]) +
                              S.material[2] * fdiv(w2, w2 - S.material[5])));
        } else if (mirror) {
            n2 = n1;  // _materials.py:135-139
        } else {
            n2 = 1.0;  // _materials.py:95-99
        }
        // optika/materials/_snells_law.py:341-366
        const double a2 = r.dx * r.dx + r.dy * r.dy + r.dz * r.dz;
        const double au = r.dx * nx + r.dy * ny + r.dz * nz;
        cos_incidence = -au;  // -direction @ normal on the effective direction (_multilayers.py:858, 927)
        double ratio = 1.0, inv_r2 = 1.0;
        if (n1 != n2) {  // n1 == n2: r = 1 and 1 / r^2 = 1 exactly, skip the divisions
            ratio = fdiv(n1, n2);
            inv_r2 = frcp(ratio * ratio);
            r.w = fdiv(r.w, ratio);  // surfaces.py:165
        }
        const double sgn = -copysign(1.0, au);
        // sqrt(1/r^2 + (a.u)^2 - |a|^2).  For an undiffracted ray in an unchanged medium
        // the radicand is (a.u)^2 + e with e = 1/r^2 - |a|^2 ~ 1e-16: the root is
        // |a.u| + e / (2 |a.u|) to better than 1e-26 relative, no square root needed.
        const double au2 = au * au;
        const double e = inv_r2 - a2;
        double root;
        if (fabs(e) < 1e-13 * au2) {
            const double m = fabs(au);
            root = fma(0.5 * e, frcp(m), m);
        } else {
            root = fsqrt(au2 + e);
        }
        const double d = -au + sgn * (mirror ? 1.0 : -1.0) * root;
        r.dx = ratio * (r.dx + d * nx);
        r.dy = ratio * (r.dy + d * ny);
        r.dz = ratio * (r.dz + d * nz);
        if (!mirror && !pass) r.att = 0.0;  // _materials.py:101-105, 141-145, 440-444
        r.n = n2;
    }

    // 9. aperture.clip_rays on the outgoing ray, local coordinates  (surfaces.py:192-193)
    if ((stages & OPTK_STAGE_CLIP) && S.aperture_kind != OPTK_APERTURE_NONE) {
        bool m;
        if (flags & OPTK_F_APERTURE_ANGULAR)
            m = aperture_test(S, r.dx, r.dy, r.dz);  // dimensionless aperture: test the direction
        else
            m = aperture_test(S, r.px, r.py, r.pz);
        r.unv = r.unv && m;
    }

    // 10. local -> global  (surfaces.py:195-196)
    if (!(flags & OPTK_F_LOCAL_OUT)) {
        if (flags & OPTK_F_TRANSLATION_ONLY) {
            r.px += S.transform.t[0];
            r.py += S.transform.t[1];
            r.pz += S.transform.t[2];
        } else if (flags & OPTK_F_TRANSFORM) {
            affine_forward(S.transform, r.px, r.py, r.pz, false);
            affine_forward(S.transform, r.dx, r.dy, r.dz, true);
        }
    }
}

__device__ __forceinline__ void store_ray(const optk_rays_out_t& out, long long o, const Ray& r) {
    if (out.field[OPTK_WAVELENGTH]) out.field[OPTK_WAVELENGTH][o] = r.w;
    if (out.field[OPTK_PX]) out.field[OPTK_PX][o] = r.px;
    if (out.field[OPTK_PY]) out.field[OPTK_PY][o] = r.py;
    if (out.field[OPTK_PZ]) out.field[OPTK_PZ][o] = r.pz;
    if (out.field[OPTK_DX]) out.field[OPTK_DX][o] = r.dx;
    if (out.field[OPTK_DY]) out.field[OPTK_DY][o] = r.dy;
    if (out.field[OPTK_DZ]) out.field[OPTK_DZ][o] = r.dz;
    if (out.field[OPTK_INTENSITY]) out.field[OPTK_INTENSITY][o] = r.intensity;
    if (out.field[OPTK_ATTENUATION]) out.field[OPTK_ATTENUATION][o] = r.att;
    if (out.field[OPTK_INDEX_REFRACTION]) out.field[OPTK_INDEX_REFRACTION][o] = r.n;
    if (out.unvignetted) out.unvignetted[o] = r.unv ? 1 : 0;
}

// FULL:  every surface runs the full operator with the sag normal (surface_full, R = 2 rays
//        per thread); otherwise the generic path with stage masks / caller normals (R = 1).
// DENSE: every input is a dense array indexed by the ray index.  VEC: the dense arrays are
//        16-byte aligned, so the two rays of a thread move as one 128-bit load / store.
// ACC:   write the state after every surface.   IMAGE: bin the final rays.
template <int R, bool DENSE>
__device__ __forceinline__ void load_rays(const TraceParams& P, long long i0, long long j0, const long long* base,
                                          const bool (&valid)[R], Ray (&r)[R], bool normal_given, double& gnx,
                                          double& gny, double& gnz) {
    long long off[OPTK_NUM_FIELDS + 1];
    long long offn[3] = {0, 0, 0};
    uint32_t last_index = 0;
    bool have_offsets = false;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        if (!valid[k]) continue;
        const long long i = i0 + k;
        if (DENSE) {
            r[k].w = __ldg(P.in.field[OPTK_WAVELENGTH] + i);
            r[k].px = __ldg(P.in.field[OPTK_PX] + i);
            r[k].py = __ldg(P.in.field[OPTK_PY] + i);
            r[k].pz = __ldg(P.in.field[OPTK_PZ] + i);
            r[k].dx = __ldg(P.in.field[OPTK_DX] + i);
            r[k].dy = __ldg(P.in.field[OPTK_DY] + i);
            r[k].dz = __ldg(P.in.field[OPTK_DZ] + i);
            r[k].intensity = __ldg(P.in.field[OPTK_INTENSITY] + i);
            r[k].att = __ldg(P.in.field[OPTK_ATTENUATION] + i);
            r[k].n = __ldg(P.in.field[OPTK_INDEX_REFRACTION] + i);
            r[k].unv = P.in.unvignetted ? (__ldg(P.in.unvignetted + i) != 0) : true;
            if (normal_given) {
                gnx = __ldg(P.in.normal[0] + i);
                gny = __ldg(P.in.normal[1] + i);
                gnz = __ldg(P.in.normal[2] + i);
            }
        } else {
            // broadcast view: the CTA-level offsets (leading axes) are in `base`, the thread adds
            // the trailing axes of its own ray.  `off`, `offn` and `last_index` live across the
            // rays of the thread: the second ray is the first one's neighbour along the last axis
            // and only adds that axis' strides unless it wraps.
#ifdef OPTK_JIT_LAYOUT
            // A kernel compiled for this input layout (jit.cu): the number of axes, which of them
            // the thread resolves itself, and WHICH FIELD VARIES ALONG WHICH AXIS are compile-time
            // constants, so only the strides that are not zero cost a multiply-add (a separable
            // grid: two or three out of sixty) and an absent mask costs nothing.
            {
                constexpr int n_axes = OPTK_JIT_N_AXES, first = OPTK_JIT_FIRST, last = n_axes - 1;
                if (k > 0 && have_offsets && last_index + 1u < P.div[last].divisor) {
                    ++last_index;
#pragma unroll
                    for (int f = 0; f <= OPTK_NUM_FIELDS; ++f)
                        if (OPTK_JIT_VARIES(f, last)) off[f] += P.stride32[f][last];
                } else {
                    uint32_t rem = (uint32_t)(j0 + k + P.index_offset);
                    int o32[OPTK_NUM_FIELDS + 1];
#pragma unroll
                    for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) o32[f] = 0;
#pragma unroll
                    for (int a = n_axes - 1; a >= first; --a) {
                        uint32_t q, idx;
                        if (a == first) {
                            idx = rem;
                        } else {
                            divmod(rem, P.div[a], q, idx);
                            rem = q;
                        }
                        if (a == last) last_index = idx;
#pragma unroll
                        for (int f = 0; f <= OPTK_NUM_FIELDS; ++f)
                            if (OPTK_JIT_VARIES(f, a)) o32[f] += (int)idx * P.stride32[f][a];
                    }
#pragma unroll
                    for (int f = 0; f <= OPTK_NUM_FIELDS; ++f)
                        off[f] = (OPTK_JIT_VARIES_OUTER(f) ? base[f] : 0) + o32[f];
                    have_offsets = first < n_axes;
                }
            }
            r[k].w = __ldg(P.in.field[OPTK_WAVELENGTH] + off[OPTK_WAVELENGTH]);
            r[k].px = __ldg(P.in.field[OPTK_PX] + off[OPTK_PX]);
            r[k].py = __ldg(P.in.field[OPTK_PY] + off[OPTK_PY]);
            r[k].pz = __ldg(P.in.field[OPTK_PZ] + off[OPTK_PZ]);
            r[k].dx = __ldg(P.in.field[OPTK_DX] + off[OPTK_DX]);
            r[k].dy = __ldg(P.in.field[OPTK_DY] + off[OPTK_DY]);
            r[k].dz = __ldg(P.in.field[OPTK_DZ] + off[OPTK_DZ]);
            r[k].intensity = __ldg(P.in.field[OPTK_INTENSITY] + off[OPTK_INTENSITY]);
            r[k].att = __ldg(P.in.field[OPTK_ATTENUATION] + off[OPTK_ATTENUATION]);
            r[k].n = __ldg(P.in.field[OPTK_INDEX_REFRACTION] + off[OPTK_INDEX_REFRACTION]);
            r[k].unv = OPTK_JIT_HAS_MASK ? (__ldg(P.in.unvignetted + off[OPTK_NUM_FIELDS]) != 0) : true;
            continue;
#endif
            const int last = P.in.n_axes - 1;
            if (k > 0 && have_offsets && !normal_given && last_index + 1u < P.div[last].divisor) {
                ++last_index;
#pragma unroll
                for (int f = 0; f < OPTK_NUM_FIELDS; ++f) off[f] += P.in.stride[f][last];
                off[OPTK_NUM_FIELDS] += P.in.mask_stride[last];
            } else {
#pragma unroll
            for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) off[f] = base[f];
            offn[0] = base[OPTK_NUM_FIELDS + 1];
            offn[1] = base[OPTK_NUM_FIELDS + 2];
            offn[2] = base[OPTK_NUM_FIELDS + 3];
            uint32_t rem = (uint32_t)(j0 + k + P.index_offset);
            const int first = P.in.n_axes - P.n_inner_axes;
            if (P.offsets32 && !normal_given) {
                // small broadcast arrays: 32-bit multiply-adds, one widening add per field at the end
                int o32[OPTK_NUM_FIELDS + 1];
#pragma unroll
                for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) o32[f] = 0;
                for (int a = P.in.n_axes - 1; a >= first; --a) {
                    uint32_t q, idx;
                    if (a == first) {
                        idx = rem;
                    } else {
                        divmod(rem, P.div[a], q, idx);
                        rem = q;
                    }
                    if (a == last) last_index = idx;
#pragma unroll
                    for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) o32[f] += (int)idx * P.stride32[f][a];
                }
#pragma unroll
                for (int f = 0; f <= OPTK_NUM_FIELDS; ++f) off[f] += o32[f];
            } else
            for (int a = P.in.n_axes - 1; a >= first; --a) {
                uint32_t q, idx;
                if (a == first) {
                    idx = rem;
                } else {
                    divmod(rem, P.div[a], q, idx);
                    rem = q;
                }
                if (a == last) last_index = idx;
#pragma unroll
                for (int f = 0; f < OPTK_NUM_FIELDS; ++f) off[f] += (long long)idx * P.in.stride[f][a];
                off[OPTK_NUM_FIELDS] += (long long)idx * P.in.mask_stride[a];
                if (normal_given) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) offn[c] += (long long)idx * P.in.normal_stride[c][a];
                }
            }
            have_offsets = P.n_inner_axes > 0;
            }
            if (normal_given) {
                gnx = __ldg(P.in.normal[0] + offn[0]);
                gny = __ldg(P.in.normal[1] + offn[1]);
                gnz = __ldg(P.in.normal[2] + offn[2]);
            }
            r[k].w = __ldg(P.in.field[OPTK_WAVELENGTH] + off[OPTK_WAVELENGTH]);
            r[k].px = __ldg(P.in.field[OPTK_PX] + off[OPTK_PX]);
            r[k].py = __ldg(P.in.field[OPTK_PY] + off[OPTK_PY]);
            r[k].pz = __ldg(P.in.field[OPTK_PZ] + off[OPTK_PZ]);
            r[k].dx = __ldg(P.in.field[OPTK_DX] + off[OPTK_DX]);
            r[k].dy = __ldg(P.in.field[OPTK_DY] + off[OPTK_DY]);
            r[k].dz = __ldg(P.in.field[OPTK_DZ] + off[OPTK_DZ]);
            r[k].intensity = __ldg(P.in.field[OPTK_INTENSITY] + off[OPTK_INTENSITY]);
            r[k].att = __ldg(P.in.field[OPTK_ATTENUATION] + off[OPTK_ATTENUATION]);
            r[k].n = __ldg(P.in.field[OPTK_INDEX_REFRACTION] + off[OPTK_INDEX_REFRACTION]);
            r[k].unv = P.in.unvignetted ? (__ldg(P.in.unvignetted + off[OPTK_NUM_FIELDS]) != 0) : true;
        }
    }
}

// ---------------------------------------------------------------------------
// on-device ray grid (optk_trace_grid)
// ---------------------------------------------------------------------------

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so the draw of a ray depends only on
// (seed, index of its cell in the whole grid), never on the launch geometry.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// t = (bits + 1/2) 2^-25 of a 25-bit integer: strictly inside (0, 1)
__device__ __forceinline__ double grid_sample(const double* vertices, int i, bool jitter, uint32_t bits25) {
    const double lo = __ldg(vertices + i), hi = __ldg(vertices + i + 1);
    if (!jitter) return 0.5 * (lo + hi);
    const double t = ((double)bits25 + 0.5) * 2.98023223876953125e-08;  // 2^-25
    return fma(t, hi - lo, lo);
}

// sin and cos of the stratified sample of an ANGULAR cell from the host's packed record
// {sin v_i, cos v_i, v_{i+1} - v_i, 0} (optk_grid_t::angular_cells): the addition theorems with a
// seven-term series in the offset d = t (v_{i+1} - v_i), |d| <= 0.01 rad (checked by the caller):
// truncation < 3e-21, so the result is sin / cos of the sampled angle to rounding -- 16 fp64
// instructions instead of the ~35 of a full-range sincos, for each of the two angles of every ray.
__device__ __forceinline__ void cell_sincos(const double* __restrict__ cells, int i, bool jitter, uint32_t bits25,
                                            double& s, double& c) {
    const double2 sc = __ldg(reinterpret_cast<const double2*>(cells) + 2 * i);
    const double w = __ldg(cells + 4 * i + 2);
    const double t = jitter ? ((double)bits25 + 0.5) * 2.98023223876953125e-08 : 0.5;
    const double d = t * w, d2 = d * d;
    const double sd = fma(d * d2, fma(d2, fma(d2, -1.0 / 5040.0, 1.0 / 120.0), -1.0 / 6.0), d);   // sin d
    const double cm = d2 * fma(d2, fma(d2, -1.0 / 720.0, 1.0 / 24.0), -0.5);                      // cos d - 1
    s = fma(sc.x, cm, fma(sc.y, sd, sc.x));
    c = fma(sc.y, cm, fma(-sc.x, sd, sc.y));
}

// Curvilinear grids (field_2d / pupil_2d): bilinear sample of a cell of two 2-D vertex arrays
// [n_a + 1][n_b + 1].  Out of line: the separable case stays the short one.
static __device__ __noinline__ Vec3 grid_sample_2d(const double* vx, const double* vy, int ia, int ib, int row, bool jitter,
                                                   uint32_t bits_a, uint32_t bits_b) {
    const double ta = jitter ? ((double)bits_a + 0.5) * 2.98023223876953125e-08 : 0.5;
    const double tb = jitter ? ((double)bits_b + 0.5) * 2.98023223876953125e-08 : 0.5;
    const long long o = (long long)ia * row + ib;
    Vec3 out;
    {
        const double v00 = __ldg(vx + o), v01 = __ldg(vx + o + 1), v10 = __ldg(vx + o + row), v11 = __ldg(vx + o + row + 1);
        const double lo = fma(ta, v10 - v00, v00), hi = fma(ta, v11 - v01, v01);  // along axis a, then b
        out.x = fma(tb, hi - lo, lo);
    }
    {
        const double v00 = __ldg(vy + o), v01 = __ldg(vy + o + 1), v10 = __ldg(vy + o + row), v11 = __ldg(vy + o + row + 1);
        const double lo = fma(ta, v10 - v00, v00), hi = fma(ta, v11 - v01, v01);
        out.y = fma(tb, hi - lo, lo);
    }
    out.z = 0.0;
    return out;
}

// Launch-wide facts about the grid: run-time tests in the table-driven kernels, compile-time constants
// in the run-time specialised ones (jit.cu defines OPTK_JIT_GRID_FLAGS for the variant).
#define OPTK_GRID_AT_INFINITY 0
#define OPTK_GRID_PACKED 1
#define OPTK_GRID_JITTER 2
#define OPTK_GRID_FRAME 3
#define OPTK_GRID_WEIGHT_SCENE 4
#define OPTK_GRID_WEIGHT_PUPIL 5
#ifdef OPTK_JIT_GRID_FLAGS
#define OPTK_GRID_FLAG(bit, runtime) ((((OPTK_JIT_GRID_FLAGS) >> (bit)) & 1) != 0)
#else
#define OPTK_GRID_FLAG(bit, runtime) (runtime)
#endif

// Chromatic axes (optk_grid_t::chromatic): the vertices of axis a form one row per WAVELENGTH vertex,
// [n_0 + 1][n_a + 1] -- field / pupil extents that come from a stop solution per wavelength
// (optika/systems/_sequential.py:748-789).  The sample is bilinear in (wavelength, a), along the
// wavelength first.  Out of line, in the curvilinear instantiations only.
static __device__ __noinline__ double grid_sample_chromatic(const double* v, int iw, int i, int row, double tw, bool jitter,
                                                            uint32_t bits25) {
    const double t = jitter ? ((double)bits25 + 0.5) * 2.98023223876953125e-08 : 0.5;
    const long long o = (long long)iw * row + i;
    const double v00 = __ldg(v + o), v01 = __ldg(v + o + 1), v10 = __ldg(v + o + row), v11 = __ldg(v + o + row + 1);
    if (!jitter) return 0.5 * (0.5 * (v00 + v10) + 0.5 * (v01 + v11));
    const double lo = fma(tw, v10 - v00, v00), hi = fma(tw, v11 - v01, v01);
    return fma(t, hi - lo, lo);
}

// SequentialSystem._rayfunction_from_vertices + _calc_rayfunction_input
// (optika/systems/_sequential.py:1055-1086, 791-828) for the R consecutive rays of a thread.
// `j0` is the C-order index of the first ray in the sub-box of this launch (< 2^31).
template <int R, bool CURVILINEAR>
__device__ __forceinline__ void generate_rays(const TraceParams& P, uint32_t j0, const bool (&valid)[R],
                                              Ray (&r)[R]) {
    const optk_grid_t& G = P.grid;
    int idx[5];
    {
        uint32_t rem = j0;
#pragma unroll
        for (int a = 4; a >= 1; --a) {
            uint32_t q, i;
            divmod(rem, P.div[a], q, i);
            idx[a] = (int)i;
            rem = q;
        }
        idx[0] = (int)rem;
    }
    const bool jitter = OPTK_GRID_FLAG(OPTK_GRID_JITTER, G.jitter != 0);
    const bool at_infinity = OPTK_GRID_FLAG(OPTK_GRID_AT_INFINITY, G.at_infinity != 0);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        if (k > 0) {
            // next ray of the thread: increment with carry
            ++idx[4];
#pragma unroll
            for (int a = 4; a >= 1; --a) {
                if (idx[a] == (int)P.div[a].divisor) {
                    idx[a] = 0;
                    ++idx[a - 1];
                }
            }
        }
        if (!valid[k]) continue;
        int g[5];
        unsigned long long cell = 0;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            g[a] = idx[a] + G.begin[a];
            cell += (unsigned long long)(unsigned)g[a] * P.cell_stride[a];
        }
        uint32_t x[4] = {0u, 0u, 0u, 0u};
        if (jitter)
            philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), 0u, 0u, (uint32_t)G.seed,
                          (uint32_t)(G.seed >> 32), x);
        // five 25-bit integers from the 128 bits: the top 25 bits of each word, and the
        // low 7 bits of the four words side by side (include/optk.h)
        const uint32_t low = (x[0] & 127u) | ((x[1] & 127u) << 7) | ((x[2] & 127u) << 14) | ((x[3] & 15u) << 21);
        const double w = grid_sample(G.vertices[0], g[0], jitter, x[0] >> 7);
        double fx = 0.0, fy = 0.0, px = 0.0, py = 0.0;
        // the angular pair (field for an object at infinity, else pupil) enters only through its sines
        // and cosines: with the host's packed cell records they come from the addition theorems
        const bool packed = !CURVILINEAR && OPTK_GRID_FLAG(OPTK_GRID_PACKED, G.angular_cells[0] != nullptr);
        const bool field_angular = at_infinity;
        double sx, cx, sy, cy;
        const double tw = jitter ? ((double)(x[0] >> 7) + 0.5) * 2.98023223876953125e-08 : 0.5;
        if (packed && field_angular) {
            cell_sincos(G.angular_cells[0], g[1], jitter, x[1] >> 7, sx, cx);
            cell_sincos(G.angular_cells[1], g[2], jitter, x[2] >> 7, sy, cy);
        } else if (CURVILINEAR && (G.chromatic & 6)) {
            fx = (G.chromatic & 2) ? grid_sample_chromatic(G.vertices[1], g[0], g[1], G.n[1] + 1, tw, jitter, x[1] >> 7)
                                   : grid_sample(G.vertices[1], g[1], jitter, x[1] >> 7);
            fy = (G.chromatic & 4) ? grid_sample_chromatic(G.vertices[2], g[0], g[2], G.n[2] + 1, tw, jitter, x[2] >> 7)
                                   : grid_sample(G.vertices[2], g[2], jitter, x[2] >> 7);
        } else if (CURVILINEAR && G.field_2d) {
            const Vec3 f = grid_sample_2d(G.vertices[1], G.vertices[2], g[1], g[2], G.n[2] + 1, jitter, x[1] >> 7, x[2] >> 7);
            fx = f.x;
            fy = f.y;
        } else {
            fx = grid_sample(G.vertices[1], g[1], jitter, x[1] >> 7);
            fy = grid_sample(G.vertices[2], g[2], jitter, x[2] >> 7);
        }
        if (packed && !field_angular) {
            cell_sincos(G.angular_cells[0], g[3], jitter, x[3] >> 7, sx, cx);
            cell_sincos(G.angular_cells[1], g[4], jitter, low, sy, cy);
        } else if (CURVILINEAR && (G.chromatic & 24)) {
            px = (G.chromatic & 8) ? grid_sample_chromatic(G.vertices[3], g[0], g[3], G.n[3] + 1, tw, jitter, x[3] >> 7)
                                   : grid_sample(G.vertices[3], g[3], jitter, x[3] >> 7);
            py = (G.chromatic & 16) ? grid_sample_chromatic(G.vertices[4], g[0], g[4], G.n[4] + 1, tw, jitter, low)
                                    : grid_sample(G.vertices[4], g[4], jitter, low);
        } else if (CURVILINEAR && G.pupil_2d) {
            const Vec3 p = grid_sample_2d(G.vertices[3], G.vertices[4], g[3], g[4], G.n[4] + 1, jitter, x[3] >> 7, low);
            px = p.x;
            py = p.y;
        } else {
            px = grid_sample(G.vertices[3], g[3], jitter, x[3] >> 7);
            py = grid_sample(G.vertices[4], g[4], jitter, low);
        }
        // position / angles by the location of the object (:797-802)
        if (!packed) {
            const double ax = at_infinity ? fx : px, ay = at_infinity ? fy : py;
            fsincos(ax, &sx, &cx);
            fsincos(ay, &sy, &cy);
        }
        r[k].w = w;
        r[k].px = at_infinity ? px : fx;
        r[k].py = at_infinity ? py : fy;
        r[k].pz = 0.0;
        r[k].dx = -cy * sx;  // optika.direction, optika/_util.py:64-73
        r[k].dy = -sy;
        r[k].dz = cy * cx;
        double weight = 1.0;
        if (OPTK_GRID_FLAG(OPTK_GRID_WEIGHT_SCENE, G.weight_scene != nullptr))
            weight = __ldg(G.weight_scene + (g[0] * G.n[1] + g[1]) * (long long)G.n[2] + g[2]);
        if (OPTK_GRID_FLAG(OPTK_GRID_WEIGHT_PUPIL, G.weight_pupil != nullptr)) {
            long long at = g[3] * (long long)G.n[4] + g[4];
            if (CURVILINEAR && G.weight_pupil_chromatic) at += (long long)g[0] * G.n[3] * G.n[4];
            weight *= __ldg(G.weight_pupil + at);
        }
        r[k].intensity = weight;
        r[k].att = 0.0;
        r[k].n = 1.0;
        r[k].unv = true;
        if (OPTK_GRID_FLAG(OPTK_GRID_FRAME, G.has_frame != 0)) {
            affine_forward(G.frame, r[k].px, r[k].py, r[k].pz, false);
            affine_forward(G.frame, r[k].dx, r[k].dy, r[k].dz, true);
        }
    }
}

// 128-bit access to two consecutive rays of one field
__device__ __forceinline__ void load_pair(const double* p, long long i, double& a, double& b) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p + i));
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void store_pair(double* p, long long i, double a, double b) {
    if (p) *reinterpret_cast<double2*>(p + i) = make_double2(a, b);
}

__device__ __forceinline__ void store_rays_vec(const optk_rays_out_t& out, long long o, const Ray (&r)[2]) {
    store_pair(out.field[OPTK_WAVELENGTH], o, r[0].w, r[1].w);
    store_pair(out.field[OPTK_PX], o, r[0].px, r[1].px);
    store_pair(out.field[OPTK_PY], o, r[0].py, r[1].py);
    store_pair(out.field[OPTK_PZ], o, r[0].pz, r[1].pz);
    store_pair(out.field[OPTK_DX], o, r[0].dx, r[1].dx);
    store_pair(out.field[OPTK_DY], o, r[0].dy, r[1].dy);
    store_pair(out.field[OPTK_DZ], o, r[0].dz, r[1].dz);
    store_pair(out.field[OPTK_INTENSITY], o, r[0].intensity, r[1].intensity);
    store_pair(out.field[OPTK_ATTENUATION], o, r[0].att, r[1].att);
    store_pair(out.field[OPTK_INDEX_REFRACTION], o, r[0].n, r[1].n);
    if (out.unvignetted)
        *reinterpret_cast<uchar2*>(out.unvignetted + o) = make_uchar2(r[0].unv ? 1 : 0, r[1].unv ? 1 : 0);
}

#ifdef OPTK_JIT_WALK
__device__ __forceinline__ void optk_jit_walk(const TraceParams& P, Ray (&r)[2], unsigned& newton_iterations,
                                              WalkState& state);
#endif

// SPEC: 0 the surface list is walked from the table; 1 a run-time compiled walk (R = 2, no ACC)
template <int R, bool FULL, bool DENSE, bool VEC, bool ACC, bool IMAGE, int GRID, bool EFF, int SPEC = 0>
__device__ __forceinline__ void trace_body(const TraceParams& P) {
    uint32_t block = blockIdx.x;
    if (IMAGE && !DENSE && P.cta_rows) {
        block = (block & 511u) * (uint32_t)P.cta_rows + (block >> 9);  // strided visiting order, params.cuh
        if (block >= (uint32_t)P.cta_count) return;                     // padding of the last row
    }
    __shared__ ImageGuess guess_shared;
    // one barrier for all per-CTA set-up: the image guess is filled by the first thread of the
    // second warp while the first warp computes the outer offsets.  With the caller's range hint
    // (optk_image_t::has_range; compiled in for the run-time specialised kernels) the guess is a set of
    // constant-bank operands: nothing to fill, and launches without per-CTA offsets need no barrier.
#if defined(OPTK_JIT_IMAGE_FLAGS) && ((OPTK_JIT_IMAGE_FLAGS) & 1)
    constexpr bool guess_const = true;
#else
    const bool guess_const = IMAGE && P.image.has_range;
#endif
    if (IMAGE && !guess_const && threadIdx.x == 32) image_guess_fill(P.image, &guess_shared);

    // Dense input: ray index = thread index.  Broadcast input: the CTA owns one index of the
    // leading ("outer") axes and a tile of the trailing ("inner") axes; its outer offsets are
    // computed once by OPTK_NUM_FIELDS + 4 threads and shared.
    __shared__ long long base[OPTK_NUM_FIELDS + 4];
    long long i0, j0 = 0;
    long long limit = P.n_rays;
    if (DENSE || GRID) {
        i0 = ((long long)block * blockDim.x + threadIdx.x) * R;
    } else {
        uint32_t outer32, tile32;
        divmod(block, P.div_tiles, outer32, tile32);  // block / tiles_per_outer without a 64-bit division
        const long long outer = outer32, tile = tile32;
        j0 = (tile * blockDim.x + threadIdx.x) * R;
        i0 = outer * P.inner_size + j0;
        limit = (outer + 1) * P.inner_size;
        if (threadIdx.x < OPTK_NUM_FIELDS + 4) {
            const int f = threadIdx.x;
            long long o = 0;
            uint32_t rem = (uint32_t)outer;
            for (int a = P.in.n_axes - P.n_inner_axes - 1; a >= 0; --a) {
                uint32_t q, idx;
                if (a == 0) {
                    idx = rem;
                } else {
                    divmod(rem, P.div[a], q, idx);
                    rem = q;
                }
                const long long st = f < OPTK_NUM_FIELDS ? P.in.stride[f][a]
                                     : (f == OPTK_NUM_FIELDS ? P.in.mask_stride[a]
                                                             : P.in.normal_stride[f - OPTK_NUM_FIELDS - 1][a]);
                o += (long long)idx * st;
            }
            base[f] = o;
        }
    }
    if ((IMAGE && !guess_const) || !(DENSE || GRID)) __syncthreads();
    bool valid[R];
    Ray r[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        valid[k] = i0 + k < limit;
        // dummy ray for idle lanes; its NaN wavelength keeps it out of every table statistic (table2d_lookup)
        r[k] = Ray{OPTK_NAN, 0.0, 0.0, -1.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1.0, false};
    }
    unsigned newton_iterations = 0;
    double cos_incidence = 0.0;  // generic path only: captured at the last traced surface
    const bool normal_given = !FULL && P.in.normal[0] != nullptr;
    double gnx = 0.0, gny = 0.0, gnz = -1.0;

    // Memory-level parallelism: each CTA computes for ~10k cycles between its loads and
    // its stores, so too few bytes are in flight to cover HBM latency.  Every thread asks L2
    // for the lines that the CTA scheduled one "wave" later will load, so DRAM streams while
    // the SMs compute and those loads become L2 hits.
    if (DENSE && P.prefetch_distance > 0) {
        const long long ip = i0 + P.prefetch_distance;
        if (ip < P.n_rays) {
#pragma unroll
            for (int f = 0; f < OPTK_NUM_FIELDS; ++f)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.in.field[f] + ip));
        }
    }

    // the whole thread is "vector" when both of its rays exist (only the last thread may not be)
    const bool pair = R == 2 && VEC && valid[R - 1];
    if (pair) {
        load_pair(P.in.field[OPTK_WAVELENGTH], i0, r[0].w, r[R - 1].w);
        load_pair(P.in.field[OPTK_PX], i0, r[0].px, r[R - 1].px);
        load_pair(P.in.field[OPTK_PY], i0, r[0].py, r[R - 1].py);
        load_pair(P.in.field[OPTK_PZ], i0, r[0].pz, r[R - 1].pz);
        load_pair(P.in.field[OPTK_DX], i0, r[0].dx, r[R - 1].dx);
        load_pair(P.in.field[OPTK_DY], i0, r[0].dy, r[R - 1].dy);
        load_pair(P.in.field[OPTK_DZ], i0, r[0].dz, r[R - 1].dz);
        load_pair(P.in.field[OPTK_INTENSITY], i0, r[0].intensity, r[R - 1].intensity);
        load_pair(P.in.field[OPTK_ATTENUATION], i0, r[0].att, r[R - 1].att);
        load_pair(P.in.field[OPTK_INDEX_REFRACTION], i0, r[0].n, r[R - 1].n);
        if (P.in.unvignetted) {
            const uchar2 m = *reinterpret_cast<const uchar2*>(P.in.unvignetted + i0);
            r[0].unv = m.x != 0;
            r[R - 1].unv = m.y != 0;
        } else {
            r[0].unv = r[R - 1].unv = true;
        }
    } else if (GRID) {
        generate_rays<R, GRID == 2>(P, (uint32_t)i0, valid, r);
    } else {
        load_rays<R, DENSE>(P, i0, j0, base, valid, r, normal_given, gnx, gny, gnz);
    }

    // generated rays start with zero attenuation: the Beer-Lambert branch is dead code there
    // (generated directions are (-cos b sin a, -sin b, cos b cos a): unit to 4e-16 by construction)
    WalkState state = {false, GRID != 0, GRID != 0 && !ACC && FULL};
    if (FULL && !GRID) {
#pragma unroll
        for (int k = 0; k < R; ++k) state.attenuating = state.attenuating || (r[k].att != 0.0);
    }

    // No `if (valid)` around the walk: threads past the end trace a harmless dummy ray, so the
    // surface loop stays warp-convergent and its per-surface decisions and parameter loads
    // can use the uniform datapath (only the stores are predicated).
    {
        if constexpr (SPEC == 1) {
            // a kernel compiled at run time for one system (jit.cu): the walk is a straight sequence of
            // surface_full<2, EFF, FixedKinds<...>>(P.surf[k], ...) with compile-time k
#ifdef OPTK_JIT_WALK
            optk_jit_walk(P, r, newton_iterations, state);
#endif
        } else
        for (int s = 0; s < P.n_surf; ++s) {
            if (FULL)
                surface_full<R, EFF>(P.surf[s], r, newton_iterations, state);
            else
                surface_generic(P.surf[s], r[0], newton_iterations, normal_given, gnx, gny, gnz, cos_incidence);
            if (ACC && P.has_out) {
                const long long o = (long long)s * P.accumulate_stride + i0;
                bool done = false;
                if constexpr (R == 2 && VEC) {
                    if (pair) {
                        store_rays_vec(P.out, o, r);
                        done = true;
                    }
                }
                if (!done) {
#pragma unroll
                    for (int k = 0; k < R; ++k)
                        if (valid[k]) store_ray(P.out, o + k, r[k]);
                }
            }
        }
        if (state.deferred) {
#pragma unroll
            for (int k = 0; k < R; ++k) poison_intensity(r[k]);
        }
        if (!ACC && P.has_out) {  // one uniform test instead of eleven null checks per ray
            bool done = false;
            if constexpr (R == 2 && VEC) {
                if (pair) {
                    store_rays_vec(P.out, i0, r);
                    done = true;
                }
            }
            if (!done) {
#pragma unroll
                for (int k = 0; k < R; ++k)
                    if (valid[k]) store_ray(P.out, i0 + k, r[k]);
            }
        }
    }

    if (!FULL && P.out.cos_incidence && valid[0]) P.out.cos_incidence[i0] = cos_incidence;

    if (IMAGE) {
        // AbstractImagingSensor.collect on the final rays in sensor-local coordinates
        // (optika/systems/_sequential.py:983-986, optika/sensors/_sensors.py:125-161);
        // IdealSensorMaterial: cos = -direction . (0, 0, -1) = d_z.
        int bin[R];
        double w_flux[R], w_real[R];
        unsigned count[R];
        const ImageGuess guess = guess_const ? image_guess_const(P.image) : guess_shared;
        // Group accumulators exist only in kernels compiled at run time with OPTK_JIT_GROUPS (jit.cu):
        // as a run-time branch they cost the detector path 4 % (measured), and launch_trace refuses
        // group launches that no such kernel serves.
#ifdef OPTK_JIT_GROUPS
        constexpr bool groups = true;
#else
        constexpr bool groups = false;
#endif
        if constexpr (groups) {
            // Not a detector: one accumulator per group of consecutive rays (the pupil of a field
            // point), optk_image_t::group_size -- sums of intensity, x, y and the number of the
            // unvignetted rays (SequentialSystem.distortion / vignetting / area_effective).  The sum
            // of y goes first, in a pass of its own, so that the detector path below carries no third
            // weight (two more live doubles cost the fused image kernels 3-4 %, measured).
            int gb[R];
            double wy[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                double x = r[k].px, y = r[k].py, z = r[k].pz;
                if (P.has_frame) affine_inverse(P.frame, x, y, z, false);
                uint32_t q, rem;
                divmod((uint32_t)(i0 + k), P.image.div_group, q, rem);
                gb[k] = (valid[k] && r[k].unv) ? (int)q : -1;
                wy[k] = y;
            }
            if (R == 2 && gb[0] == gb[R - 1] && gb[0] >= 0) {
                wy[0] += wy[R - 1];
                gb[R - 1] = -1;
            }
            ImageDev sum_y = {};
            sum_y.flux = P.image.moment_imag;
#pragma unroll
            for (int k = 0; k < R; ++k) image_add(sum_y, gb[k], wy[k], 0.0, 0.0, 0u);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            double x = r[k].px, y = r[k].py, z = r[k].pz, cx = r[k].dx, cy = r[k].dy, cz = r[k].dz;
            if (P.has_frame) {
                affine_inverse(P.frame, x, y, z, false);
                affine_inverse(P.frame, cx, cy, cz, true);
            }
            if constexpr (groups) {
                uint32_t q, rem;
                divmod((uint32_t)(i0 + k), P.image.div_group, q, rem);
                bin[k] = (valid[k] && r[k].unv) ? (int)q : -1;
            } else {
                bin[k] = image_bin_index(P.image, guess, valid[k], r[k].w, x, y, r[k].unv);
            }
            w_flux[k] = r[k].intensity;
            w_real[k] = groups ? x : r[k].intensity * cz;
            count[k] = 1u;
        }
        // the two rays of a thread are pupil neighbours: usually the same pixel, merged here
        if (R == 2 && bin[0] == bin[R - 1] && bin[0] >= 0) {
            w_flux[0] += w_flux[R - 1];
            w_real[0] += w_real[R - 1];
            count[0] += count[R - 1];
            bin[R - 1] = -1;
        }
        // (with groups the imaginary-moment plane holds the sum of y added above: leave it alone)
        if constexpr (groups) {
            ImageDev planes = P.image;
            planes.moment_imag = nullptr;
#pragma unroll
            for (int k = 0; k < R; ++k) image_add(planes, bin[k], w_flux[k], w_real[k], 0.0, count[k]);
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) image_add(P.image, bin[k], w_flux[k], w_real[k], 0.0, count[k]);
        }
    }

    if (P.stats) {
        const unsigned full = 0xffffffffu;
        unsigned n_unv = 0, n_val = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            n_unv += __popc(__ballot_sync(full, valid[k] && r[k].unv));
            n_val += __popc(__ballot_sync(full, valid[k]));
        }
        unsigned n_it = __reduce_add_sync(full, newton_iterations);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&P.stats->n_rays, (unsigned long long)n_val);
            atomicAdd(&P.stats->n_unvignetted, (unsigned long long)n_unv);
            if (n_it) atomicAdd(&P.stats->n_newton_iterations, (unsigned long long)n_it);
        }
    }
}

// One kernel per (FULL, DENSE, VEC, ACC, IMAGE): uniform decisions are made once on the host
// instead of per ray per surface.  The streamlined kernels carry two rays per thread in
// <= 80 registers (3 CTAs of 256 threads per SM: the measured best).
// GRID: 0 rays from memory, 1 on-device separable vertex grid, 2 curvilinear (2-D vertex arrays)
template <int MINB, int R, bool FULL, bool DENSE, bool VEC, bool ACC, bool IMAGE, int GRID = 0, bool EFF = false, int SPEC = 0>
__global__ void __launch_bounds__(256, MINB) trace_kernel(const __grid_constant__ TraceParams P) {
    trace_body<R, FULL, DENSE, VEC, ACC, IMAGE, GRID, EFF, SPEC>(P);
}

typedef void (*trace_kernel_t)(const TraceParams);

// defined in trace_full.cu / trace_generic.cu (separate translation units: parallel compilation)
trace_kernel_t select_full_kernel(bool dense, bool vec, bool acc, bool image);
trace_kernel_t select_generic_kernel(bool dense, bool acc, bool image);
trace_kernel_t select_grid_kernel(bool full, bool acc, bool image, bool curvilinear);
trace_kernel_t select_heavy_kernel(bool grid, bool acc, bool image);  // trace_heavy.cu
trace_kernel_t select_efficiency_kernel(bool grid, bool dense, bool acc, bool image);  // trace_eff.cu

}  // namespace optk
