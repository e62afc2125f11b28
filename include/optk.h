/*
 * optk.h -- C ABI of liboptk, the B200-native (sm_100a) engine for optika's
 * sequential-raytrace hot path.
 *
 * The reference (sun-data/optika) is pure Python and has no FFI; the seam this
 * library replaces is the set of Python calls listed beside each entry point
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add at each of them.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, POD structs.  No torch / CUDA types.
 *   - `stream` is a CUDA stream handle (cudaStream_t / CUstream) passed as
 *     void*; NULL is the legacy default stream.  Calls with device pointers are
 *     asynchronous and stream ordered; calls taking host pointers return after
 *     the results are in the host buffers.
 *   - all lengths are millimetres, angles radians, attenuation 1/mm.
 *   - every function returns 0 on success or a negative optk_status_t;
 *     optk_last_error() returns a thread-local message.  There is NO CPU
 *     fallback: without a CUDA device every compute entry point fails with
 *     OPTK_ERR_CUDA.
 *   - ownership: inputs are never written; outputs are caller-allocated.
 */
#ifndef OPTK_H
#define OPTK_H

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define OPTK_ABI_VERSION 9

#if defined(__GNUC__)
#define OPTK_API __attribute__((visibility("default")))
#else
#define OPTK_API
#endif

#define OPTK_MAX_SURFACES 22 /* surfaces per launch; longer systems are chained   */
#define OPTK_MAX_VERTICES 32 /* polygon aperture vertices                          */
#define OPTK_MAX_COEFF 8     /* terms of a Polynomial1dRulingSpacing               */
#define OPTK_MAX_AXES 8      /* named axes of a ray grid                           */
#define OPTK_NUM_FIELDS 10   /* fp64 fields of a ray                               */

typedef enum optk_status {
    OPTK_OK = 0,
    OPTK_ERR_INVALID = -1,     /* bad argument                     -> ValueError          */
    OPTK_ERR_UNSUPPORTED = -2, /* unsupported surface kind         -> NotImplementedError */
    OPTK_ERR_CUDA = -3,        /* CUDA runtime error / no device   -> RuntimeError        */
    OPTK_ERR_NOMEM = -4
} optk_status_t;

/* Order of the fp64 ray fields (optika/rays/_ray_vectors.py:256-278). */
typedef enum optk_field {
    OPTK_WAVELENGTH = 0,
    OPTK_PX = 1, OPTK_PY = 2, OPTK_PZ = 3,
    OPTK_DX = 4, OPTK_DY = 5, OPTK_DZ = 6,
    OPTK_INTENSITY = 7,
    OPTK_ATTENUATION = 8,
    OPTK_INDEX_REFRACTION = 9
} optk_field_t;

typedef enum optk_sag_kind {
    OPTK_SAG_FLAT = 0,        /* optika/sags/_flat.py:26-64          */
    OPTK_SAG_SPHERICAL = 1,   /* optika/sags/_spherical.py:100-191   */
    OPTK_SAG_CYLINDRICAL = 2, /* optika/sags/_cylindrical.py:75-160  */
    OPTK_SAG_CONIC = 3,       /* optika/sags/_conic.py:31-170        */
    OPTK_SAG_PARABOLIC = 4,   /* optika/sags/_parabolic.py:40-158    */
    OPTK_SAG_TOROIDAL = 5     /* optika/sags/_toroidal.py:38-88 + _abc.py:76-107 */
} optk_sag_kind_t;

typedef enum optk_material_kind {
    OPTK_MAT_VACUUM = 0, /* optika/materials/_materials.py:82-116  */
    OPTK_MAT_MIRROR = 1, /* optika/materials/_materials.py:120-175 */
    OPTK_MAT_GLASS = 2,  /* optika/materials/_materials.py:428-455 */
    /* unit operation optika.materials.snells_law (optika/materials/_snells_law.py:41-47):
     * the new index is given explicitly in material[0]; transmit / reflect */
    OPTK_MAT_INDEX = 3,
    OPTK_MAT_INDEX_MIRROR = 4,
    /* index and attenuation pass through, no reflection: AbstractMultilayerFilm
     * (optika/materials/_multilayers.py:795-805, 835-837) */
    OPTK_MAT_PASS = 5
} optk_material_kind_t;

typedef enum optk_efficiency_kind {
    OPTK_EFF_UNIT = 0,
    OPTK_EFF_LUT = 1,
    /* efficiency(wavelength, cos(incidence)) from a table in DEVICE memory: multilayer coatings
     * (MultilayerMirror / MultilayerFilm .efficiency, optika/materials/_multilayers.py:839-866, 908-935)
     * tabulated once with optk_multilayer instead of one stack evaluation per ray; see
     * optk_surface_t.material_efficiency */
    OPTK_EFF_TABLE2D = 2
} optk_efficiency_kind_t;

typedef enum optk_ruling_profile {
    OPTK_PROFILE_IDEAL = 0,       /* Rulings.efficiency = 1, optika/rulings/_rulings.py:246-251 */
    OPTK_PROFILE_SINUSOIDAL = 1,
    OPTK_PROFILE_SQUARE = 2,
    OPTK_PROFILE_SAWTOOTH = 3,
    OPTK_PROFILE_TRIANGULAR = 4,
    OPTK_PROFILE_RECTANGULAR = 5,
    OPTK_PROFILE_MEASURED = 6
} optk_ruling_profile_t;

typedef enum optk_ruling_kind {
    OPTK_RULING_NONE = 0,
    OPTK_RULING_CONSTANT = 1,    /* optika/rulings/_spacing.py:45-74   */
    OPTK_RULING_POLYNOMIAL = 2,  /* optika/rulings/_spacing.py:78-128  */
    OPTK_RULING_HOLOGRAPHIC = 3  /* optika/rulings/_spacing.py:295-328 */
} optk_ruling_kind_t;

typedef enum optk_aperture_kind {
    OPTK_APERTURE_NONE = 0,
    OPTK_APERTURE_CIRCULAR = 1,    /* optika/apertures/_apertures.py:292-314 */
    OPTK_APERTURE_RECTANGULAR = 2, /* optika/apertures/_apertures.py:941-968 */
    OPTK_APERTURE_POLYGON = 3,     /* optika/apertures/_apertures.py:739-778 */
    OPTK_APERTURE_ELLIPTICAL = 4,  /* optika/apertures/_apertures.py:640-663 */
    OPTK_APERTURE_SECTOR = 5       /* optika/apertures/_apertures.py:438-480 */
} optk_aperture_kind_t;

/* optk_surface_t.flags */
#define OPTK_F_TRANSFORM 0x001          /* surface.transformation is not None      */
#define OPTK_F_SAG_TRANSFORM 0x002      /* sag.transformation is not None          */
#define OPTK_F_APERTURE_TRANSFORM 0x004 /* aperture.transformation is not None     */
#define OPTK_F_RULING_TRANSFORM 0x008   /* Polynomial1dRulingSpacing.transformation */
#define OPTK_F_APERTURE_INVERTED 0x010
#define OPTK_F_APERTURE_ACTIVE 0x020
#define OPTK_F_APERTURE_ANGULAR 0x040   /* clip on direction instead of position   */
#define OPTK_F_HOLO_DIVERGING_1 0x080
#define OPTK_F_HOLO_DIVERGING_2 0x100
#define OPTK_F_TRANSLATION_ONLY 0x400 /* set by the library: `transform` has R == identity */
/* Set by the library (optk_system_create): the polygon's vertices are in strictly convex position, listed
 * counter-clockwise (_CONVEX) or clockwise (_CONVEX | _CLOCKWISE).  aperture[0] then holds B = max |vertex
 * coordinate| and aperture[1] the width 1e-12 B^2 of the band around the edge lines inside which the kernels
 * decide with the exact even-odd arithmetic; everywhere else eight half-plane tests give the same answer. */
#define OPTK_F_APERTURE_CONVEX 0x4000
#define OPTK_F_APERTURE_CLOCKWISE 0x8000
/* Set by the library per launch (never by callers): the rays arrive in the LOCAL frame of the previous surface
 * of the walk and `sag_transform` (unused by the full operator) holds the relative map previous-local ->
 * this-local, applied forwards; _TRANSLATION: its rotation is the identity; _IDENTITY: both surfaces share
 * one frame, nothing to apply. */
#define OPTK_F_RELATIVE_IN 0x800
#define OPTK_F_RELATIVE_TRANSLATION 0x1000
#define OPTK_F_RELATIVE_IDENTITY 0x2000
#define OPTK_F_LOCAL_OUT 0x200 /* skip the final local -> global step: the rays leave the
                                  surface in its LOCAL frame (sensor.transformation.inverse,
                                  optika/systems/_sequential.py:983-986)                    */

/*
 * optk_surface_t.stages: which steps of AbstractSurface.propagate_rays
 * (optika/surfaces.py:123-198) run.  OPTK_STAGE_ALL is the full operator; the
 * partial masks implement the reference's unit operations (sag.intercept,
 * aperture.clip_rays, ...) on the same kernel.
 */
#define OPTK_STAGE_INTERCEPT 0x01 /* sag.intercept           (sags/_abc.py:76-107)        */
#define OPTK_STAGE_ATTENUATE 0x02 /* Beer-Lambert            (sags/_abc.py:109-122)       */
#define OPTK_STAGE_RULINGS 0x04   /* incident_effective      (rulings/_rulings.py:170-204) */
#define OPTK_STAGE_REFRACT 0x08   /* index, wavelength, Snell (surfaces.py:156-190)        */
#define OPTK_STAGE_CLIP 0x10      /* aperture.clip_rays      (apertures/_apertures.py:82-102) */
#define OPTK_STAGE_NORMAL_OUT 0x20 /* write sag.normal(position) into the direction fields */
#define OPTK_STAGE_SAG_OUT 0x40    /* write sag(position) into the z position field        */
#define OPTK_STAGE_KAPPA_OUT 0x80  /* write spacing_(position, normal) into direction      */
#define OPTK_STAGE_EFFICIENCY_OUT 0x100 /* write material.efficiency * rulings.efficiency into intensity */
#define OPTK_STAGE_ALL 0x1f

/* x -> R x + t, row-major R.  The inverse is evaluated as R^T (x - t). */
typedef struct optk_affine {
    double r[9];
    double t[3];
} optk_affine_t;

/*
 * One optical surface, lowered from optika.surfaces.Surface
 * (optika/surfaces.py:282-395) for ONE configuration: every parameter a scalar.
 */
typedef struct optk_surface {
    int32_t sag_kind;
    int32_t material_kind;
    int32_t ruling_kind;
    int32_t aperture_kind;
    int32_t flags;
    int32_t stages;
    int32_t n_vertices;
    int32_t n_coeff;

    optk_affine_t transform;          /* surface-local -> global                        */
    optk_affine_t sag_transform;      /* sag-local -> surface-local                     */
    optk_affine_t aperture_transform; /* aperture-local -> surface-local                */
    optk_affine_t ruling_transform;   /* applied FORWARDS to the position (spacing.py:118-119) */

    /* sag: [0] radius (spherical/cylindrical/conic/toroidal minor) or focal length
     * (parabolic); [1] conic constant; [2] radius of rotation (toroidal);
     * [3] reserved: optk_system_create stores 1 / [0] there.                      */
    double sag[4];

    /* Glass: Sellmeier b1 b2 b3 c1 c2 c3 (c in mm^2). */
    double material[6];

    double ruling_order;        /* diffraction order m                               */
    double ruling_normal[3];    /* unit vector normal to the ruling planes           */
    double ruling_coeff[OPTK_MAX_COEFF]; /* constant: [0] = spacing; polynomial: c_k  */
    int32_t ruling_power[OPTK_MAX_COEFF];
    double holo_x1[3];
    double holo_x2[3];
    double holo_wavelength;

    /* aperture: circular [0]=radius; rectangular [0..1]=half widths; elliptical
     * [0..1]=radii; sector [0]=radius [1]=angle_start [2]=angle_stop.              */
    double aperture[4];
    double vertices_x[OPTK_MAX_VERTICES];
    double vertices_y[OPTK_MAX_VERTICES];

    /* Efficiencies multiplied into the intensity next to Snell's law
     * (optika/surfaces.py:175-179), evaluated on the ray AFTER rulings.incident_effective
     * and BEFORE the wavelength rescale.  0 = unit efficiency (Vacuum / Mirror / Glass,
     * ideal Rulings).
     * material_efficiency OPTK_EFF_LUT: MeasuredMirror (optika/materials/_materials.py:279-305),
     *   numpy.interp of the ray wavelength in (lut_x, lut_y), ends clamped.
     * ruling_profile: the thin-grating groove efficiencies of optika/rulings/_rulings.py
     *   (Magnusson & Gaylord 1978, Table 1), with the reference's
     *   cos(theta) = -(a - (a . p)) . n  (p = normalized(n x g), the scalar a . p subtracted
     *   from every component, :446-449) and gamma = pi ruling_depth / (wavelength cos(theta)):
     *   SINUSOIDAL  J_m(2 gamma), unsquared as in the reference (:404-457);
     *   SQUARE (:548-614), SAWTOOTH (:705-758), TRIANGULAR (:849-911), RECTANGULAR with
     *   ruling_duty = ratio_duty (:1008-1073); `ruling_depth` is the physical depth in mm, the
     *   per-profile amplitude normalisation is applied by the kernel;
     *   MEASURED: MeasuredRulings (:287-313), numpy.interp like the mirror.
     * LUT arrays are HOST pointers when the table is given to optk_system_create, which
     * copies them to the current device; they must be ascending in x.
     * material_efficiency OPTK_EFF_TABLE2D (materials OPTK_MAT_MIRROR / OPTK_MAT_PASS only, whose
     *   material[] is otherwise unused): the efficiency is a function of the ray's wavelength w and of
     *   c = -(a . n), the cosine of incidence on the effective direction -- what the reference passes to
     *   multilayer_efficiency (_multilayers.py:858, 927) for a ray in vacuum.  All pointers are DEVICE
     *   pointers owned by the caller (nothing is copied):
     *     material_lut_x  wavelength nodes, strictly ascending, material_lut_n >= 2 of them (they may be
     *                     non-uniform: the distinct wavelengths of a ray grid, or a refinement that has
     *                     a node on every kink of the optical constants);
     *     material_lut_y  values [material_lut_n][n_c], n_c = (int) material[2] >= 4 uniform cosine
     *                     nodes c_k = material[0] + k / material[1]  (material[1] = 1 / spacing);
     *     ruling_lut_x    when the surface has no measured ruling profile: a uint64 counter in device
     *                     memory (or NULL) that the kernels increment for every ray outside the table.
     *   Lookup: linear between the two wavelength nodes around w (exact AT a node), cubic Lagrange
     *   through the four cosine nodes around c; valid for material_lut_x[0] <= w <= last node and
     *   c_1 <= c <= c_{n_c - 2}.  A ray outside takes the nearest edge value and is counted; NaN
     *   inputs give NaN.                                                                     */
    int32_t material_efficiency;
    int32_t ruling_profile;
    int32_t material_lut_n;
    int32_t ruling_lut_n;
    double ruling_depth;
    double ruling_duty;
    const double* material_lut_x;
    const double* material_lut_y;
    const double* ruling_lut_x;
    const double* ruling_lut_y;
} optk_surface_t;

typedef struct optk_system optk_system_t;

/*
 * Rays as structure-of-arrays.  `field[f]` may be a broadcast view: element
 * (i_0, ..., i_{n_axes-1}) of field f lives at
 * field[f][sum_a i_a * stride[f][a]]  (strides in ELEMENTS, 0 = broadcast), and
 * the mask likewise with mask_stride.  This is how a separable ray grid
 * (wavelength x field x pupil, optika/systems/_sequential.py:791-828) is passed
 * without being materialised.  Rays are enumerated in C order of `dims`.
 */
typedef struct optk_rays_in {
    int32_t n_axes;
    int64_t dims[OPTK_MAX_AXES];
    const double* field[OPTK_NUM_FIELDS];
    int64_t stride[OPTK_NUM_FIELDS][OPTK_MAX_AXES];
    const uint8_t* unvignetted; /* NULL = all true */
    int64_t mask_stride[OPTK_MAX_AXES];
    /* Optional caller-supplied surface normal (NULL = use sag.normal): the `normal`
     * argument of rulings.incident_effective (optika/rulings/_rulings.py:170-204),
     * AbstractRulingSpacing.__call__ (optika/rulings/_spacing.py:27-41) and
     * snells_law (optika/materials/_snells_law.py:41-47).  Applies to every surface. */
    const double* normal[3];
    int64_t normal_stride[3][OPTK_MAX_AXES];
} optk_rays_in_t;

/* Dense outputs, prod(dims) elements each (times the number of traced surfaces
 * when accumulating).  A NULL field pointer skips that output. */
typedef struct optk_rays_out {
    double* field[OPTK_NUM_FIELDS];
    uint8_t* unvignetted;
    /* Optional, one value per ray: -a . n at the LAST traced surface, a = the incident
     * direction after rulings.incident_effective, n = the surface normal: the `direction`
     * argument MultilayerMirror.efficiency / MultilayerFilm.efficiency hand to
     * multilayer_efficiency (optika/materials/_multilayers.py:858, 927).  NULL = not wanted. */
    double* cos_incidence;
} optk_rays_out_t;

/* Detector binning target (optika/sensors/_sensors.py:139-161). Planes are
 * [n_wavelength][n_x][n_y], C order, caller-zeroed; contributions are ADDED.
 * n_wavelength * n_x * n_y < 2^31 (split the wavelength axis beyond that). */
typedef struct optk_image {
    int32_t n_wavelength, n_x, n_y;
    const double* edges_wavelength; /* n_wavelength + 1, device or host like the rays */
    const double* edges_x;          /* n_x + 1 */
    const double* edges_y;          /* n_y + 1 */
    double* flux;                   /* sum of intensity            (may be NULL) */
    double* moment_real;            /* sum of intensity * Re cos   (may be NULL) */
    double* moment_imag;            /* sum of intensity * Im cos   (may be NULL) */
    unsigned long long* counts;     /* number of rays              (may be NULL) */
    /* Optional: the first and last edge of every axis, as known to the caller
     * (wavelength, x, y).  With has_range != 0 the kernels do not have to fetch them from
     * device memory at the start of every CTA.  Must equal the array values exactly. */
    int32_t has_range;
    /* Optional promise about the pixel edges (needs has_range): bit 0 -- edges_x, bit 1 -- edges_y are
     * UNIFORM: every edge lies within 1e-9 of a bin width of first + i (last - first) / n, which is what
     * numpy.linspace produces (optika/sensors/_sensors.py:141-149).  The bin of a sample is then
     * floor((v - first) n / (last - first)) and the edge arrays are consulted only for samples within
     * 1e-6 of a bin width of an edge -- identical results, no loads on the common path.  Without the
     * promise the exact search runs for every sample. */
    int32_t uniform_edges;
    double range[6];
    /* 0: a detector image as described above.  > 0 (optk_trace / optk_trace_grid only): not a
     * detector but one accumulator per GROUP of `group_size` consecutive rays of the launch (C order
     * of the ray axes; the engine keeps the pupil axes innermost, so a group is the pupil of one
     * field point) -- the reductions over the pupil of SURVEY.md section 8f-4 fused into the trace,
     * no ray is written: flux[g] = sum of intensity, moment_real[g] = sum of x, moment_imag[g] =
     * sum of y (final frame of the trace), counts[g] = number, all over the UNVIGNETTED rays of
     * group g.  n_x = number of groups, n_wavelength = n_y = 1; the edge arrays are not used.
     * Served by the run-time compiled kernels only (full surface operator, no accumulate; a launch
     * they cannot serve returns OPTK_ERR_UNSUPPORTED: reduce the traced rays with
     * optk_reduce_groups instead), so that the detector path carries no code for it. */
    int64_t group_size;
} optk_image_t;

/* Counters returned by the trace (device-side reductions, optional). */
typedef struct optk_trace_stats {
    unsigned long long n_rays;
    unsigned long long n_unvignetted;
    unsigned long long n_newton_iterations;
    unsigned long long n_binned;
} optk_trace_stats_t;

/* ---- library --------------------------------------------------------------- */
OPTK_API int optk_abi_version(void);
OPTK_API const char* optk_last_error(void);
/* Number of visible CUDA devices, or a negative status. */
OPTK_API int optk_device_count(void);

/* ---- system handle ---------------------------------------------------------
 * Replaces the Python list `SequentialSystem.surfaces_all`
 * (optika/systems/_sequential.py:93-108) as consumed by
 * optika.propagators.propagate_rays / accumulate_rays (optika/propagators.py:19-73).
 * `table` is [n_config][n_surface], copied.  Unsupported kinds are rejected here.
 */
OPTK_API int optk_system_create(const optk_surface_t* table, int32_t n_surface, int32_t n_config,
                       optk_system_t** out);
OPTK_API int optk_system_destroy(optk_system_t* sys);
OPTK_API int optk_system_size(const optk_system_t* sys, int32_t* n_surface, int32_t* n_config);
/* One surface of one configuration AS THE LIBRARY KEEPS IT: the caller's values plus what optk_system_create
 * derives once per surface -- sag[3] = 1 / sag[0], aperture[3] = the squared-radius threshold of circular and sector
 * apertures, OPTK_F_TRANSLATION_ONLY, OPTK_F_APERTURE_CONVEX / _CLOCKWISE with aperture[0..1] for convex polygons.
 * Efficiency-table pointers are the device copies.  Needs no GPU (inspection, tests). */
OPTK_API int optk_system_surface(const optk_system_t* sys, int32_t config, int32_t index, optk_surface_t* out);

/* ---- fused sequential trace (kernel 1) --------------------------------------
 * Replaces optika.propagators.propagate_rays (optika/propagators.py:19-41) and,
 * with accumulate != 0, accumulate_rays (:44-73): each ray walks surfaces
 * surf_begin, surf_begin + surf_step, ... (surf_count of them) of configuration
 * `config`, running AbstractSurface.propagate_rays (optika/surfaces.py:123-198)
 * at each.  surf_step = -1 gives the backwards trace of the stop solver
 * (optika/systems/_sequential.py:640-654).
 * With accumulate, the state after the k-th traced surface is written at element
 * offset k * accumulate_stride of every output array.
 * `image` (may be NULL) additionally bins the FINAL rays, transformed by the
 * inverse of `image_frame` (NULL = already local; this is the sensor's
 * transformation, optika/systems/_sequential.py:983-986), as
 * AbstractImagingSensor.collect does (optika/sensors/_sensors.py:92-171).
 * `out` may be NULL when only the image is wanted.
 * All pointers are DEVICE pointers.
 */
OPTK_API int optk_trace(const optk_system_t* sys, int32_t config,
               const optk_rays_in_t* in, const optk_rays_out_t* out,
               int32_t surf_begin, int32_t surf_count, int32_t surf_step,
               int32_t accumulate, int64_t accumulate_stride,
               const optk_image_t* image, const optk_affine_t* image_frame,
               optk_trace_stats_t* stats_device, void* stream);

/* Same operation with HOST pointers everywhere (rays, edges, image planes).
 * The library streams the rays through the device in slabs of `slab_rays`
 * (0 = default), overlapping H2D, kernel and D2H on three streams; `pinned`
 * says the host buffers are page-locked.  Returns when outputs are complete. */
OPTK_API int optk_trace_host(const optk_system_t* sys, int32_t config,
                    const optk_rays_in_t* in, const optk_rays_out_t* out,
                    int32_t surf_begin, int32_t surf_count, int32_t surf_step,
                    int32_t accumulate, int64_t accumulate_stride,
                    const optk_image_t* image, const optk_affine_t* image_frame,
                    optk_trace_stats_t* stats_host, int64_t slab_rays, int32_t pinned);

/* ---- on-device ray grid (kernel 1 with a generator in front) ------------------
 * Replaces SequentialSystem._rayfunction_from_vertices + _calc_rayfunction_input
 * (optika/systems/_sequential.py:1002-1086, 791-828) for a separable grid given on
 * cell VERTICES: grid.cell_centers(axis=(wavelength, field, pupil), random=True)
 * (:1066-1069) becomes a stratified sample drawn inside the kernel, the flux
 * radiance * cell_area (:1071-1077, optika/vectors/_vectors_object.py:42-133)
 * becomes weight_scene * weight_pupil, and optika.direction (optika/_util.py:41-73)
 * is evaluated per ray.  No input ray is ever read from memory.
 *
 * Axes (fixed order, C order of the outputs, pupil_y fastest):
 *   0 wavelength [mm], 1 field_x, 2 field_y, 3 pupil_x, 4 pupil_y.
 * at_infinity = 1 (object at infinity, :797-799): field = angles [rad], pupil =
 * position [mm]: position = (pupil_x, pupil_y, 0), direction = direction(field).
 * at_infinity = 0 (:800-802): field = position [mm], pupil = angles [rad].
 * Every ray starts with attenuation 0, index_refraction 1, unvignetted = true and
 * intensity = weight_scene[i0][i1][i2] * weight_pupil[i3][i4] (NULL = 1).
 *
 * Sample of ray (i0..i4), indices in the WHOLE grid n[]:
 *   v_a = lo_a + t_a * (hi_a - lo_a),  lo_a = vertices[a][i_a], hi_a = vertices[a][i_a + 1];
 *   jitter = 0: t_a = 1/2 exactly as (lo + hi) / 2;
 *   jitter = 1: t_a = (b_a + 1/2) * 2^-25 with five 25-bit integers b_a cut from the
 *   four words x_0..x_3 = Philox4x32-10(counter = (cell_lo, cell_hi, 0, 0), key =
 *   (seed_lo, seed_hi)): b_a = x_a >> 7 for a = 0..3 and b_4 = (x_0 & 127) |
 *   (x_1 & 127) << 7 | (x_2 & 127) << 14 | (x_3 & 15) << 21; cell = C-order index of
 *   the ray in the whole grid.  The stream therefore does not depend on how the grid is
 *   split into sub-boxes (begin/count), launches or GPUs.  n[0] * n[1] < 2^31.
 * `frame`, when has_frame, maps the generated rays (object-local) to the coordinates
 * the first traced surface expects: object.transformation followed by the inverse of
 * the system transformation (:823-826, :908-909). */
typedef struct optk_grid {
    int32_t n[5];     /* cells of the whole grid along each axis                      */
    int32_t begin[5]; /* first cell of the sub-box this call traces                   */
    int32_t count[5]; /* cells of the sub-box; outputs are dense in C order over it   */
    int32_t at_infinity;
    int32_t jitter;
    int32_t has_frame;
    uint64_t seed;
    const double* vertices[5];  /* device; n[a] + 1 values each (see field_2d / pupil_2d) */
    const double* weight_scene; /* device; [n0][n1][n2] or NULL                       */
    const double* weight_pupil; /* device; [n3][n4] or NULL                           */
    optk_affine_t frame;
    /* Curvilinear (e.g. polar) field / pupil grids: with field_2d the x and y vertex arrays
     * vertices[1] and vertices[2] are both 2-D, [n1 + 1][n2 + 1], and the sample of cell
     * (i1, i2) is the bilinear combination of its four corner vertices with weights
     * (t_1, t_2) (cell_centers applied along one axis after the other); likewise pupil_2d
     * for vertices[3], vertices[4] with [n3 + 1][n4 + 1] and (t_3, t_4). */
    int32_t field_2d;
    int32_t pupil_2d;
    /* Optional, separable grids only: for the two ANGULAR axes (field x, y for an object at infinity,
     * else pupil x, y) one packed record per CELL i, {sin v_i, cos v_i, v_{i+1} - v_i, 0} (device,
     * 32-byte records, 16-byte aligned), every cell at most 0.01 rad wide.  The kernel then forms the
     * sine and cosine of the sampled angle v_i + t (v_{i+1} - v_i) by the addition theorems and a short
     * series in the offset (truncation < 3e-21) instead of a full-range sincos per ray and angle; the
     * vertex arrays of those two axes are not read.  Both NULL: sincos of the sampled angle. */
    const double* angular_cells[2];
    /* Chromatic axes: bit a (a = 1 .. 4) set -- vertices[a] is a 2-D array [n[0] + 1][n[a] + 1], one row of
     * vertices per WAVELENGTH vertex (field / pupil extents from a stop solution per wavelength,
     * optika/systems/_sequential.py:748-789); the sample of cell (i_0, i_a) is bilinear in (t_0, t_a), along
     * the wavelength first.  Not combinable with field_2d / pupil_2d for the same pair, nor with
     * angular_cells.  weight_pupil_chromatic != 0: weight_pupil is [n0][n3][n4] (cell areas per wavelength
     * cell). */
    int32_t chromatic;
    int32_t weight_pupil_chromatic;
} optk_grid_t;

/* As optk_trace, with the rays of `grid` as input.  surf_count = 0 returns the
 * generated rays themselves. */
OPTK_API int optk_trace_grid(const optk_system_t* sys, int32_t config, const optk_grid_t* grid,
                    const optk_rays_out_t* out,
                    int32_t surf_begin, int32_t surf_count, int32_t surf_step,
                    int32_t accumulate, int64_t accumulate_stride,
                    const optk_image_t* image, const optk_affine_t* image_frame,
                    optk_trace_stats_t* stats_device, void* stream);

/* ---- detector binning (kernel 2) --------------------------------------------
 * Replaces the three na.histogram calls of AbstractImagingSensor.collect
 * (optika/sensors/_sensors.py:139-161) for rays already in sensor-local
 * coordinates: sample (wavelength, x, y), weight intensity * unvignetted
 * [* cos_real / cos_imag].  cos_* may be NULL (IdealSensorMaterial: cos = d_z,
 * optika/sensors/materials/_materials.py:1603-1616 is then taken from `dz`).
 * Device pointers; contributions are added to the planes of `image`.
 */
OPTK_API int optk_bin(int64_t n_rays, const double* wavelength, const double* x, const double* y,
             const double* dz, const double* intensity, const uint8_t* unvignetted,
             const optk_image_t* image, void* stream);

/* ---- multilayer transfer-matrix efficiency (kernel 3) -----------------------
 * Replaces optika.materials.multilayer_efficiency
 * (optika/materials/_multilayers.py:240-532).  The evaluation grid has n_axes
 * named axes of sizes dims[]; every input is a broadcast view with element
 * strides per axis.  Layers are listed top (ambient side) to bottom; the LAST
 * layer is the substrate (its thickness is ignored, _multilayers.py:190-193).
 * Segments express PeriodicLayerSequence (optika/materials/_layers.py:611-645):
 * layers [first, first + count) repeated `repeat` times.  Optical constants
 * n_j(lambda) are interpolated by the caller (optika/chemicals/_chemicals.py:101-144).
 * Outputs are dense, prod(dims) elements.  Device pointers.
 */
#define OPTK_ML_MAX_AXES 4
#define OPTK_ML_MAX_LAYERS 256

typedef struct optk_ml_layer {
    const double* n_re;
    const double* n_im;
    int64_t n_stride[OPTK_ML_MAX_AXES];
    const double* thickness; /* NULL = 0 */
    int64_t thickness_stride[OPTK_ML_MAX_AXES];
    const double* width; /* interface profile width, NULL = no profile */
    int64_t width_stride[OPTK_ML_MAX_AXES];
    int32_t profile_kind; /* 0 none, 1 erf, 2 exponential, 3 linear, 4 sinusoidal
                             (optika/materials/profiles.py:221-222, 320-325, 424-431, 532-541) */
    int32_t reserved;
} optk_ml_layer_t;

typedef struct optk_ml_segment {
    int32_t first;
    int32_t count;
    int32_t repeat;
    int32_t reserved;
} optk_ml_segment_t;

typedef struct optk_ml_input {
    int32_t n_axes;
    int64_t dims[OPTK_ML_MAX_AXES];
    const double* wavelength;
    int64_t wavelength_stride[OPTK_ML_MAX_AXES];
    const double* direction_re; /* cosine of the incidence angle in the ambient medium */
    const double* direction_im; /* NULL = real */
    int64_t direction_stride[OPTK_ML_MAX_AXES];
    const double* n_re; /* ambient index of refraction */
    const double* n_im; /* NULL = real */
    int64_t n_stride[OPTK_ML_MAX_AXES];
} optk_ml_input_t;

OPTK_API int optk_multilayer(const optk_ml_input_t* input,
                    int32_t n_layers, const optk_ml_layer_t* layers,
                    int32_t n_segments, const optk_ml_segment_t* segments,
                    double* reflectivity_s, double* reflectivity_p,
                    double* transmissivity_s, double* transmissivity_p,
                    void* stream);

/* ---- per-ray multilayer efficiency (rows a35) -----------------------------------
 * MultilayerMirror.efficiency / MultilayerFilm.efficiency (optika/materials/
 * _multilayers.py:839-866, 908-935) evaluate multilayer_efficiency for every ray.  The
 * host chains: optk_trace up to the coated surface with `cos_incidence` captured ->
 * optk_interp of every layer's optical constants at the ray wavelengths
 * (Chemical.n, optika/chemicals/_chemicals.py:136-142) -> optk_multilayer on the dense
 * per-ray arrays -> optk_apply_efficiency -> optk_trace of the remaining surfaces. */

/* numpy.interp(x, xp, fp) for n device values x; fp complex as (re, im), im may be NULL
 * (then out_im may be NULL); xp ascending, m >= 1 entries; ends clamped. */
OPTK_API int optk_interp(int64_t n, const double* x, int32_t m, const double* xp, const double* fp_re,
                const double* fp_im, double* out_re, double* out_im, void* stream);

/* intensity[i] *= (e_s[i] + e_p[i]) / 2   (PolarizationVectorArray.average) */
OPTK_API int optk_apply_efficiency(int64_t n, double* intensity, const double* e_s, const double* e_p, void* stream);

/* ---- reductions over the pupil -----------------------------
 * SequentialSystem.distortion / vignetting / area_effective reduce the traced rays over the pupil
 * axes for every (configuration, wavelength, field point): unvignetted.any(axis_pupil),
 * mean(position.xy, axis_pupil, where=unvignetted | ~where), unvignetted.mean(axis_pupil),
 * intensity.sum(axis_pupil, where=unvignetted) (optika/systems/_sequential.py:1266-1285,
 * 1351-1368, 1501-1506).  optk_reduce_groups computes the sums those are made of from dense
 * device arrays of n_groups * n_inner rays, group g = rays [g n_inner, (g + 1) n_inner):
 *   count[g]         number of unvignetted rays            (the caller zeroes every output first)
 *   sum_intensity[g] sum of intensity over them            (intensity NULL = 1)
 *   sum_x[g], sum_y[g]          sums of x, y over them
 *   sum_x_all[g], sum_y_all[g]  sums of x, y over ALL rays of the group (the reference's
 *                               fallback where no ray of a field point survives); may be NULL
 * unvignetted NULL = every ray counts.  Integer counts are exact, fp64 sums are independent of
 * the order up to rounding. */
OPTK_API int optk_reduce_groups(int64_t n_groups, int64_t n_inner, const double* x, const double* y,
                       const double* intensity, const uint8_t* unvignetted, double* sum_intensity, double* sum_x,
                       double* sum_y, uint64_t* count, double* sum_x_all, double* sum_y_all, void* stream);

/* ---- stop solver -------------------------------------------
 * SequentialSystem._calc_rayfunction_stops_only (optika/systems/_sequential.py:396-623) finds
 * the rays that connect a grid of points on one stop surface with a grid of points on the other
 * one by a 2-D Newton iteration with a finite-difference Jacobian (na.optimize.root_newton /
 * na.jacobian, :586-606) around _ray_error (:363-394).  optk_solve_stops runs that iteration on
 * the device, one thread per unknown ray, around the same surface walk optk_trace uses.
 *
 * The rays start in the GLOBAL frame on surface `surf_first` (they are not propagated through
 * it) and are traced through surf_first + 1 ... surf_last.  `variable` says which vector of the
 * ray is solved for -- its global x, y components are the unknowns, its z follows from them
 * (direction: sqrt(1 - x^2 - y^2); position: the sag of surf_first at (x, y), which must not
 * carry a sag transformation) -- the other vector is `fixed`.  `target` says which vector of the
 * traced ray, in the local frame of surf_last, must equal (target_x, target_y).
 * Per-ray device arrays of length n; x, y hold the initial guess on entry and the solution on
 * exit, z receives the third component.  A ray stops iterating once both residual components
 * are <= max_abs_error; *n_unconverged (device, caller-zeroed) counts rays that did not get
 * there within max_iterations (the reference raises "Max iterations exceeded"). */
#define OPTK_STOP_DIRECTION 0
#define OPTK_STOP_POSITION 1
typedef struct optk_stop_problem_t {
    int32_t variable;       /* OPTK_STOP_DIRECTION / OPTK_STOP_POSITION */
    int32_t target;         /* OPTK_STOP_DIRECTION / OPTK_STOP_POSITION */
    int32_t surf_first;
    int32_t surf_last;      /* > surf_first, at most OPTK_MAX_SURFACES - 1 surfaces apart */
    int32_t max_iterations; /* the reference's default: 100 */
    int32_t reserved;
    double step;            /* forward-difference step of the Jacobian (:586-591) */
    double max_abs_error;   /* :561-568 */
} optk_stop_problem_t;

OPTK_API int optk_solve_stops(const optk_system_t* sys, int32_t config, const optk_stop_problem_t* problem, int64_t n,
                     const double* wavelength, const double* fixed_x, const double* fixed_y, const double* fixed_z,
                     const double* target_x, const double* target_y, double* x, double* y, double* z,
                     uint32_t* n_unconverged, void* stream);

/* ---- run-time specialisation ---------------------------------------------------
 * Long launches of the streamlined kernels (>= 2^25 rays -- >= 2^20 when the kernel is already in
 * the disk cache -- full operator, no accumulate) are
 * served by a kernel compiled with NVRTC for exactly the traced surface list (every kind and
 * flag a compile-time constant, the walk unrolled; ~1.5 s once per system shape and kernel
 * variant, cached for the life of the process and, as a cubin, on disk across processes).  mode: -1 automatic (default; also the
 * environment variable OPTK_JIT=-1), 0 never, 1 for every eligible launch.  Results equal
 * those of the table-driven kernels to rounding (the same expressions; the compiler may contract
 * different multiply-add pairs into FMAs), masks and NaN / inf patterns exactly.  If libnvrtc / libcuda cannot be loaded or the
 * compilation fails, the table-driven kernels run (a message goes to stderr). */
OPTK_API int optk_jit_mode(int32_t mode);
/* Number of specialised kernels made available so far in this process (compiled, or loaded
 * from the disk cache $OPTK_JIT_CACHE, default ~/.cache/optika_b200/jit; empty string = no cache). */
OPTK_API int64_t optk_jit_compiled(void);

/* ---- measurement helpers ----------------------------------------------------
 * FP64 DFMA peak micro-benchmark (the roofline denominator that
 * MEASURED_PEAKS.json does not record): returns achieved FLOP/s. */
OPTK_API int optk_measure_fp64_peak(double* flops_per_second, void* stream);
/* Bandwidth (GB/s, 162 B per ray) of the trace kernel's access pattern with no arithmetic:
 * ten fp64 arrays + a byte mask in, the same out.  Allocates 162 * n_rays bytes. */
OPTK_API int optk_measure_soa_copy(int64_t n_rays, double* gbytes_per_second, void* stream);
/* The kernels' own fp64 division / reciprocal / square-root sequences (csrc/common.cuh: MUFU seed, one
 * third-order step or one Newton step + residual correction, no branches), element by element on device
 * arrays, so that their error can be measured against correctly rounded results (tests/test_gpu_math.py:
 * <= 1 ulp; zero / inf / NaN as the IEEE operators).  op: 0 a / b, 1 1 / a, 2 sqrt(a), 3 1 / sqrt(a),
 * 4 / 5 the variants of 1 / a and 1 / sqrt(a) without the repair of a = 0 and a = inf, 6 a / b for a divisor
 * that is never infinite (<= 1.5 ulp), 7 a / b to 2^-38 for self-correcting iterations (b is read for op 0, 6, 7). */
OPTK_API int optk_debug_math(int32_t op, int64_t n, const double* a, const double* b, double* out, void* stream);

/* ---- detector physics after binning (kernel 4) ----------------------------------
 * Replaces the numba kernel optika/sensors/materials/_ramanathan_2020/_ramanathan_2020.py:762-876
 * (_electrons_measured_numba, called by electrons_measured :468-690): a Monte-Carlo model of a
 * back-illuminated CCD.  For every photon absorbed in a pixel: the number of electron-hole pairs
 * (below 50 eV from the tabulated pair-number distribution `cmf` / `n_values`, above from a rounded
 * normal of mean energy / energy_pair_inf and variance fano_inf * mean), the absorption depth
 * (exponential of rate `absorption`, truncated to the substrate), the charge-collection efficiency at
 * that depth (cce_backsurface at the back surface rising linearly to 1 across thickness_implant) applied
 * as a binomial thinning, and for every surviving electron a Gaussian lateral step of standard deviation
 * z_ff sqrt(1 - z / z_ff) (z_ff = thickness_substrate - thickness_depletion; none beyond z_ff) from the
 * photon's uniformly random position in the pixel, rounded to whole pixels; electrons leaving the grid
 * wrap around (wrap != 0) or are lost.
 * One record per image plane (everything but the photon counts is per plane: the wavelength axis);
 * lengths in mm, energies in eV.  `cmf` (cumulative pair-number distribution) and `n_values` are DEVICE
 * arrays of n_pmf entries (ignored above 50 eV).
 * Random numbers: Philox4x32-10, key = seed, counter = {pixel index in the whole [n_plane][n_x][n_y]
 * array (64 bit), photon index, draw}: draw 0 = pair number (words 0-1 -> a 53-bit uniform; above 50 eV
 * words 2-3 -> the second uniform of a Box-Muller pair, first normal used), draw 1 = depth (words 0-1),
 * position in the pixel (word 2, word 3: (w + 1/2) 2^-32 - 1/2), draw 2 + e = the lateral step of
 * electron e (two 53-bit uniforms -> Box-Muller -> x, y), draw 2^31 + 2 + e = its survival.  The result
 * is independent of the launch geometry; oracle/detector.py reproduces it count for count. */
typedef struct optk_ccd_plane {
    double energy;              /* photon energy, eV                                  */
    double absorption;          /* absorption coefficient of the substrate, 1 / mm    */
    double thickness_implant;   /* mm                                                 */
    double thickness_depletion; /* mm                                                 */
    double thickness_substrate; /* mm                                                 */
    double width_pixel_x;       /* mm                                                 */
    double width_pixel_y;       /* mm                                                 */
    double cce_backsurface;     /* charge-collection efficiency at the back surface   */
    double energy_pair_inf;     /* asymptotic pair-creation energy, eV                */
    double fano_inf;            /* asymptotic Fano factor                             */
    int32_t n_pmf;
    int32_t reserved;
    const double* cmf;          /* device, n_pmf: cumulative probability of n_values  */
    const double* n_values;     /* device, n_pmf: numbers of pairs                    */
} optk_ccd_plane_t;

/* photons: device int64 [n_plane][n_x][n_y] (absorbed photons per pixel); electrons: device uint64, same
 * shape, caller-zeroed, counts are ADDED.  `planes` is a HOST array of n_plane records. */
OPTK_API int optk_electrons_measured(int32_t n_plane, int32_t n_x, int32_t n_y, const optk_ccd_plane_t* planes,
                                     const int64_t* photons, uint64_t* electrons, int32_t wrap, uint64_t seed,
                                     void* stream);

/* ---- host memory ---------------------------------------------------------------
 * Page-lock (cudaHostRegister, portable) / release a host range the caller owns, e.g. a POSIX
 * shared-memory mapping that several single-GPU processes read detector planes back into
 * (optika_b200/distributed.py): copies into registered memory run at PCIe speed and are
 * asynchronous.  The range must stay mapped until it is unregistered. */
OPTK_API int optk_host_register(void* data, int64_t n_bytes);
OPTK_API int optk_host_unregister(void* data);
/* cudaMemcpyAsync(dst, src, n_bytes, cudaMemcpyDefault) on `stream`: device, registered-host or PEER
 * pointers (memory of another GPU opened through CUDA IPC, optika_b200/distributed.py): the copy
 * engines move detector planes over NVLink without occupying an SM. */
OPTK_API int optk_memcpy_async(void* dst, const void* src, int64_t n_bytes, void* stream);
/* cudaDeviceEnablePeerAccess(peer_device) for the CURRENT device (already enabled is not an error):
 * without it a copy from peer memory is staged through the host (20 GB/s instead of NVLink). */
OPTK_API int optk_enable_peer_access(int32_t peer_device);

#ifdef __cplusplus
}
#endif
#endif /* OPTK_H */
