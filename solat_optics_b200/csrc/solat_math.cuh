// solat_math.cuh -- FP64 per-ray geometric optics for the LATR optics tube (near field) and the
// LAT mirrors (far field).  Product code: every function here runs inside the sm_100a kernels of
// solat_kernels.cu.  The functions are also marked __host__ so that tests/host_shadow can compile
// the SAME source with g++ and check it against the golden fixtures in a container without a
// GPU; under nvcc every function here is __device__ only.
//
// Formulation (differs from the reference on purpose; see DESIGN.md "Kernel math"):
//  * Near field: every ot_geo frame is a +-pi/2 tilt about x (ot_geo.py:61-66), so all surfaces
//    are written directly in the telescope frame as  y = -Y0 + sag(x^2+z^2)  and no per-surface
//    frame transform is evaluated.  Intersection = Newton on the ray parameter from the
//    vertex-plane hit (3 updates reach 1e-13 mm) instead of scipy brentq.
//  * The conic term uses  c*u/(1+S) == (1-S)/((1+k)c)  and the gradient collapses to
//    G = c/S + 2a1 + 4a2 u + 6a3 u^2 + 8a4 u^3  (ot_geo.py:125-135 sum to exactly this),
//    so one rsqrt serves both sag and slope.
//  * Far field: all latrt_geo frame changes for one (az, el) pixel are folded into four affine
//    maps that are built once per pixel (FfFrames) -- including the reference's (+el,+az)
//    back-rotation that is not the inverse of its (-el,-az) forward rotation
//    (solat_apert_field.py:249-257 vs latrt_geo.py:280-288), which parity needs.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SOLAT_HD __device__ __forceinline__
#define SOLAT_TABLE __constant__          // coefficient tables live in the constant bank: DFMA reads c[bank][ofs]
#else                                     // directly, where a literal would cost two UMOVs per use
#define SOLAT_HD inline
#define SOLAT_TABLE static const
#endif

// Newton steps after the MUFU seed.  Measured on B200 (tools/variant_bench.py math, profiles/r01_variants.md):
// seed 9.5e-7, 1 step 8.9e-13, 2 steps 1.1e-16 (full); rsqrt third-order step alone 2.2e-16 (full).
#ifndef SOLAT_RCP_ITERS
#define SOLAT_RCP_ITERS 2        // geometry: full precision
#endif
#ifndef SOLAT_RCP_STEP_ITERS
#define SOLAT_RCP_STEP_ITERS 1   // Newton step sizes: 9e-13 relative, self-correcting
#endif
#ifndef SOLAT_RSQRT_POLISH
#define SOLAT_RSQRT_POLISH 0
#endif

namespace solat {

// ------------------------------------------------------------------------------------------
// status flags (per ray / OR-reduced per launch)
// ------------------------------------------------------------------------------------------
enum : int {
    ST_OK = 0,
    ST_BRACKET = 1,   // root outside the reference's brentq bracket (reference: ValueError)
    ST_NOCONV = 2,    // Newton did not reach 1e-9 mm
    ST_NAN = 4        // NaN in the optical path (total internal reflection / sqrt domain)
};

struct vec3 {
    double x, y, z;
};
SOLAT_HD vec3 mk(double x, double y, double z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
SOLAT_HD double dot(const vec3& a, const vec3& b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
SOLAT_HD vec3 axpy(const vec3& p, const vec3& d, double t) { return mk(fma(d.x, t, p.x), fma(d.y, t, p.y), fma(d.z, t, p.z)); }

// Reciprocal / reciprocal-sqrt.  Device: MUFU seed (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64)
// + Newton in FMA form, without the IEEE slow-path branch that the compiler's a/b carries.  Operands on
// this path are O(1e-4 .. 1e9) and never denormal.  ITERS Newton steps; accuracy per variant is measured on
// the device by solat_selftest_math (tests/test_gpu_parity.py::test_fast_math_primitives).
template <int ITERS>
SOLAT_HD double rcp_n(double a)
{
#if defined(__CUDA_ARCH__) && !defined(SOLAT_EXACT_DIV)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#pragma unroll
    for (int i = 0; i < ITERS; i++) {
        const double e = fma(-a, r, 1.0);
        r = fma(r, e, r);
    }
    return r;
#else
    return 1.0 / a;
#endif
}
SOLAT_HD double rcp_fast(double a) { return rcp_n<SOLAT_RCP_ITERS>(a); }      // geometry (path lengths, normals)
SOLAT_HD double rcp_step(double a) { return rcp_n<SOLAT_RCP_STEP_ITERS>(a); } // Newton step sizes: self-correcting
template <int POLISH>
SOLAT_HD double rsqrt_n(double a)
{
#if defined(__CUDA_ARCH__) && !defined(SOLAT_EXACT_DIV)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    // e = 1 - a y^2 ;  y <- y (1 + e/2 + 3 e^2/8)   (third order), then optional second-order polish
    double e = fma(-a * y, y, 1.0);
    y = fma(y * e, fma(0.375, e, 0.5), y);
#pragma unroll
    for (int i = 0; i < POLISH; i++) {
        e = fma(-a * y, y, 1.0);
        y = fma(y * 0.5, e, y);
    }
    return y;
#else
    return 1.0 / sqrt(a);
#endif
}
SOLAT_HD double rsqrt_fast(double a) { return rsqrt_n<SOLAT_RSQRT_POLISH>(a); }

// Snell refraction, same vector law as ray_trace_int.py:20-30 written without cross products:
//   s2 = mu (s - (N.s) N) - N sqrt(1 - mu^2 (1 - (N.s)^2)),   N pointing against propagation.
// Returns NaN components on total internal reflection exactly like the reference's np.sqrt.
SOLAT_HD vec3 refract(const vec3& s, const vec3& N, double mu)
{
    double c = -dot(N, s);
    double k = fma(-mu * mu, fma(-c, c, 1.0), 1.0);
    double f = fma(mu, c, -sqrt(k));
    return mk(fma(f, N.x, mu * s.x), fma(f, N.y, mu * s.y), fma(f, N.z, mu * s.z));
}
SOLAT_HD vec3 reflect(const vec3& s, const vec3& N)
{
    double c = -2.0 * dot(s, N);
    return mk(fma(c, N.x, s.x), fma(c, N.y, s.y), fma(c, N.z, s.z));
}

// ------------------------------------------------------------------------------------------
// Near field: even aspheres of ot_geo.py:81-100,193-212,297-338 and axial offsets :68-71.
// ------------------------------------------------------------------------------------------
struct Asphere {
    double c, K, Kinv;         // 1/R, (1+k)c^2, 1/((1+k)c)
    double a1, a2, a3, a4;     // sag polynomial in u = r^2
    double g1, g2, g3, g4;     // slope polynomial 2a1, 4a2, 6a3, 8a4
    double y0;                 // surface vertex at y = -y0 in the telescope frame
};
constexpr Asphere make_asphere(double R, double k, double a1, double a2, double a3, double a4, double y0)
{
    return Asphere{1.0 / R, (1.0 + k) * (1.0 / R) * (1.0 / R), 1.0 / ((1.0 + k) * (1.0 / R)),
                   a1, a2, a3, a4, 2.0 * a1, 4.0 * a2, 6.0 * a3, 8.0 * a4, y0};
}
constexpr double LENS1_Y = 975.08;                                             // ot_geo.py:68
constexpr double LENS2_Y = 592.76;                                             // ot_geo.py:69
constexpr double LENS3_Y = 180.8 - ((2.5 + 1.513) - (35.16 - 32.19));          // ot_geo.py:70
constexpr double LYOT_Y = 453.62;                                              // ot_geo.py:71
// plane offsets: [0] 1B (ot_geo.py:166), [1] 2B (:277), [2] Lyot (:526)
SOLAT_TABLE double K_NF_PLANE[] = {LENS1_Y, LENS2_Y, LYOT_Y};
#define Y_1B (K_NF_PLANE[0])
#define Y_2B (K_NF_PLANE[1])
#define Y_LYOT (K_NF_PLANE[2])
// ot_geo.py:85-91, :197-203, :301-307, :323-329 with the offsets of :178, :289, :478, :490
enum { SURF_1A = 0, SURF_2A = 1, SURF_3A = 2, SURF_3B = 3 };
struct AsphereF {   // single-precision copy for the coarse first Newton update
    float c, K, a1, a2, a3, a4, g1, g2, g3, g4;
};
constexpr AsphereF make_asphere_f(double R, double k, double a1, double a2, double a3, double a4)
{
    return AsphereF{(float)(1.0 / R), (float)((1.0 + k) * (1.0 / R) * (1.0 / R)), (float)a1, (float)a2, (float)a3, (float)a4,
                    (float)(2.0 * a1), (float)(4.0 * a2), (float)(6.0 * a3), (float)(8.0 * a4)};
}
SOLAT_TABLE AsphereF K_ASPHF[4] = {
    make_asphere_f(-1510.61, 12.0966, -5.876941e-5, 1.769198e-9, -1.310183e-15, 1.994766e-19),
    make_asphere_f(-871.25, -10.201256, -1.147640e-04, 1.103643e-10, -2.969608e-14, 5.980271e-19),
    make_asphere_f(6608.30, 414.142695, -2.981594e-06, 3.955202e-10, 3.180727e-16, 0.0),
    make_asphere_f(604.44, -3.177341, 2.366829e-05, -1.000417e-10, -6.565681e-15, 0.0)};
SOLAT_TABLE Asphere K_ASPH[4] = {
    make_asphere(-1510.61, 12.0966, -5.876941e-5, 1.769198e-9, -1.310183e-15, 1.994766e-19, LENS1_Y - 19.32),
    make_asphere(-871.25, -10.201256, -1.147640e-04, 1.103643e-10, -2.969608e-14, 5.980271e-19, LENS2_Y - 26.47),
    make_asphere(6608.30, 414.142695, -2.981594e-06, 3.955202e-10, 3.180727e-16, 0.0, LENS3_Y),
    make_asphere(604.44, -3.177341, 2.366829e-05, -1.000417e-10, -6.565681e-15, 0.0, LENS3_Y + 35.16)};

// Hit of the ray P + t d with asphere S.  On return H is the hit point, N the unit normal in the
// telescope frame with N.y > 0 (= reference n_hat rotated back), t the ray parameter.
SOLAT_HD float rsqrt_f(float a)
{
#if defined(__CUDA_ARCH__)
    return rsqrtf(a);
#else
    return 1.0f / sqrtf(a);
#endif
}
SOLAT_HD float div_f(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}

// Newton schedule shared by all curved surfaces (DESIGN.md "Mixed-precision Newton"):
//   update 1 from the vertex-plane hit is evaluated in FP32 on the otherwise idle FMA pipe -- its own
//            truncation error (curvature x step^2 ~ 1e-4 .. 1e-2 mm) dwarfs FP32 rounding (~1e-5 mm);
//   updates 2 and 3 are FP64 and converge quadratically to < 1e-12 mm; the unit normal uses the slope of the
//            last evaluation (the last step is < 1e-9 mm, i.e. < 1e-13 rad in the normal).
template <int S>
SOLAT_HD int asphere_hit(const vec3& P, const vec3& d, double lo, double hi, double& t, vec3& H, vec3& N)
{
    const Asphere& A = K_ASPH[S];
    const double py = P.y + A.y0;
    t = -py * rcp_step(d.y);
    {   // coarse update in single precision; on the vertex plane the residual is just -sag
        const AsphereF& F = K_ASPHF[S];
        const float dxf = (float)d.x, dyf = (float)d.y, dzf = (float)d.z;
        const float xf = (float)fma(d.x, t, P.x), zf = (float)fma(d.z, t, P.z);
        const float uf = fmaf(xf, xf, zf * zf);
        const float rsf = rsqrt_f(fmaf(-F.K, uf, 1.0f));
        const float sagf = fmaf(div_f(F.c * uf, 1.0f + div_f(1.0f, rsf)), 1.0f, uf * fmaf(uf, fmaf(uf, fmaf(uf, F.a4, F.a3), F.a2), F.a1));
        const float Gf = fmaf(F.c, rsf, fmaf(uf, fmaf(uf, fmaf(uf, F.g4, F.g3), F.g2), F.g1));
        const float fpf = fmaf(-Gf, fmaf(xf, dxf, zf * dzf), dyf);
        t += (double)div_f(sagf, fpf);
    }
    double x, z, u, w, rs, G, dt = 0.0;
#define SOLAT_ASPH_UPDATE()                                                            \
    {                                                                                  \
        x = fma(d.x, t, P.x);                                                          \
        z = fma(d.z, t, P.z);                                                          \
        u = fma(x, x, z * z);                                                          \
        w = fma(-A.K, u, 1.0);                                                         \
        rs = rsqrt_fast(w);                                                            \
        G = fma(A.c, rs, fma(u, fma(u, fma(u, A.g4, A.g3), A.g2), A.g1));              \
        const double sag = fma(fma(-w, rs, 1.0), A.Kinv, u * fma(u, fma(u, fma(u, A.a4, A.a3), A.a2), A.a1)); \
        const double f = fma(d.y, t, py) - sag;                                        \
        const double fp = fma(-G, fma(x, d.x, z * d.z), d.y);                          \
        dt = f * rcp_step(fp);                                                         \
        t -= dt;                                                                       \
    }
    SOLAT_ASPH_UPDATE()
    SOLAT_ASPH_UPDATE()
    for (int it = 3; fabs(dt) > 1e-9 && it < 10; ++it) SOLAT_ASPH_UPDATE()   // off-nominal rays only
#undef SOLAT_ASPH_UPDATE
    // hit point: (x, z) of the last evaluation moved by the last (sub-1e-9 mm) step
    H = mk(fma(d.x, t, P.x), fma(d.y, t, P.y), fma(d.z, t, P.z));
    const double nn = rsqrt_fast(fma(G * G, u, 1.0));
    N = mk(-x * G * nn, nn, -z * G * nn);
    int st = ST_OK;
    if (!(fabs(dt) <= 1e-9)) st |= (dt != dt) ? ST_NAN : ST_NOCONV;
    if (!(t >= lo && t <= hi)) st |= ST_BRACKET;
    return st;
}

SOLAT_HD int plane_hit(const vec3& P, const vec3& d, double y0, double lo, double hi, double& t, vec3& H)
{
    t = (-y0 - P.y) * rcp_fast(d.y);
    H = mk(fma(d.x, t, P.x), fma(d.y, t, P.y), fma(d.z, t, P.z));
    return (t >= lo && t <= hi) ? ST_OK : ST_BRACKET;
}

// -- source_to_lyot_model.trace_rays, ray_trace_int.py:929-1373 --------------------------------
struct NfSourceRay {
    vec3 lyot, p2b, p2a, p1b, p1a, d_out;   // hit points; d_out = direction arriving at the Lyot stop
    double opl;
    int valid;                              // 0 when vignetted (reference returns a zero column)
    int status;
};
SOLAT_HD void nf_source_ray(const vec3& p0, const vec3& d0, double n_si, double n_vac, NfSourceRay& r)
{
    const double mu_in = n_vac / n_si, mu_out = n_si / n_vac;
    double t1, t2, t3, t4, t5;
    vec3 N, d;
    int st = plane_hit(p0, d0, Y_1B, 5.0, 14000.0, t1, r.p1b);                 // rti:967
    d = refract(d0, mk(0.0, -1.0, 0.0), mu_in);                                 // rti:992
    const int st1a = asphere_hit<SURF_1A>(r.p1b, d, 1.0, 900.0, t2, r.p1a, N); // rti:1033
    r.valid = 0; r.opl = 0.0; r.status = st;
    // rti:1086 (>= 196 -> zero column).  A NaN radius means the iterate left the domain of the conic
    // (r > 417 mm), where the reference's root function substitutes sag = 0 and then vignettes too.
    if (!(sqrt(fma(r.p1a.x, r.p1a.x, r.p1a.z * r.p1a.z)) < 196.0)) return;
    st |= st1a;
    d = refract(d, mk(-N.x, -N.y, -N.z), mu_out);                               // rti:1058
    st |= plane_hit(r.p1a, d, Y_2B, 5.0, 600.0, t3, r.p2b);                    // rti:1130
    r.status = st;
    if (!(sqrt(fma(r.p2b.x, r.p2b.x, r.p2b.z * r.p2b.z)) < 176.0)) return;     // rti:1177
    d = refract(d, mk(0.0, -1.0, 0.0), mu_in);                                  // rti:1154 (-n_hat)
    st |= asphere_hit<SURF_2A>(r.p2b, d, 0.0, 100.0, t4, r.p2a, N);            // rti:1228
    d = refract(d, mk(-N.x, -N.y, -N.z), mu_out);                               // rti:1254
    st |= plane_hit(r.p2a, d, Y_LYOT, 10.0, 900.0, t5, r.lyot);                // rti:1295
    r.d_out = d;
    r.opl = (t1) + (t2 * n_si) + (t3) + (t4 * n_si) + t5;                       // rti:1341-1343
    if (r.opl != r.opl) st |= ST_NAN;
    r.valid = 1;
    r.status = st;
}

// -- rx_to_lyot_model.trace_rays, ray_trace_int.py:614-876 -------------------------------------
struct NfRxRay {
    vec3 lyot, p3b, p3a, d_out;
    double opl;
    int status;
};
SOLAT_HD double dist3(const vec3& a, const vec3& b)
{
    double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return sqrt(fma(dx, dx, fma(dy, dy, dz * dz)));
}
SOLAT_HD void nf_rx_ray(const vec3& p0, const vec3& d0, double n_si, double n_vac, NfRxRay& r)
{
    double t1, t2, t3;
    vec3 N, d;
    int st = asphere_hit<SURF_3A>(p0, d0, 40.0, 700.0, t1, r.p3a, N);          // rti:653
    d = refract(d0, N, n_vac / n_si);                                           // rti:679
    st |= asphere_hit<SURF_3B>(r.p3a, d, 5.0, 200.0, t2, r.p3b, N);            // rti:719
    d = refract(d, N, n_si / n_vac);                                            // rti:743
    st |= plane_hit(r.p3b, d, Y_LYOT, 1.0, 2000.0, t3, r.lyot);                // rti:784
    r.d_out = d;
    // rti:848: Euclidean leg lengths, n_si inside lens 3
    r.opl = (dist3(r.p3b, r.p3a) * n_si) + dist3(r.p3a, p0) + dist3(r.lyot, r.p3b);
    if (r.opl != r.opl) st |= ST_NAN;
    r.status = st;
}

// -- aperature_fields.trace_rays (focal plane -> window toy trace), ray_trace_int.py:70-565 ----
struct NfToyRay {
    vec3 pos_ap, p1b, p1a, p2b, p2a, p3b, p3a, d_out, d_first;
    double opl;
    int status;
};
SOLAT_HD void nf_toy_ray(const vec3& p0, const vec3& d0, double y_ap, double n_si, double n_vac, NfToyRay& r)
{
    const double mu_in = n_vac / n_si, mu_out = n_si / n_vac;
    double t;
    vec3 N, d;
    int st = asphere_hit<SURF_3A>(p0, d0, 2.0, 1600.0, t, r.p3a, N);           // rti:111
    r.d_first = d0;
    d = refract(d0, N, mu_in);                                                  // rti:138
    st |= asphere_hit<SURF_3B>(r.p3a, d, 5.0, 200.0, t, r.p3b, N);             // rti:178
    d = refract(d, N, mu_out);                                                  // rti:202
    st |= asphere_hit<SURF_2A>(r.p3b, d, 10.0, 2000.0, t, r.p2a, N);           // rti:243
    d = refract(d, N, mu_in);                                                   // rti:268
    st |= plane_hit(r.p2a, d, Y_2B, 5.0, 400.0, t, r.p2b);                     // rti:308
    d = refract(d, mk(0.0, 1.0, 0.0), mu_out);                                  // rti:332
    st |= asphere_hit<SURF_1A>(r.p2b, d, 5.0, 500.0, t, r.p1a, N);             // rti:373
    d = refract(d, N, mu_in);                                                   // rti:398
    st |= plane_hit(r.p1a, d, Y_1B, 5.0, 200.0, t, r.p1b);                     // rti:440
    // rti:465: snell(n_si, n_vac, N_hat, -tan) with the 1B normal (tele frame (0,-1,0))
    d = refract(mk(-d.x, -d.y, -d.z), mk(0.0, -1.0, 0.0), mu_out);
    r.d_out = d;
    double d_ap = fabs((y_ap - r.p1b.y) / d.y);                                 // rti:489
    r.opl = dist3(r.p3a, p0) + dist3(r.p3b, r.p3a) * n_si + dist3(r.p2a, r.p3b) + dist3(r.p2b, r.p2a) * n_si +
            dist3(r.p1a, r.p2b) + dist3(r.p1b, r.p1a) * n_si + d_ap;            // rti:490-498
    r.pos_ap = mk(fma(-d_ap, d.x, r.p1b.x), fma(-d_ap, d.y, r.p1b.y), fma(-d_ap, d.z, r.p1b.z));   // rti:500
    if (r.opl != r.opl) st |= ST_NAN;
    r.status = st;
}

// ------------------------------------------------------------------------------------------
// Far field: LAT mirrors, latrt_geo.py:54-150 (2-D polynomials in x/3000, y/3000).
// Only even powers of x occur; Horner in v = y/R per power of u^2.
// ------------------------------------------------------------------------------------------
constexpr double FF_RN = 3000.0;
// [0] 1/R_N, [1] 2/R_N, [2] y_cent_m1 (saf:19), [3] aperture radius^2 (saf:378), [4] aperture plane z (latrt_geo.py:244)
SOLAT_TABLE double K_FF[] = {1.0 / 3000.0, 2.0 / 3000.0, -7201.003729431267, 3e3 * 3e3, 4000.0};
#define FF_INV_RN (K_FF[0])
#define Y_CENT_M1 (K_FF[2])

// Horner tables: rows i = 0, 2, 4, 6 of a1 / a2 (latrt_geo.py:55-61, :63-71) as polynomials in v = y/R,
// followed by their v-derivatives.  z = P0 + w (P2 + w (P4 + w P6)), w = (x/R)^2.
SOLAT_TABLE double K_M1[] = {
    /* P0/v^2 [0..4] */ -57.74022, 1.5373825, 1.154294, -0.441762, 0.0906601,
    /* P2 [5..9]     */ -72.17349, 1.8691899, 2.8859421, -1.026471, 0.2610568,
    /* P4 [10..12]   */ 1.8083973, -0.603195, 0.2177414,
    /* P6 [13]       */ 0.0394559,
    /* dP0/v [14..18]*/ 2 * -57.74022, 3 * 1.5373825, 4 * 1.154294, 5 * -0.441762, 6 * 0.0906601,
    /* dP2 [19..22]  */ 1.8691899, 2 * 2.8859421, 3 * -1.026471, 4 * 0.2610568,
    /* dP4 [23..24]  */ -0.603195, 2 * 0.2177414,
    /* 2P4, 3P6 scale handled in code */};
SOLAT_TABLE double K_M2[] = {
    /* P0/v^2 [0..5] */ 103.90461, 6.6513025, 2.8405781, -0.7819705, -0.0400483, 0.0896645,
    /* P2 [6..11]    */ 115.44758, 7.3024355, 5.7640389, -1.578144, -0.0354326, 0.2781226,
    /* P4 [12..15]   */ 2.9130983, -0.8104051, -0.0185283, 0.2626023,
    /* P6 [16..17]   */ -0.0250794, 0.0709672,
    /* dP0/v [18..23]*/ 2 * 103.90461, 3 * 6.6513025, 4 * 2.8405781, 5 * -0.7819705, 6 * -0.0400483, 7 * 0.0896645,
    /* dP2 [24..28]  */ 7.3024355, 2 * 5.7640389, 3 * -1.578144, 4 * -0.0354326, 5 * 0.2781226,
    /* dP4 [29..31]  */ -0.8104051, 2 * -0.0185283, 3 * 0.2626023,
    /* dP6 [32]      */ 0.0709672};

// single-precision copies (same layout) for the coarse first Newton update
SOLAT_TABLE float K_M1F[] = {
    -57.74022f, 1.5373825f, 1.154294f, -0.441762f, 0.0906601f, -72.17349f, 1.8691899f, 2.8859421f, -1.026471f, 0.2610568f,
    1.8083973f, -0.603195f, 0.2177414f, 0.0394559f,
    2 * -57.74022f, 3 * 1.5373825f, 4 * 1.154294f, 5 * -0.441762f, 6 * 0.0906601f,
    1.8691899f, 2 * 2.8859421f, 3 * -1.026471f, 4 * 0.2610568f, -0.603195f, 2 * 0.2177414f};
SOLAT_TABLE float K_M2F[] = {
    103.90461f, 6.6513025f, 2.8405781f, -0.7819705f, -0.0400483f, 0.0896645f,
    115.44758f, 7.3024355f, 5.7640389f, -1.578144f, -0.0354326f, 0.2781226f,
    2.9130983f, -0.8104051f, -0.0185283f, 0.2626023f, -0.0250794f, 0.0709672f,
    2 * 103.90461f, 3 * 6.6513025f, 4 * 2.8405781f, 5 * -0.7819705f, 6 * -0.0400483f, 7 * 0.0896645f,
    7.3024355f, 2 * 5.7640389f, 3 * -1.578144f, 4 * -0.0354326f, 5 * 0.2781226f,
    -0.8104051f, 2 * -0.0185283f, 3 * 0.2626023f, 0.0709672f};

// The Horner bodies are written once for both precisions (T = double with fma, T = float with fmaf).
#define SOLAT_M1_BODY(T, FMA, K, INV_RN, TWO_INV_RN)                                                     \
    const T u = x * INV_RN, v = y * INV_RN, w = u * u;                                                   \
    const T p0 = v * v * FMA(v, FMA(v, FMA(v, FMA(v, K[4], K[3]), K[2]), K[1]), K[0]);                   \
    const T p2 = FMA(v, FMA(v, FMA(v, FMA(v, K[9], K[8]), K[7]), K[6]), K[5]);                           \
    const T p4 = FMA(v, FMA(v, K[12], K[11]), K[10]);                                                    \
    const T p6 = K[13];                                                                                  \
    z = FMA(w, FMA(w, FMA(w, p6, p4), p2), p0);                                                          \
    const T q0 = v * FMA(v, FMA(v, FMA(v, FMA(v, K[18], K[17]), K[16]), K[15]), K[14]);                  \
    const T q2 = FMA(v, FMA(v, FMA(v, K[22], K[21]), K[20]), K[19]);                                     \
    const T q4 = FMA(v, K[24], K[23]);                                                                   \
    zy = FMA(w, FMA(w, q4, q2), q0) * INV_RN;                                                            \
    zx = TWO_INV_RN * u * FMA(w, FMA(w, (T)3 * p6, (T)2 * p4), p2);
#define SOLAT_M2_BODY(T, FMA, K, INV_RN, TWO_INV_RN)                                                     \
    const T u = x * INV_RN, v = y * INV_RN, w = u * u;                                                   \
    const T p0 = v * v * FMA(v, FMA(v, FMA(v, FMA(v, FMA(v, K[5], K[4]), K[3]), K[2]), K[1]), K[0]);     \
    const T p2 = FMA(v, FMA(v, FMA(v, FMA(v, FMA(v, K[11], K[10]), K[9]), K[8]), K[7]), K[6]);           \
    const T p4 = FMA(v, FMA(v, FMA(v, K[15], K[14]), K[13]), K[12]);                                     \
    const T p6 = FMA(v, K[17], K[16]);                                                                   \
    z = FMA(w, FMA(w, FMA(w, p6, p4), p2), p0);                                                          \
    const T q0 = v * FMA(v, FMA(v, FMA(v, FMA(v, FMA(v, K[23], K[22]), K[21]), K[20]), K[19]), K[18]);   \
    const T q2 = FMA(v, FMA(v, FMA(v, FMA(v, K[28], K[27]), K[26]), K[25]), K[24]);                      \
    const T q4 = FMA(v, FMA(v, K[31], K[30]), K[29]);                                                    \
    const T q6 = K[32];                                                                                  \
    zy = FMA(w, FMA(w, FMA(w, q6, q4), q2), q0) * INV_RN;                                                \
    zx = TWO_INV_RN * u * FMA(w, FMA(w, (T)3 * p6, (T)2 * p4), p2);

SOLAT_HD void m1_sag_f(float x, float y, float& z, float& zx, float& zy)
{
    SOLAT_M1_BODY(float, fmaf, K_M1F, (1.0f / 3000.0f), (2.0f / 3000.0f))
}
SOLAT_HD void m2_sag_f(float x, float y, float& z, float& zx, float& zy)
{
    SOLAT_M2_BODY(float, fmaf, K_M2F, (1.0f / 3000.0f), (2.0f / 3000.0f))
}

// z and its partials wrt the LOCAL coordinates x, y (mm)
SOLAT_HD void m1_sag(double x, double y, double& z, double& zx, double& zy)
{
    SOLAT_M1_BODY(double, fma, K_M1, K_FF[0], K_FF[1])
}
SOLAT_HD void m2_sag(double x, double y, double& z, double& zx, double& zy)
{
    SOLAT_M2_BODY(double, fma, K_M2, K_FF[0], K_FF[1])
}

struct Affine {          // y = A x + b, row-major A
    double a[9];
    double b[3];
};
SOLAT_HD vec3 apply(const Affine& T, const vec3& p)
{
    return mk(fma(T.a[0], p.x, fma(T.a[1], p.y, fma(T.a[2], p.z, T.b[0]))),
              fma(T.a[3], p.x, fma(T.a[4], p.y, fma(T.a[5], p.z, T.b[1]))),
              fma(T.a[6], p.x, fma(T.a[7], p.y, fma(T.a[8], p.z, T.b[2]))));
}
SOLAT_HD vec3 apply_lin(const double* a, const vec3& p)
{
    return mk(fma(a[0], p.x, fma(a[1], p.y, a[2] * p.z)), fma(a[3], p.x, fma(a[4], p.y, a[5] * p.z)),
              fma(a[6], p.x, fma(a[7], p.y, a[8] * p.z)));
}

// Telescope geometry that does not depend on the pixel (latrt_geo.py:7-49, ff_sim.py:72-79)
struct FfGeom {
    double th_1, th2;      // mirror tilts (th2 = -pi/2 - atan(1/3))
    double rotc[3];        // rotation centre [mm]
    double src[3];         // source (tower) position [mm]
    double foc_y;          // focal plane at y' = foc_y in the (+el,+az) frame: rx_y*1e3 (saf:582,955-957)
};

// Everything about one (az, el) pixel.  65 doubles; lives in shared memory.
struct FfFrames {
    double ap_row[3];      // third row of R(-el,-az)
    double ap_num;         // 4000 - C.z - ap_row.(src - C)            (aperture plane, latrt_geo.py:237-245)
    Affine t1;             // telescope -> M1 local        (latrt_geo.py:183-199 with (-el,-az))
    Affine t2;             // telescope -> M2 local        (latrt_geo.py:202-218 with (-el,-az))
    double b1[9];          // M1 local vector -> telescope (saf:215-257: R(+el,+az) Rx(th1) diag(-1,1,-1))
    double b2[9];          // M2 local vector -> telescope (saf:885-930)
    double foc_row[3];     // second row of R(+el,+az)
    double foc_off;        // C.y - foc_row.C
    double centre[3];      // focal-plane centre, saf:323-339
};

SOLAT_HD void mat3_mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// R(el,az) = Ry(az) Rx(el), latrt_geo.py:280-288
SOLAT_HD void rot_el_az(double el, double az, double* R)
{
    const double ce = cos(el), se = sin(el), ca = cos(az), sa = sin(az);
    R[0] = ca;  R[1] = sa * se;  R[2] = sa * ce;
    R[3] = 0.0; R[4] = ce;       R[5] = -se;
    R[6] = -sa; R[7] = ca * se;  R[8] = ca * ce;
}
SOLAT_HD void ff_build_frames(const FfGeom& g, double el, double az, FfFrames& f)
{
    double Rm[9], Rp[9];
    rot_el_az(-el, -az, Rm);
    rot_el_az(el, az, Rp);
    const double C[3] = {g.rotc[0], g.rotc[1], g.rotc[2]};
    // aperture plane
    f.ap_row[0] = Rm[6]; f.ap_row[1] = Rm[7]; f.ap_row[2] = Rm[8];
    f.ap_num = K_FF[4] - C[2] - (Rm[6] * (g.src[0] - C[0]) + Rm[7] * (g.src[1] - C[1]) + Rm[8] * (g.src[2] - C[2]));
    // F1 = diag(-1,1,-1) Rx(-th1), F2 = diag(-1,-1,1) Rx(-th2)
    const double c1 = cos(g.th_1), s1 = sin(g.th_1), c2 = cos(g.th2), s2 = sin(g.th2);
    const double F1[9] = {-1, 0, 0, 0, c1, s1, 0, s1, -c1};
    const double F2[9] = {-1, 0, 0, 0, -c2, -s2, 0, -s2, c2};
    const double o1[3] = {0.0, 7200.0, 3600.0}, o2[3] = {0.0, 4800.0 + 7200.0, 0.0};
    double w[3];   // C - Rm C
    for (int i = 0; i < 3; i++) w[i] = C[i] - (Rm[3 * i] * C[0] + Rm[3 * i + 1] * C[1] + Rm[3 * i + 2] * C[2]);
    mat3_mul(F1, Rm, f.t1.a);
    mat3_mul(F2, Rm, f.t2.a);
    for (int i = 0; i < 3; i++) {
        f.t1.b[i] = F1[3 * i] * (w[0] + o1[0]) + F1[3 * i + 1] * (w[1] + o1[1]) + F1[3 * i + 2] * (w[2] + o1[2]);
        f.t2.b[i] = F2[3 * i] * (w[0] + o2[0]) + F2[3 * i + 1] * (w[1] + o2[1]) + F2[3 * i + 2] * (w[2] + o2[2]);
    }
    // back-rotations use R(+el,+az) (reference behaviour) and the transposes of F1, F2
    const double F1t[9] = {F1[0], F1[3], F1[6], F1[1], F1[4], F1[7], F1[2], F1[5], F1[8]};
    const double F2t[9] = {F2[0], F2[3], F2[6], F2[1], F2[4], F2[7], F2[2], F2[5], F2[8]};
    mat3_mul(Rp, F1t, f.b1);
    mat3_mul(Rp, F2t, f.b2);
    f.foc_row[0] = Rp[3]; f.foc_row[1] = Rp[4]; f.foc_row[2] = Rp[5];
    f.foc_off = C[1] - (Rp[3] * C[0] + Rp[4] * C[1] + Rp[5] * C[2]);
    // focal centre: mean over the r<3600 part of row 50 / column 50 of a 100x100 grid on +-4000 mm
    // pushed through foc_into_tele(el, az) (latrt_geo.py:248-258).  The grid is symmetric, so the
    // means are taken point by point exactly as the reference does.
    const double step = 8000.0 / 99.0, a50 = -4000.0 + 50 * step;
    double sx = 0, sy = 0, sz = 0;
    int cnt = 0;
    for (int i = 0; i < 100; i++) {
        const double ai = (i == 99) ? 4000.0 : -4000.0 + i * step;
        if (sqrt(ai * ai + a50 * a50) < 3600.0) {
            // row 50: (X, Y) = (ai, a50) -> focal->tele (x, 0, Y); column 50: (a50, ai)
            const double rx_ = ai - C[0], ry_ = 0.0 - C[1], rz_ = a50 - C[2];
            sx += Rp[0] * rx_ + Rp[1] * ry_ + Rp[2] * rz_ + C[0];
            const double cx_ = a50 - C[0], cy_ = 0.0 - C[1], cz_ = ai - C[2];
            sy += Rp[3] * cx_ + Rp[4] * cy_ + Rp[5] * cz_ + C[1];
            sz += Rp[6] * cx_ + Rp[7] * cy_ + Rp[8] * cz_ + C[2];
            cnt++;
        }
    }
    f.centre[0] = sx / cnt; f.centre[1] = sy / cnt; f.centre[2] = sz / cnt;
}

// Panel-error polynomial of one panel (rows 2..9 of the reference's panel model, saf:667-674)
struct PanelErr {
    double a, b, c, d, e, x0, y0;
};

// Newton hit of P + t D (mirror-local frame) with mirror M (1 or 2), starting on the local z=0 plane.
// One FP32 coarse update, then two FP64 updates; the gradient of the last evaluation is also used for
// the normal (the last step is ~1e-11 mm, so the normal is exact to 1e-16).  PE: optional panel
// error, evaluated along its own affine image (pe + t de) of the ray (saf:676-703, :820-852).
template <int M, bool PANEL>
SOLAT_HD int mirror_hit(const vec3& P, const vec3& D, double lo, double hi, const PanelErr* pe, const vec3& pE,
                        const vec3& dE, double& t, vec3& H, vec3& N)
{
    t = -P.z * rcp_step(D.z);
    {   // coarse update in single precision (see asphere_hit); on the local z = 0 plane the residual is -z
        const float xf = (float)fma(D.x, t, P.x), yf = (float)fma(D.y, t, P.y);
        float zf, zxf, zyf;
        if (M == 1) m1_sag_f(xf, yf, zf, zxf, zyf); else m2_sag_f(xf, yf, zf, zxf, zyf);
        const float fpf = (float)D.z - fmaf(zxf, (float)D.x, zyf * (float)D.y);
        t += (double)div_f(zf, fpf);
    }
    double dt = 0.0, z, zx, zy;
    auto update = [&]() {
        const double x = fma(D.x, t, P.x), y = fma(D.y, t, P.y);
        if (M == 1) m1_sag(x, y, z, zx, zy); else m2_sag(x, y, z, zx, zy);
        double f = fma(D.z, t, P.z) - z;
        double fp = D.z - fma(zx, D.x, zy * D.y);
        if (PANEL) {
            // xpc = -x_e - x0, ypc = y_e - y0 with (x_e, y_e) the error-frame image of the point
            const double xe = fma(dE.x, t, pE.x), ye = fma(dE.y, t, pE.y);
            const double xpc = -xe - pe->x0, ypc = ye - pe->y0;
            f -= pe->a + pe->b * xpc + pe->c * ypc + pe->d * (xpc * xpc + ypc * ypc) + pe->e * (xpc * ypc);
            fp -= -dE.x * (pe->b + 2.0 * pe->d * xpc + pe->e * ypc) + dE.y * (pe->c + 2.0 * pe->d * ypc + pe->e * xpc);
        }
        dt = f * rcp_step(fp);
        t -= dt;
    };
    update();
    update();                                             // ~1e-11 mm on nominal rays
    for (int it = 3; fabs(dt) > 1e-8 && it < 12; ++it) update();   // off-nominal rays only
    H = mk(fma(D.x, t, P.x), fma(D.y, t, P.y), fma(D.z, t, P.z));
    if (PANEL) {   // reference normal = ideal-mirror gradient at the final point (saf:718-720, :869-871)
        if (M == 1) m1_sag(H.x, H.y, z, zx, zy); else m2_sag(H.x, H.y, z, zx, zy);
    }
    const double nn = rsqrt_fast(fma(zx, zx, fma(zy, zy, 1.0)));
    N = mk(-zx * nn, -zy * nn, nn);
    int st = ST_OK;
    if (!(fabs(dt) <= 1e-8)) st |= (dt != dt) ? ST_NAN : ST_NOCONV;
    if (!(t >= lo && t <= hi)) st |= ST_BRACKET;
    return st;
}

// cheap first stage of every far-field ray: aperture-plane hit and the 3 m radius test
// (saf:359-379 == :644-665).  Returns 1 if the ray continues.
SOLAT_HD int ff_aperture(const FfGeom& g, const FfFrames& f, const vec3& d, double& t_ap, vec3& Pa)
{
    t_ap = f.ap_num * rcp_fast(dot(mk(f.ap_row[0], f.ap_row[1], f.ap_row[2]), d));
    Pa = mk(fma(d.x, t_ap, g.src[0]), fma(d.y, t_ap, g.src[1]), fma(d.z, t_ap, g.src[2]));
    const double dy = Pa.y - Y_CENT_M1;
    return !(fma(Pa.x, Pa.x, dy * dy) > K_FF[3]);
}

struct FfRay {
    vec3 m1, m2, n2, foc, d_out;   // telescope-frame M1/M2 hits, M2 normal, focal hit, outgoing direction
    double opl;
    int status;
};
// Remainder of the ray after ff_aperture.  STAGE1 = ray_mirror_pts brackets (saf:396,490), no focal
// plane; otherwise aperature_fields_from_panel_model brackets (saf:706,855,962).
template <bool STAGE1, bool PANEL>
SOLAT_HD void ff_ray_trace(const FfFrames& f, const vec3& d, double t_ap, const vec3& Pa, double foc_y,
                           const PanelErr* pe1, const PanelErr* pe2, const double* t2e_a, FfRay& r)
{
    double t1, t2;
    vec3 H, N;
    const vec3 none = mk(0, 0, 0);
    // M1
    vec3 p1 = apply(f.t1, Pa), d1 = apply_lin(f.t1.a, d);
    int st = mirror_hit<1, PANEL>(p1, d1, 0.0, STAGE1 ? 5e4 : 1e4, pe1, p1, d1, t1, H, N);
    r.m1 = axpy(Pa, d, t1);
    vec3 dt2 = apply_lin(f.b1, reflect(d1, N));                                // saf:208-257
    // M2
    vec3 p2 = apply(f.t2, r.m1), d2 = apply_lin(f.t2.a, dt2);
    vec3 pE = none, dE = none;
    if (PANEL) {   // error coordinates use the z-stretched point, saf:828-832
        Affine te; for (int i = 0; i < 9; i++) te.a[i] = t2e_a[i];
        te.b[0] = f.t2.b[0]; te.b[1] = f.t2.b[1]; te.b[2] = f.t2.b[2];
        pE = apply(te, r.m1); dE = apply_lin(te.a, dt2);
    }
    st |= mirror_hit<2, PANEL>(p2, d2, STAGE1 ? 0.0 : 1e3, STAGE1 ? 20e3 : 13e3, pe2, pE, dE, t2, H, N);
    r.m2 = axpy(r.m1, dt2, t2);
    r.opl = 0.0;
    if (!STAGE1) {
        r.n2 = apply_lin(f.b2, N);
        vec3 dt3 = apply_lin(f.b2, reflect(d2, N));                            // saf:883-930
        const vec3 fr = mk(f.foc_row[0], f.foc_row[1], f.foc_row[2]);
        const double tf = (foc_y - (dot(fr, r.m2) + f.foc_off)) * rcp_fast(dot(fr, dt3));   // saf:947-962
        if (!(tf >= 1e3 && tf <= 15e3)) st |= ST_BRACKET;
        r.foc = axpy(r.m2, dt3, tf);
        r.d_out = dt3;
        r.opl = ((t_ap + t1) + t2) + tf;                                        // saf:970
        if (r.opl != r.opl) st |= ST_NAN;
    }
    r.status = st;
}

// np.searchsorted(a, v, side='left') over a strided sorted array
SOLAT_HD int searchsorted_left(const double* a, int n, int stride, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[(long)mid * stride] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// exp(i x) for x up to ~1e9 rad with absolute error ~1e-15: three-term Cody-Waite reduction by
// 2 pi in FMA form, then the library sincos on |r| <= pi (fast path).  far_field_latrt.py:28-30
// feeds phases k*L/1e3 ~ 2.9e8 rad, which would take the library's Payne-Hanek slow path.
SOLAT_HD void cis_large(double x, double& c, double& s)
{
    const double INV_2PI = 0.15915494309189535;
    const double P1 = 6.283185307179586, P2 = 2.4492935982947064e-16, P3 = -5.9895396194366790e-33;
    const double n = rint(x * INV_2PI);
    double r = fma(-n, P1, x);
    r = fma(-n, P2, r);
    r = fma(-n, P3, r);
#if defined(__CUDA_ARCH__)
    sincos(r, &s, &c);
#else
    s = sin(r); c = cos(r);
#endif
}

// ------------------------------------------------------------------------------------------
// Row packers: the reference's `.output()` column layouts.
// ------------------------------------------------------------------------------------------
SOLAT_HD vec3 launch_dir(double sin_th, double cos_th, double sin_ph, double cos_ph)
{
    return mk(sin_th * cos_ph, sin_th * sin_ph, cos_th);   // rti:628, saf:345
}

// source_to_lyot_model rows, rti:1346-1373.  rows[26]
SOLAT_HD void pack_nf_source_rows(const NfSourceRay& r, const vec3& p0, double n_si, double n_vac, double* rows)
{
    if (!r.valid) { for (int i = 0; i < 26; i++) rows[i] = 0.0; return; }     // rti:1086-1114
    // rti:1305-1339: Snell (si -> vac) applied to the vector pointing from the Lyot hit back to 2A
    const vec3 og = refract(mk(-r.d_out.x, -r.d_out.y, -r.d_out.z), mk(0.0, 1.0, 0.0), n_si / n_vac);
    const double v[26] = {r.lyot.x, r.lyot.y, r.lyot.z, r.opl, 1.0, 0.0, 1.0, 0.0, og.x, og.y, og.z,
                          r.p2b.x, r.p2b.y, r.p2b.z, r.p2a.x, r.p2a.y, r.p2a.z, r.p1b.x, r.p1b.y, r.p1b.z,
                          r.p1a.x, r.p1a.y, r.p1a.z, p0.x, p0.y, p0.z};
    for (int i = 0; i < 26; i++) rows[i] = v[i];
}
// rx_to_lyot_model rows, rti:855-876.  rows[20].  inv_2sig2 = 0.5 / (th_fwhp / sqrt(8 ln 2))^2
SOLAT_HD void pack_nf_rx_rows(const NfRxRay& r, const vec3& p0, double dth, double dph, double inv_2sig2, double* rows)
{
    const double amp = exp(-(dth * dth + dph * dph) * inv_2sig2);             // rti:849-853
    const double v[20] = {r.lyot.x, r.lyot.y, r.lyot.z, r.opl, amp, 0.0, 1.0, 0.0, r.d_out.x, r.d_out.y, r.d_out.z,
                          r.p3b.x, r.p3b.y, r.p3b.z, r.p3a.x, r.p3a.y, r.p3a.z, p0.x, p0.y, p0.z};
    for (int i = 0; i < 20; i++) rows[i] = v[i];
}
// aperature_fields rows, rti:532-565.  rows[32]
SOLAT_HD void pack_nf_toy_rows(const NfToyRay& r, const vec3& p0, double inv_2sig2, double* rows)
{
    const vec3& t = r.d_first;
    const double de_ve = atan(t.x / (-t.y));                                   // rti:503
    const double de_ho = atan(t.z / sqrt(t.x * t.x + t.y * t.y));              // rti:504-506
    const double amp = exp(-(de_ho * de_ho + de_ve * de_ve) * inv_2sig2);     // rti:527-531
    const double v[32] = {r.pos_ap.x, r.pos_ap.y, r.pos_ap.z, r.p1b.x, r.p1b.y, r.p1b.z, r.p1a.x, r.p1a.y, r.p1a.z,
                          r.p2b.x, r.p2b.y, r.p2b.z, r.p2a.x, r.p2a.y, r.p2a.z, r.p3b.x, r.p3b.y, r.p3b.z,
                          r.p3a.x, r.p3a.y, r.p3a.z, p0.x, p0.y, p0.z, 0.0, 0.0, 1.0, r.d_out.x, r.d_out.y, r.d_out.z,
                          r.opl, amp};
    for (int i = 0; i < 32; i++) rows[i] = v[i];
}
// aperature_fields_from_panel_model rows, saf:985-1008.  rows[17]
SOLAT_HD void pack_ff_rows(const FfRay& r, const double* centre, double* rows)
{
    const double v[17] = {r.m2.x, r.m2.y, r.m2.z, r.m1.x, r.m1.y, r.m1.z, r.n2.x, r.n2.y, r.n2.z,
                          r.foc.x - centre[0], r.foc.y - centre[1], r.foc.z - centre[2],
                          r.d_out.x, r.d_out.y, r.d_out.z, r.opl, 1.0};
    for (int i = 0; i < 17; i++) rows[i] = v[i];
}

// far_field_latrt.py:14-26 for one ray: focal-plane offsets (mm) -> cell index iy*M+ix, or -1 when
// the reference skips the ray.  gx = data_x[0,:] (M), gy = data_y (M x M row-major).
SOLAT_HD int ff_cell(double fx_mm, double fz_mm, const double* gx, const double* gy, int M, double gx_max, double gy_max)
{
    const double xx = fx_mm / 1e1, yy = fz_mm / 1e1;
    if (gx_max < xx || gy_max < yy) return -1;
    const int ix = searchsorted_left(gx, M, 1, xx);
    const int iy = searchsorted_left(gy + ix, M, M, yy);
    return iy * M + ix;
}

}  // namespace solat
