// solat_math.cuh -- FP64 per-ray geometric optics for the LATR optics tube (near field) and the
// LAT mirrors (far field).  Product code: every function here runs inside the sm_100a kernels of
// solat_kernels.cu.  tests/host_shadow compiles the SAME source with g++ so that the per-ray math can
// be checked against the golden fixtures in a container without a GPU; under nvcc every function
// here is __device__ only.
//
// Formulation (differs from the reference on purpose; see DESIGN.md "Kernel math"):
//  * Near field: every ot_geo frame is a +-pi/2 tilt about x (ot_geo.py:61-66), so all surfaces
//    are written directly in the telescope frame as  y = -Y0 + sag(x^2+z^2)  and no per-surface
//    frame transform is evaluated.  Intersection = Newton on the ray parameter from the
//    vertex-plane hit instead of scipy brentq.
//  * The conic term uses  c*u/(1+S) == (1-S)/((1+k)c)  and the gradient collapses to
//    G = c/S + 2a1 + 4a2 u + 6a3 u^2 + 8a4 u^3  (ot_geo.py:125-135 sum to exactly this),
//    so one rsqrt serves both sag and slope.
//  * Far field: all latrt_geo frame changes for one (az, el) pixel are folded into four affine
//    maps that are built once per pixel (FfFrames) -- including the reference's (+el,+az)
//    back-rotation that is not the inverse of its (-el,-az) forward rotation
//    (solat_apert_field.py:249-257 vs latrt_geo.py:280-288), which parity needs.
//  * Mixed-precision Newton: the first update from the vertex-plane hit is evaluated in FP32 on the
//    otherwise idle FMA pipe (its truncation error dwarfs FP32 rounding); two FP64 updates follow.
//  * Every routine is a template over the real type R: `double` traces one ray, `d2` traces two rays in
//    lock step.  The fused map kernels use d2: two independent dependency chains per thread hide the
//    8-cycle DFMA latency, and every coefficient / frame load and every barrier serves two rays.
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define SOLAT_HD __device__ __forceinline__
#define SOLAT_TABLE __constant__          // coefficient tables live in the constant bank (a literal double
#else                                     // costs two UMOVs per use on sm_100)
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

using ::fabs;   // the pack overloads below must not hide the scalar library functions
using ::fma;
using ::fmaf;
using ::sqrt;

// ------------------------------------------------------------------------------------------
// status flags (per ray / OR-reduced per launch)
// ------------------------------------------------------------------------------------------
enum : int {
    ST_OK = 0,
    ST_BRACKET = 1,   // root outside the reference's brentq bracket (reference: ValueError)
    ST_NOCONV = 2,    // Newton did not reach 1e-9 mm
    ST_NAN = 4,       // NaN in the optical path (total internal reflection / sqrt domain)
    ST_RANGE = 8      // far-field path length not exactly representable in the winner key
};

// ------------------------------------------------------------------------------------------
// scalar primitives
// ------------------------------------------------------------------------------------------
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
SOLAT_HD double rcp_fast(double a) { return rcp_n<SOLAT_RCP_ITERS>(a); }      // geometry (path lengths, normals)
SOLAT_HD double rcp_step(double a) { return rcp_n<SOLAT_RCP_STEP_ITERS>(a); } // Newton step sizes
SOLAT_HD double rsqrt_fast(double a) { return rsqrt_n<SOLAT_RSQRT_POLISH>(a); }
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
SOLAT_HD float to_f(double a) { return (float)a; }
SOLAT_HD double to_r(float a, double) { return (double)a; }
SOLAT_HD bool lt(double a, double b) { return a < b; }
SOLAT_HD bool gt(double a, double b) { return a > b; }
SOLAT_HD bool le(double a, double b) { return a <= b; }
SOLAT_HD bool ge(double a, double b) { return a >= b; }
SOLAT_HD bool is_nan(double a) { return a != a; }
// Status tests on the integer pipe (a DSETP costs two FP64-pipe cycles; these compare the high word only, so a
// boundary is resolved to ~1e-6 relative -- ample for "did the root stay in the reference's bracket").
SOLAT_HD int hi_word(double a)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(a);
#else
    long long b;
    memcpy(&b, &a, 8);
    return (int)(b >> 32);
#endif
}
// true unless lo <= t <= hi (0 <= lo < hi); NaN and negative t report out of range
SOLAT_HD bool out_of_range(double t, double lo, double hi)
{
    const unsigned h = (unsigned)hi_word(t);              // negative numbers and NaN have huge unsigned high words
    return h < (unsigned)hi_word(lo) || h > (unsigned)hi_word(hi);
}
SOLAT_HD bool abs_gt(double a, double tol) { return (hi_word(a) & 0x7fffffff) > hi_word(tol); }   // NaN: true
SOLAT_HD bool not_finite(double a) { return (hi_word(a) & 0x7ff00000) == 0x7ff00000; }
SOLAT_HD bool bnot(bool a) { return !a; }
SOLAT_HD bool band(bool a, bool b) { return a && b; }
SOLAT_HD bool bor(bool a, bool b) { return a || b; }
SOLAT_HD bool any(bool a) { return a; }
SOLAT_HD double select(bool m, double a, double b) { return m ? a : b; }
SOLAT_HD int flag(bool m, int bit) { return m ? bit : 0; }
SOLAT_HD int imask(bool m, int v) { return m ? v : 0; }

// ------------------------------------------------------------------------------------------
// two-ray packs: all operations element-wise
// ------------------------------------------------------------------------------------------
struct d2 {
    double a, b;
    SOLAT_HD d2() {}
    SOLAT_HD d2(double s) : a(s), b(s) {}
    SOLAT_HD d2(double x, double y) : a(x), b(y) {}
};
struct f2 {
    float a, b;
    SOLAT_HD f2() {}
    SOLAT_HD f2(float s) : a(s), b(s) {}
    SOLAT_HD f2(float x, float y) : a(x), b(y) {}
};
struct b2 {
    bool a, b;
};
struct i2 {
    int a, b;
    SOLAT_HD i2() {}
    SOLAT_HD i2(int s) : a(s), b(s) {}
    SOLAT_HD i2(int x, int y) : a(x), b(y) {}
};
SOLAT_HD b2 mkb(bool a, bool b) { b2 r; r.a = a; r.b = b; return r; }
SOLAT_HD d2 operator+(d2 x, d2 y) { return d2(x.a + y.a, x.b + y.b); }
SOLAT_HD d2 operator-(d2 x, d2 y) { return d2(x.a - y.a, x.b - y.b); }
SOLAT_HD d2 operator*(d2 x, d2 y) { return d2(x.a * y.a, x.b * y.b); }
SOLAT_HD d2 operator-(d2 x) { return d2(-x.a, -x.b); }
SOLAT_HD d2& operator-=(d2& x, d2 y) { x.a -= y.a; x.b -= y.b; return x; }
SOLAT_HD d2& operator+=(d2& x, d2 y) { x.a += y.a; x.b += y.b; return x; }
SOLAT_HD d2 fma(d2 x, d2 y, d2 z) { return d2(fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)); }
SOLAT_HD d2 sqrt(d2 x) { return d2(sqrt(x.a), sqrt(x.b)); }
SOLAT_HD d2 fabs(d2 x) { return d2(fabs(x.a), fabs(x.b)); }
SOLAT_HD d2 rcp_fast(d2 x) { return d2(rcp_fast(x.a), rcp_fast(x.b)); }
SOLAT_HD d2 rcp_step(d2 x) { return d2(rcp_step(x.a), rcp_step(x.b)); }
SOLAT_HD d2 rsqrt_fast(d2 x) { return d2(rsqrt_fast(x.a), rsqrt_fast(x.b)); }
SOLAT_HD f2 operator+(f2 x, f2 y) { return f2(x.a + y.a, x.b + y.b); }
SOLAT_HD f2 operator-(f2 x, f2 y) { return f2(x.a - y.a, x.b - y.b); }
SOLAT_HD f2 operator*(f2 x, f2 y) { return f2(x.a * y.a, x.b * y.b); }
SOLAT_HD f2 operator-(f2 x) { return f2(-x.a, -x.b); }
SOLAT_HD f2 fmaf(f2 x, f2 y, f2 z) { return f2(fmaf(x.a, y.a, z.a), fmaf(x.b, y.b, z.b)); }
SOLAT_HD f2 rsqrt_f(f2 x) { return f2(rsqrt_f(x.a), rsqrt_f(x.b)); }
SOLAT_HD f2 div_f(f2 x, f2 y) { return f2(div_f(x.a, y.a), div_f(x.b, y.b)); }
SOLAT_HD f2 to_f(d2 x) { return f2((float)x.a, (float)x.b); }
SOLAT_HD d2 to_r(f2 x, d2) { return d2((double)x.a, (double)x.b); }
SOLAT_HD b2 lt(d2 x, d2 y) { return mkb(x.a < y.a, x.b < y.b); }
SOLAT_HD b2 gt(d2 x, d2 y) { return mkb(x.a > y.a, x.b > y.b); }
SOLAT_HD b2 le(d2 x, d2 y) { return mkb(x.a <= y.a, x.b <= y.b); }
SOLAT_HD b2 ge(d2 x, d2 y) { return mkb(x.a >= y.a, x.b >= y.b); }
SOLAT_HD b2 is_nan(d2 x) { return mkb(x.a != x.a, x.b != x.b); }
SOLAT_HD b2 out_of_range(d2 t, double lo, double hi) { return mkb(out_of_range(t.a, lo, hi), out_of_range(t.b, lo, hi)); }
SOLAT_HD b2 abs_gt(d2 a, double tol) { return mkb(abs_gt(a.a, tol), abs_gt(a.b, tol)); }
SOLAT_HD b2 not_finite(d2 a) { return mkb(not_finite(a.a), not_finite(a.b)); }
SOLAT_HD b2 bnot(b2 x) { return mkb(!x.a, !x.b); }
SOLAT_HD b2 band(b2 x, b2 y) { return mkb(x.a && y.a, x.b && y.b); }
SOLAT_HD b2 bor(b2 x, b2 y) { return mkb(x.a || y.a, x.b || y.b); }
SOLAT_HD bool any(b2 x) { return x.a || x.b; }
SOLAT_HD d2 select(b2 m, d2 x, d2 y) { return d2(m.a ? x.a : y.a, m.b ? x.b : y.b); }
SOLAT_HD i2 flag(b2 m, int bit) { return i2(m.a ? bit : 0, m.b ? bit : 0); }
SOLAT_HD i2 imask(b2 m, i2 v) { return i2(m.a ? v.a : 0, m.b ? v.b : 0); }
SOLAT_HD i2 operator|(i2 x, i2 y) { return i2(x.a | y.a, x.b | y.b); }
SOLAT_HD i2& operator|=(i2& x, i2 y) { x.a |= y.a; x.b |= y.b; return x; }

template <class R> struct Pack;
template <> struct Pack<double> { using F = float; using I = int; using B = bool; };
template <> struct Pack<d2> { using F = f2; using I = i2; using B = b2; };

template <class R>
struct vec3t {
    R x, y, z;
};
using vec3 = vec3t<double>;
template <class R> SOLAT_HD vec3t<R> mk(R x, R y, R z) { vec3t<R> r; r.x = x; r.y = y; r.z = z; return r; }
SOLAT_HD vec3 mk(double x, double y, double z) { return mk<double>(x, y, z); }
template <class R> SOLAT_HD R dot(const vec3t<R>& a, const vec3t<R>& b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
template <class R> SOLAT_HD vec3t<R> axpy(const vec3t<R>& p, const vec3t<R>& d, R t)
{
    return mk<R>(fma(d.x, t, p.x), fma(d.y, t, p.y), fma(d.z, t, p.z));
}
template <class R> SOLAT_HD vec3t<R> neg(const vec3t<R>& a) { return mk<R>(-a.x, -a.y, -a.z); }

// sqrt without the library's correctly-rounded slow path: k * rsqrt(k), <= 2 ulp; NaN for k < 0 like sqrt.
template <class R> SOLAT_HD R sqrt_fast(R k) { return k * rsqrt_fast(k); }

// Snell refraction, same vector law as ray_trace_int.py:20-30 written without cross products:
//   s2 = mu (s - (N.s) N) - N sqrt(1 - mu^2 (1 - (N.s)^2)),   N pointing against propagation.
// Returns NaN components on total internal reflection exactly like the reference's np.sqrt.
template <class R>
SOLAT_HD vec3t<R> refract(const vec3t<R>& s, const vec3t<R>& N, double mu)
{
    const R c = -dot(N, s);
    const R k = fma(R(-mu * mu), fma(-c, c, R(1.0)), R(1.0));
    const R f = fma(R(mu), c, -sqrt_fast(k));
    return mk<R>(fma(f, N.x, mu * s.x), fma(f, N.y, mu * s.y), fma(f, N.z, mu * s.z));
}
// refraction at a flat whose normal is (0, ny, 0), ny = +-1 pointing against propagation
template <class R>
SOLAT_HD vec3t<R> refract_flat(const vec3t<R>& s, double ny, double mu)
{
    const R c = -(ny * s.y);
    const R k = fma(R(-mu * mu), fma(-c, c, R(1.0)), R(1.0));
    const R f = fma(R(mu), c, -sqrt_fast(k));
    // the x and z components keep the reference's NaN behaviour: f * 0 is NaN when f is
    return mk<R>(fma(f, R(0.0), mu * s.x), fma(f, R(ny), mu * s.y), fma(f, R(0.0), mu * s.z));
}
// refraction INTO the denser medium (mu < 1: no total internal reflection) at a flat with normal (0, -1, 0), ray
// travelling towards +y: (mu sx, sqrt(1 - mu^2 (1 - sy^2)), mu sz)
template <class R>
SOLAT_HD vec3t<R> refract_flat_in(const vec3t<R>& s, double mu)
{
    const R k = fma(R(mu * mu), fma(s.y, s.y, R(-1.0)), R(1.0));
    return mk<R>(mu * s.x, sqrt_fast(k), mu * s.z);
}
template <class R>
SOLAT_HD vec3t<R> reflect(const vec3t<R>& s, const vec3t<R>& N)
{
    const R c = R(-2.0) * dot(s, N);
    return mk<R>(fma(c, N.x, s.x), fma(c, N.y, s.y), fma(c, N.z, s.z));
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
struct AsphereF {              // single-precision copy for the coarse first Newton update
    float c, K, a1, a2, a3, a4, g1, g2, g3, g4, Kinv;
};
constexpr Asphere make_asphere(double R, double k, double a1, double a2, double a3, double a4, double y0)
{
    return Asphere{1.0 / R, (1.0 + k) * (1.0 / R) * (1.0 / R), 1.0 / ((1.0 + k) * (1.0 / R)),
                   a1, a2, a3, a4, 2.0 * a1, 4.0 * a2, 6.0 * a3, 8.0 * a4, y0};
}
constexpr AsphereF make_asphere_f(double R, double k, double a1, double a2, double a3, double a4)
{
    return AsphereF{(float)(1.0 / R), (float)((1.0 + k) * (1.0 / R) * (1.0 / R)), (float)a1, (float)a2, (float)a3, (float)a4,
                    (float)(2.0 * a1), (float)(4.0 * a2), (float)(6.0 * a3), (float)(8.0 * a4), (float)(1.0 / ((1.0 + k) * (1.0 / R)))};
}
constexpr double LENS1_Y = 975.08;                                             // ot_geo.py:68
constexpr double LENS2_Y = 592.76;                                             // ot_geo.py:69
constexpr double LENS3_Y = 180.8 - ((2.5 + 1.513) - (35.16 - 32.19));          // ot_geo.py:70
constexpr double LYOT_Y = 453.62;                                              // ot_geo.py:71
// ot_geo.py:85-91, :197-203, :301-307, :323-329 with the offsets of :178, :289, :478, :490
enum { SURF_1A = 0, SURF_2A = 1, SURF_3A = 2, SURF_3B = 3 };
enum { PLANE_1B = 0, PLANE_2B = 1, PLANE_LYOT = 2 };      // ot_geo.py:166, :277, :526
SOLAT_TABLE AsphereF K_ASPHF[4] = {
    make_asphere_f(-1510.61, 12.0966, -5.876941e-5, 1.769198e-9, -1.310183e-15, 1.994766e-19),
    make_asphere_f(-871.25, -10.201256, -1.147640e-04, 1.103643e-10, -2.969608e-14, 5.980271e-19),
    make_asphere_f(6608.30, 414.142695, -2.981594e-06, 3.955202e-10, 3.180727e-16, 0.0),
    make_asphere_f(604.44, -3.177341, 2.366829e-05, -1.000417e-10, -6.565681e-15, 0.0)};

// All FP64 coefficients of the path in one table in the constant bank (every routine takes it by reference, so a
// kernel could stage it elsewhere: a shared-memory copy was measured 2-4 % slower on B200, profiles/r01_variants.md).
struct Coef {
    double m1[14];        // LAT primary,  rows i = 0,2,4,6 of a1 (latrt_geo.py:55-61) as polynomials in v = y/R
    double m2[18];        // LAT secondary, rows of a2 (latrt_geo.py:63-71)
    double ff[5];         // 1/R_N, 2/R_N, y_cent_m1 (saf:19), aperture radius^2 (saf:378), aperture plane z (latrt_geo.py:244)
    double plane[3];      // near-field plane offsets 1B, 2B, Lyot
    Asphere asph[4];      // near-field aspheres 1A, 2A, 3A, 3B
};
SOLAT_TABLE Coef K_COEF = {
    {/* H0 */ -57.74022, 1.5373825, 1.154294, -0.441762, 0.0906601, /* P2 */ -72.17349, 1.8691899, 2.8859421, -1.026471, 0.2610568,
     /* P4 */ 1.8083973, -0.603195, 0.2177414, /* P6 */ 0.0394559},
    {/* H0 */ 103.90461, 6.6513025, 2.8405781, -0.7819705, -0.0400483, 0.0896645,
     /* P2 */ 115.44758, 7.3024355, 5.7640389, -1.578144, -0.0354326, 0.2781226,
     /* P4 */ 2.9130983, -0.8104051, -0.0185283, 0.2626023, /* P6 */ -0.0250794, 0.0709672},
    {1.0 / 3000.0, 2.0 / 3000.0, -7201.003729431267, 3e3 * 3e3, 4000.0},
    {LENS1_Y, LENS2_Y, LYOT_Y},
    {make_asphere(-1510.61, 12.0966, -5.876941e-5, 1.769198e-9, -1.310183e-15, 1.994766e-19, LENS1_Y - 19.32),
     make_asphere(-871.25, -10.201256, -1.147640e-04, 1.103643e-10, -2.969608e-14, 5.980271e-19, LENS2_Y - 26.47),
     make_asphere(6608.30, 414.142695, -2.981594e-06, 3.955202e-10, 3.180727e-16, 0.0, LENS3_Y),
     make_asphere(604.44, -3.177341, 2.366829e-05, -1.000417e-10, -6.565681e-15, 0.0, LENS3_Y + 35.16)}};

// Newton schedule shared by all curved surfaces (DESIGN.md "Mixed-precision Newton"):
//   update 1 from the vertex-plane hit is evaluated in FP32 on the otherwise idle FMA pipe -- its own
//            truncation error (curvature x step^2 ~ 1e-4 .. 1e-2 mm) dwarfs FP32 rounding (~1e-5 mm);
//   mirrors: updates 2 and 3 are FP64 and converge quadratically to < 1e-12 mm; the unit normal uses the slope of
//            the last evaluation (the last step is < 1e-9 mm, i.e. < 1e-13 rad in the normal);
//   lenses:  one FP64 update already lands within ~1e-13 mm (checked: more only while |dt| > 1e-3 mm), then only
//            the slope is re-evaluated at the final point.
// Hit of the ray P + t d with asphere S.  On return H is the hit point, N the unit normal in the
// telescope frame with N.y > 0 (= reference n_hat rotated back), t the ray parameter.
template <int S, class R>
SOLAT_HD typename Pack<R>::I asphere_hit(const Coef& C, const vec3t<R>& P, const vec3t<R>& d, double lo, double hi, R& t,
                                         vec3t<R>& H, vec3t<R>& N)
{
    using F = typename Pack<R>::F;
    using I = typename Pack<R>::I;
    const Asphere& A = C.asph[S];
    const R py = P.y + A.y0;
    t = -py * rcp_step(d.y);
    {   // coarse update in single precision; on the vertex plane the residual is just -sag
        const AsphereF& Af = K_ASPHF[S];
        const F dxf = to_f(d.x), dyf = to_f(d.y), dzf = to_f(d.z);
        const F xf = to_f(fma(d.x, t, P.x)), zf = to_f(fma(d.z, t, P.z));
        const F uf = fmaf(xf, xf, zf * zf);
        const F wf = fmaf(F(-Af.K), uf, F(1.0f));
        const F rsf = rsqrt_f(wf);
        // conic term in the division-free form of the FP64 evaluation, (1 - S) / ((1 + k) c) with S = w rsqrt(w): the
        // cancellation costs ~1e-5 mm in single precision, far below this step's own truncation error
        const F sagf = fmaf(fmaf(-wf, rsf, F(1.0f)), F(Af.Kinv),
                            uf * fmaf(uf, fmaf(uf, fmaf(uf, F(Af.a4), F(Af.a3)), F(Af.a2)), F(Af.a1)));
        const F Gf = fmaf(F(Af.c), rsf, fmaf(uf, fmaf(uf, fmaf(uf, F(Af.g4), F(Af.g3)), F(Af.g2)), F(Af.g1)));
        const F fpf = fmaf(-Gf, fmaf(xf, dxf, zf * dzf), dyf);
        t += to_r(div_f(sagf, fpf), t);
    }
    R x, z, u, w, rs, G, dt = R(0.0);
#define SOLAT_ASPH_SLOPE()                                                                         \
    x = fma(d.x, t, P.x);                                                                          \
    z = fma(d.z, t, P.z);                                                                          \
    u = fma(x, x, z * z);                                                                          \
    w = fma(R(-A.K), u, R(1.0));                                                                   \
    rs = rsqrt_fast(w);                                                                            \
    G = fma(R(A.c), rs, fma(u, fma(u, fma(u, R(A.g4), R(A.g3)), R(A.g2)), R(A.g1)));
#define SOLAT_ASPH_UPDATE()                                                                        \
    {                                                                                              \
        SOLAT_ASPH_SLOPE()                                                                         \
        const R sag = fma(fma(-w, rs, R(1.0)), R(A.Kinv), u * fma(u, fma(u, fma(u, R(A.a4), R(A.a3)), R(A.a2)), R(A.a1))); \
        const R f = fma(d.y, t, py) - sag;                                                         \
        const R fp = fma(-G, fma(x, d.x, z * d.z), d.y);                                           \
        dt = f * rcp_step(fp);                                                                     \
        t -= dt;                                                                                   \
    }
    // One FP64 update after the FP32 one leaves an error of C dt^2 with C <~ 4e-5 / mm for these lenses (nominal
    // rays: dt ~ 1e-4 mm, error ~ 1e-13 mm).  More updates only while |dt| > 1e-3 mm (error bound 4e-11 mm).
    SOLAT_ASPH_UPDATE()
    for (int it = 2; any(abs_gt(dt, 1e-3)) && it < 10; ++it) SOLAT_ASPH_UPDATE()
    SOLAT_ASPH_SLOPE()                     // slope at the converged point, for the normal
#undef SOLAT_ASPH_UPDATE
#undef SOLAT_ASPH_SLOPE
    H = mk<R>(x, fma(d.y, t, P.y), z);
    const R nn = rsqrt_fast(fma(G * G, u, R(1.0)));
    N = mk<R>(-x * G * nn, nn, -z * G * nn);
    I st = imask(bnot(not_finite(dt)), flag(abs_gt(dt, 1e-3), ST_NOCONV)) | flag(not_finite(dt), ST_NAN);
    st |= flag(out_of_range(t, lo, hi), ST_BRACKET);
    return st;
}

template <class R>
SOLAT_HD typename Pack<R>::I plane_hit(const vec3t<R>& P, const vec3t<R>& d, double y0, double lo, double hi, R& t,
                                       vec3t<R>& H)
{
    t = (R(-y0) - P.y) * rcp_fast(d.y);
    H = axpy(P, d, t);
    return flag(out_of_range(t, lo, hi), ST_BRACKET);
}

// -- source_to_lyot_model.trace_rays, ray_trace_int.py:929-1373 --------------------------------
template <class R>
struct NfSourceRayT {
    vec3t<R> lyot, p2b, p2a, p1b, p1a, d_out;   // hit points; d_out = direction arriving at the Lyot stop
    R opl;
    typename Pack<R>::B valid;                  // false when vignetted (reference returns a zero column)
    typename Pack<R>::I status;
};
using NfSourceRay = NfSourceRayT<double>;
// The ray from the flat front of lens 1 on: `d` is the direction inside the silicon, `st` the status of the 1B hit.
template <class R>
SOLAT_HD void nf_source_ray_from_1b(const Coef& C, const vec3t<R>& d_in, R t1, typename Pack<R>::I st, double n_si, double n_vac,
                                    NfSourceRayT<R>& r)
{
    using I = typename Pack<R>::I;
    using B = typename Pack<R>::B;
    const double mu_in = n_vac / n_si, mu_out = n_si / n_vac;
    R t2, t3, t4, t5;
    vec3t<R> N, d = d_in;
    const I st1a = asphere_hit<SURF_1A>(C, r.p1b, d, 1.0, 900.0, t2, r.p1a, N);   // rti:1033
    // rti:1086 (radius >= 196 -> zero column; compared squared).  A NaN radius means the iterate left the domain of the conic
    // (r > 417 mm), where the reference's root function substitutes sag = 0 and then vignettes too.
    B alive = lt(fma(r.p1a.x, r.p1a.x, r.p1a.z * r.p1a.z), R(196.0 * 196.0));
    r.valid = alive; r.opl = R(0.0); r.status = st;
    if (!any(alive)) return;
    st |= imask(alive, st1a);
    d = refract(d, neg(N), mu_out);                                             // rti:1058
    st |= imask(alive, plane_hit(r.p1a, d, C.plane[PLANE_2B], 5.0, 600.0, t3, r.p2b));      // rti:1130
    alive = band(alive, lt(fma(r.p2b.x, r.p2b.x, r.p2b.z * r.p2b.z), R(176.0 * 176.0)));   // rti:1177
    r.valid = alive; r.status = st;
    if (!any(alive)) return;
    d = refract_flat_in(d, mu_in);                                              // rti:1154 (-n_hat)
    st |= imask(alive, asphere_hit<SURF_2A>(C, r.p2b, d, 0.0, 100.0, t4, r.p2a, N));   // rti:1228
    d = refract(d, neg(N), mu_out);                                             // rti:1254
    st |= imask(alive, plane_hit(r.p2a, d, C.plane[PLANE_LYOT], 10.0, 900.0, t5, r.lyot));  // rti:1295
    r.d_out = d;
    const R opl = (t1) + (t2 * n_si) + (t3) + (t4 * n_si) + t5;                 // rti:1341-1343
    st |= flag(band(alive, not_finite(opl)), ST_NAN);
    r.opl = select(alive, opl, R(0.0));
    r.status = st;
}
template <class R>
SOLAT_HD void nf_source_ray(const Coef& C, const vec3t<R>& p0, const vec3t<R>& d0, double n_si, double n_vac, NfSourceRayT<R>& r)
{
    R t1;
    const typename Pack<R>::I st = plane_hit(p0, d0, C.plane[PLANE_1B], 5.0, 14000.0, t1, r.p1b);   // rti:967
    nf_source_ray_from_1b(C, refract_flat_in(d0, n_vac / n_si), t1, st, n_si, n_vac, r);     // rti:992
}
// What of a source ray does not depend on the source's x, z (every source of a map sits on the plane y = y_ap and the
// front of lens 1 is flat): the path to the 1B plane, the direction after it and the hit's y.  The hit itself is
// (x + t1 d0.x, hit_y, z + t1 d0.z).  Two double4 per ray: {t1, d0.x, hit_y, d0.z}, {d.x, d.y, d.z, status}.
SOLAT_HD void nf_source_first_leg(const Coef& C, const vec3& d0, double y_ap, double n_si, double n_vac, double* a4, double* b4)
{
    double t1;
    vec3 H;
    const int st = plane_hit(mk(0.0, y_ap, 0.0), d0, C.plane[PLANE_1B], 5.0, 14000.0, t1, H);
    const vec3 d = refract_flat_in(d0, n_vac / n_si);
    a4[0] = t1; a4[1] = d0.x; a4[2] = H.y; a4[3] = d0.z;
    b4[0] = d.x; b4[1] = d.y; b4[2] = d.z; b4[3] = (double)st;
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
SOLAT_HD void nf_rx_ray(const Coef& C, const vec3& p0, const vec3& d0, double n_si, double n_vac, NfRxRay& r)
{
    double t1, t2, t3;
    vec3 N, d;
    int st = asphere_hit<SURF_3A>(C, p0, d0, 40.0, 700.0, t1, r.p3a, N);          // rti:653
    d = refract(d0, N, n_vac / n_si);                                           // rti:679
    st |= asphere_hit<SURF_3B>(C, r.p3a, d, 5.0, 200.0, t2, r.p3b, N);            // rti:719
    d = refract(d, N, n_si / n_vac);                                            // rti:743
    st |= plane_hit(r.p3b, d, C.plane[PLANE_LYOT], 1.0, 2000.0, t3, r.lyot);                // rti:784
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
SOLAT_HD void nf_toy_ray(const Coef& C, const vec3& p0, const vec3& d0, double y_ap, double n_si, double n_vac, NfToyRay& r)
{
    const double mu_in = n_vac / n_si, mu_out = n_si / n_vac;
    double t;
    vec3 N, d;
    int st = asphere_hit<SURF_3A>(C, p0, d0, 2.0, 1600.0, t, r.p3a, N);           // rti:111
    r.d_first = d0;
    d = refract(d0, N, mu_in);                                                  // rti:138
    st |= asphere_hit<SURF_3B>(C, r.p3a, d, 5.0, 200.0, t, r.p3b, N);             // rti:178
    d = refract(d, N, mu_out);                                                  // rti:202
    st |= asphere_hit<SURF_2A>(C, r.p3b, d, 10.0, 2000.0, t, r.p2a, N);           // rti:243
    d = refract(d, N, mu_in);                                                   // rti:268
    st |= plane_hit(r.p2a, d, C.plane[PLANE_2B], 5.0, 400.0, t, r.p2b);                     // rti:308
    d = refract_flat(d, 1.0, mu_out);                                           // rti:332
    st |= asphere_hit<SURF_1A>(C, r.p2b, d, 5.0, 500.0, t, r.p1a, N);             // rti:373
    d = refract(d, N, mu_in);                                                   // rti:398
    st |= plane_hit(r.p1a, d, C.plane[PLANE_1B], 5.0, 200.0, t, r.p1b);                     // rti:440
    // rti:465: snell(n_si, n_vac, N_hat, -tan) with the 1B normal (tele frame (0,-1,0))
    d = refract_flat(neg(d), -1.0, mu_out);
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
// Coefficient layout of Coef::m1 / Coef::m2: rows i = 0, 2, 4, 6 of a1 / a2 as polynomials in v = y/R:
// z = P0 + w (P2 + w (P4 + w P6)), w = (x/R)^2, P0 = v^2 H0(v).   m1: H0 [0..4] P2 [5..9] P4 [10..12] P6 [13];
// m2: H0 [0..5] P2 [6..11] P4 [12..15] P6 [16..17].
// single-precision copies (same layout) for the coarse first Newton update
SOLAT_TABLE float K_M1F[] = {-57.74022f, 1.5373825f, 1.154294f, -0.441762f, 0.0906601f, -72.17349f, 1.8691899f,
                             2.8859421f, -1.026471f, 0.2610568f, 1.8083973f, -0.603195f, 0.2177414f, 0.0394559f};
SOLAT_TABLE float K_M2F[] = {103.90461f, 6.6513025f, 2.8405781f, -0.7819705f, -0.0400483f, 0.0896645f,
                             115.44758f, 7.3024355f, 5.7640389f, -1.578144f, -0.0354326f, 0.2781226f,
                             2.9130983f, -0.8104051f, -0.0185283f, 0.2626023f, -0.0250794f, 0.0709672f};

// Value and v-derivative of each row polynomial by dual-number Horner (p <- p v + c, dp <- dp v + p): every
// coefficient is fetched once per evaluation -- on sm_100 a DFMA cannot take a constant-bank operand, so each
// coefficient costs a load, and separate derivative tables would double them.  Written once for both precisions.
#define SOLAT_DUAL_START(T, FMA, K, hi, p, dp) T p = FMA(T(K[hi]), v, T(K[(hi) - 1])), dp = T(K[hi]);
#define SOLAT_DUAL_STEP(T, FMA, K, k, p, dp) dp = FMA(dp, v, p); p = FMA(p, v, T(K[k]));
#define SOLAT_M1_BODY(T, FMA, K, INV_RN, TWO_INV_RN)                                                             \
    const T u = x * T(INV_RN), v = y * T(INV_RN), w = u * u;                                                     \
    SOLAT_DUAL_START(T, FMA, K, 4, h0, dh0)                                                                      \
    SOLAT_DUAL_STEP(T, FMA, K, 2, h0, dh0) SOLAT_DUAL_STEP(T, FMA, K, 1, h0, dh0) SOLAT_DUAL_STEP(T, FMA, K, 0, h0, dh0) \
    SOLAT_DUAL_START(T, FMA, K, 9, p2, dp2)                                                                      \
    SOLAT_DUAL_STEP(T, FMA, K, 7, p2, dp2) SOLAT_DUAL_STEP(T, FMA, K, 6, p2, dp2) SOLAT_DUAL_STEP(T, FMA, K, 5, p2, dp2) \
    SOLAT_DUAL_START(T, FMA, K, 12, p4, dp4)                                                                     \
    SOLAT_DUAL_STEP(T, FMA, K, 10, p4, dp4)                                                                      \
    const T p6 = T(K[13]);                                                                                       \
    const T vv = v * v;                                                                                          \
    z = FMA(w, FMA(w, FMA(w, p6, p4), p2), vv * h0);                                                             \
    const T dp0 = v * FMA(v, dh0, h0 + h0);                                                                      \
    zy = FMA(w, FMA(w, dp4, dp2), dp0) * T(INV_RN);                                                              \
    zx = T(TWO_INV_RN) * u * FMA(w, FMA(w, p6 + p6 + p6, p4 + p4), p2);
#define SOLAT_M2_BODY(T, FMA, K, INV_RN, TWO_INV_RN)                                                             \
    const T u = x * T(INV_RN), v = y * T(INV_RN), w = u * u;                                                     \
    SOLAT_DUAL_START(T, FMA, K, 5, h0, dh0)                                                                      \
    SOLAT_DUAL_STEP(T, FMA, K, 3, h0, dh0) SOLAT_DUAL_STEP(T, FMA, K, 2, h0, dh0) SOLAT_DUAL_STEP(T, FMA, K, 1, h0, dh0) \
    SOLAT_DUAL_STEP(T, FMA, K, 0, h0, dh0)                                                                       \
    SOLAT_DUAL_START(T, FMA, K, 11, p2, dp2)                                                                     \
    SOLAT_DUAL_STEP(T, FMA, K, 9, p2, dp2) SOLAT_DUAL_STEP(T, FMA, K, 8, p2, dp2) SOLAT_DUAL_STEP(T, FMA, K, 7, p2, dp2) \
    SOLAT_DUAL_STEP(T, FMA, K, 6, p2, dp2)                                                                       \
    SOLAT_DUAL_START(T, FMA, K, 15, p4, dp4)                                                                     \
    SOLAT_DUAL_STEP(T, FMA, K, 13, p4, dp4) SOLAT_DUAL_STEP(T, FMA, K, 12, p4, dp4)                              \
    const T dp6 = T(K[17]);                                                                                      \
    const T p6 = FMA(dp6, v, T(K[16]));                                                                          \
    const T vv = v * v;                                                                                          \
    z = FMA(w, FMA(w, FMA(w, p6, p4), p2), vv * h0);                                                             \
    const T dp0 = v * FMA(v, dh0, h0 + h0);                                                                      \
    zy = FMA(w, FMA(w, FMA(w, dp6, dp4), dp2), dp0) * T(INV_RN);                                                 \
    zx = T(TWO_INV_RN) * u * FMA(w, FMA(w, p6 + p6 + p6, p4 + p4), p2);

// z and its partials wrt the LOCAL coordinates x, y (mm); M = 1 primary, 2 secondary
template <int M, class R>
SOLAT_HD void mirror_sag(const Coef& C, R x, R y, R& z, R& zx, R& zy)
{
    if (M == 1) { SOLAT_M1_BODY(R, fma, C.m1, C.ff[0], C.ff[1]) } else { SOLAT_M2_BODY(R, fma, C.m2, C.ff[0], C.ff[1]) }
}
template <int M, class F>
SOLAT_HD void mirror_sag_f(F x, F y, F& z, F& zx, F& zy)
{
    if (M == 1) { SOLAT_M1_BODY(F, fmaf, K_M1F, (1.0f / 3000.0f), (2.0f / 3000.0f)) }
    else { SOLAT_M2_BODY(F, fmaf, K_M2F, (1.0f / 3000.0f), (2.0f / 3000.0f)) }
}

struct Affine {          // y = A x + b, row-major A
    double a[9];
    double b[3];
};
template <class R>
SOLAT_HD vec3t<R> apply(const Affine& T, const vec3t<R>& p)
{
    return mk<R>(fma(R(T.a[0]), p.x, fma(R(T.a[1]), p.y, fma(R(T.a[2]), p.z, R(T.b[0])))),
                 fma(R(T.a[3]), p.x, fma(R(T.a[4]), p.y, fma(R(T.a[5]), p.z, R(T.b[1])))),
                 fma(R(T.a[6]), p.x, fma(R(T.a[7]), p.y, fma(R(T.a[8]), p.z, R(T.b[2])))));
}
template <class R>
SOLAT_HD vec3t<R> apply_lin(const double* a, const vec3t<R>& p)
{
    return mk<R>(fma(R(a[0]), p.x, fma(R(a[1]), p.y, R(a[2]) * p.z)), fma(R(a[3]), p.x, fma(R(a[4]), p.y, R(a[5]) * p.z)),
                 fma(R(a[6]), p.x, fma(R(a[7]), p.y, R(a[8]) * p.z)));
}

// Telescope geometry that does not depend on the pixel (latrt_geo.py:7-49, ff_sim.py:72-79)
struct FfGeom {
    double th_1, th2;      // mirror tilts (th2 = -pi/2 - atan(1/3))
    double rotc[3];        // rotation centre [mm]
    double src[3];         // source (tower) position [mm]
    double foc_y;          // focal plane at y' = foc_y in the (+el,+az) frame: rx_y*1e3 (saf:582,955-957)
};

// Everything about one (az, el) pixel; lives in shared memory.
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
    double centre_mz;      // z of the focal-plane centre of the mirror pixel (-az, el); its x is -centre[0] (see ff_map_kernel)
    float apf[8];          // single-precision copy for the aperture prefilter: ap_row[3], ap_num, src.x, src.y - y_cent, R^2, margin
    // composites used by the fused map (ff_ray_lean): the ray stays in mirror-local coordinates from M1 on
    Affine t21;            // M1 local point  -> M2 local point   (t2 o t1^-1)
    double b12[9];         // M1 local vector -> M2 local vector  (t2.a * b1)
    double foc_l[4];       // focal-plane row in M2-local point coordinates: foc_row . (t2^-1 H) + foc_off = foc_l[0..2].H + foc_l[3]
    double foc_d[3];       // foc_row . (b2 v) = foc_d . v  for an M2-local vector v
    double fx_l[4], fx_d[3];   // telescope x of an M2-local point / vector
    double fz_l[4], fz_d[3];   // telescope z of an M2-local point / vector
};

SOLAT_HD void mat3_mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// R(el,az) = Ry(az) Rx(el), latrt_geo.py:280-288
SOLAT_HD void rot_from_trig(double se, double ce, double sa, double ca, double* R)
{
    R[0] = ca;  R[1] = sa * se;  R[2] = sa * ce;
    R[3] = 0.0; R[4] = ce;       R[5] = -se;
    R[6] = -sa; R[7] = ca * se;  R[8] = ca * ce;
}
SOLAT_HD void ff_build_frames(const Coef& C_, const FfGeom& g, double el, double az, FfFrames& f)
{
    double Rm[9], Rp[9];
    // sin is odd and cos even to the bit, so R(-el, -az) comes from the same four values as R(el, az)
    double se, ce, sa, ca;
    sincos(el, &se, &ce);
    sincos(az, &sa, &ca);
    rot_from_trig(-se, ce, -sa, ca, Rm);
    rot_from_trig(se, ce, sa, ca, Rp);
    const double C[3] = {g.rotc[0], g.rotc[1], g.rotc[2]};
    // aperture plane
    f.ap_row[0] = Rm[6]; f.ap_row[1] = Rm[7]; f.ap_row[2] = Rm[8];
    f.ap_num = C_.ff[4] - C[2] - (Rm[6] * (g.src[0] - C[0]) + Rm[7] * (g.src[1] - C[1]) + Rm[8] * (g.src[2] - C[2]));
    // F1 = diag(-1,1,-1) Rx(-th1), F2 = diag(-1,-1,1) Rx(-th2)
    double c1, s1, c2, s2;
    sincos(g.th_1, &s1, &c1);
    sincos(g.th2, &s2, &c2);
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
    {   // t1, t2 are rigid (orthogonal matrix A, offset b):  p = A^T (l - b)
        double A1t[9], A2t[9], t1i_b[3], t2i_b[3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) { A1t[3 * i + j] = f.t1.a[3 * j + i]; A2t[3 * i + j] = f.t2.a[3 * j + i]; }
        for (int i = 0; i < 3; i++) {
            t1i_b[i] = -(A1t[3 * i] * f.t1.b[0] + A1t[3 * i + 1] * f.t1.b[1] + A1t[3 * i + 2] * f.t1.b[2]);
            t2i_b[i] = -(A2t[3 * i] * f.t2.b[0] + A2t[3 * i + 1] * f.t2.b[1] + A2t[3 * i + 2] * f.t2.b[2]);
        }
        mat3_mul(f.t2.a, A1t, f.t21.a);
        for (int i = 0; i < 3; i++)
            f.t21.b[i] = f.t2.a[3 * i] * t1i_b[0] + f.t2.a[3 * i + 1] * t1i_b[1] + f.t2.a[3 * i + 2] * t1i_b[2] + f.t2.b[i];
        mat3_mul(f.t2.a, f.b1, f.b12);
        for (int j = 0; j < 3; j++) {
            f.foc_l[j] = f.foc_row[0] * A2t[j] + f.foc_row[1] * A2t[3 + j] + f.foc_row[2] * A2t[6 + j];
            f.foc_d[j] = f.foc_row[0] * f.b2[j] + f.foc_row[1] * f.b2[3 + j] + f.foc_row[2] * f.b2[6 + j];
            f.fx_l[j] = A2t[j];      f.fx_d[j] = f.b2[j];
            f.fz_l[j] = A2t[6 + j];  f.fz_d[j] = f.b2[6 + j];
        }
        f.foc_l[3] = f.foc_row[0] * t2i_b[0] + f.foc_row[1] * t2i_b[1] + f.foc_row[2] * t2i_b[2] + f.foc_off;
        f.fx_l[3] = t2i_b[0];
        f.fz_l[3] = t2i_b[2];
    }
    f.apf[0] = (float)f.ap_row[0]; f.apf[1] = (float)f.ap_row[1]; f.apf[2] = (float)f.ap_row[2]; f.apf[3] = (float)f.ap_num;
    f.apf[4] = (float)g.src[0]; f.apf[5] = (float)(g.src[1] - C_.ff[2]); f.apf[6] = (float)C_.ff[3];
    // |r^2 - R^2| below this is left to the FP64 test: FP32 carries ~3e-7 relative on a ~5 m lever arm, i.e.
    // ~2e-3 mm in radius, 10 mm^2 in r^2; the margin is 20x that (an annulus holding ~1e-5 of the rays)
    f.apf[7] = 2.0e-5f * (float)C_.ff[3];
    // Focal centre (saf:323-339): means over the r < 3600 mm part of row 50 / column 50 of a 100x100 grid on
    // +-4000 mm pushed through foc_into_tele(el, az) (latrt_geo.py:248-258).  The grid is symmetric about 0, so
    // the masked mean of the running coordinate is 0 (to the reference's own 1e-12 mm rounding) and the means
    // are the images of (0, 0, a50) [x component] and (a50, 0, 0) [y, z components], a50 = -4000 + 50*8000/99.
    const double a50 = -4000.0 + 50.0 * (8000.0 / 99.0);
    f.centre[0] = Rp[0] * (0.0 - C[0]) + Rp[1] * (0.0 - C[1]) + Rp[2] * (a50 - C[2]) + C[0];
    f.centre[1] = Rp[3] * (a50 - C[0]) + Rp[4] * (0.0 - C[1]) + Rp[5] * (0.0 - C[2]) + C[1];
    f.centre[2] = Rp[6] * (a50 - C[0]) + Rp[7] * (0.0 - C[1]) + Rp[8] * (0.0 - C[2]) + C[2];
    f.centre_mz = (-Rp[6]) * (a50 - C[0]) + Rp[7] * (0.0 - C[1]) + Rp[8] * (0.0 - C[2]) + C[2];   // R(el, -az): only Rp[6] flips
}

// Panel-error polynomial of one panel (rows 2..9 of the reference's panel model, saf:667-674)
struct PanelErr {
    double a, b, c, d, e, x0, y0;
};

// Newton hit of P + t D (mirror-local frame) with mirror M (1 or 2), starting on the local z=0 plane.
// One FP32 coarse update, then two FP64 updates; the gradient of the last evaluation is also used for
// the normal (the last step is ~1e-11 mm, so the normal is exact to 1e-16).  PE (scalar rays only): optional
// panel error, evaluated along its own affine image (pe + t de) of the ray (saf:676-703, :820-852).
// status of a mirror hit from its ray parameter and last Newton step
template <class R>
SOLAT_HD typename Pack<R>::I mirror_status(R t, R dt, double lo, double hi)
{
    typename Pack<R>::I st = imask(bnot(not_finite(dt)), flag(abs_gt(dt, 1e-8), ST_NOCONV)) | flag(not_finite(dt), ST_NAN);
    st |= flag(out_of_range(t, lo, hi), ST_BRACKET);
    return st;
}
template <int M, bool PANEL, class R>
SOLAT_HD typename Pack<R>::I mirror_hit(const Coef& C, const vec3t<R>& P, const vec3t<R>& D, double lo, double hi, const PanelErr* pe,
                                        const vec3t<R>& pE, const vec3t<R>& dE, R& t, vec3t<R>& H, vec3t<R>& N, R* dt_out = nullptr)
{
    using F = typename Pack<R>::F;
    using I = typename Pack<R>::I;
    t = -P.z * rcp_step(D.z);
    {   // coarse update in single precision (see asphere_hit); on the local z = 0 plane the residual is -z
        const F xf = to_f(fma(D.x, t, P.x)), yf = to_f(fma(D.y, t, P.y));
        F zf, zxf, zyf;
        mirror_sag_f<M, F>(xf, yf, zf, zxf, zyf);
        const F fpf = to_f(D.z) - fmaf(zxf, to_f(D.x), zyf * to_f(D.y));
        t += to_r(div_f(zf, fpf), t);
    }
    R dt = R(0.0), z, zx, zy;
    auto update = [&]() {
        const R x = fma(D.x, t, P.x), y = fma(D.y, t, P.y);
        mirror_sag<M, R>(C, x, y, z, zx, zy);
        R f = fma(D.z, t, P.z) - z;
        R fp = D.z - fma(zx, D.x, zy * D.y);
        if (PANEL) {
            // xpc = -x_e - x0, ypc = y_e - y0 with (x_e, y_e) the error-frame image of the point
            const R xe = fma(dE.x, t, pE.x), ye = fma(dE.y, t, pE.y);
            const R xpc = -xe - R(pe->x0), ypc = ye - R(pe->y0);
            f -= R(pe->a) + R(pe->b) * xpc + R(pe->c) * ypc + R(pe->d) * (xpc * xpc + ypc * ypc) + R(pe->e) * (xpc * ypc);
            fp -= -dE.x * (R(pe->b) + R(2.0 * pe->d) * xpc + R(pe->e) * ypc) + dE.y * (R(pe->c) + R(2.0 * pe->d) * ypc + R(pe->e) * xpc);
        }
        dt = f * rcp_step(fp);
        t -= dt;
    };
    update();
    update();                                             // ~1e-11 mm on nominal rays
    for (int it = 3; any(abs_gt(dt, 1e-8)) && it < 12; ++it) update();   // off-nominal rays only
    H = axpy(P, D, t);
    if (PANEL) mirror_sag<M, R>(C, H.x, H.y, z, zx, zy);   // reference normal = ideal-mirror gradient (saf:718-720, :869-871)
    const R nn = rsqrt_fast(fma(zx, zx, fma(zy, zy, R(1.0))));
    N = mk<R>(-zx * nn, -zy * nn, nn);
    if (dt_out) { *dt_out = dt; return I(0); }            // the caller forms the status itself (mirror_status)
    return mirror_status(t, dt, lo, hi);
}

// cheap first stage of every far-field ray: aperture-plane hit and the 3 m radius test
// (saf:359-379 == :644-665).  Returns the lanes that continue.
template <class R>
SOLAT_HD typename Pack<R>::B ff_aperture(const Coef& C, const FfGeom& g, const FfFrames& f, const vec3t<R>& d, R& t_ap, vec3t<R>& Pa)
{
    t_ap = R(f.ap_num) * rcp_fast(fma(R(f.ap_row[0]), d.x, fma(R(f.ap_row[1]), d.y, R(f.ap_row[2]) * d.z)));
    Pa = mk<R>(fma(d.x, t_ap, R(g.src[0])), fma(d.y, t_ap, R(g.src[1])), fma(d.z, t_ap, R(g.src[2])));
    const R dy = Pa.y - R(C.ff[2]);
    return bnot(gt(fma(Pa.x, Pa.x, dy * dy), R(C.ff[3])));
}

// unit launch direction of ray (theta, phi)
SOLAT_HD vec3 launch_dir(double sin_th, double cos_th, double sin_ph, double cos_ph)
{
    return mk(sin_th * cos_ph, sin_th * sin_ph, cos_th);   // rti:628, saf:345
}

// Accepted rays of one phi row of a launch fan whose theta is strictly monotone inside one branch of tan: the rays
// of the row hit the aperture plane along a straight line, monotonically, so the squared radius about the M1 centre
// is unimodal in the theta index and the accepted rays (ff_aperture, the same FP64 operations) are ONE interval
// [lo, hi).  First a guess: the chord of the aperture circle on the line through the row's two end rays, taken as
// uniform in the index (true for the reference's narrow linspace fans); it is accepted only if the exact test
// confirms both rim crossings (accepted at lo and hi-1, rejected at lo-1 and hi), 6 evaluations per row.  Otherwise
// bisection: the closest ray, then both crossings, ~4 log2(nth) evaluations.  Either way the interval is exactly the
// set a per-ray test would accept.  Returns false when the plane-hit denominator changes sign along the row (rays
// skimming the aperture plane's horizon: the caller must test every ray instead).
SOLAT_HD bool ff_row_interval(const Coef& C, const FfGeom& g, const FfFrames& f, const double* st, const double* ct, int nth,
                              double sp, double cp, int& lo, int& hi)
{
    auto hit = [&](int i, double& px, double& py, double& den) {
        const vec3 d = launch_dir(st[i], ct[i], sp, cp);
        den = fma(f.ap_row[0], d.x, fma(f.ap_row[1], d.y, f.ap_row[2] * d.z));
        vec3 Pa;
        double t_ap;
        ff_aperture(C, g, f, d, t_ap, Pa);
        px = Pa.x; py = Pa.y - C.ff[2];
        return fma(px, px, py * py);
    };
    auto r2 = [&](int i) {
        double px, py, den;
        return hit(i, px, py, den);
    };
    auto rejected = [&](int i) { return r2(i) > C.ff[3]; };
    double x0, y0, x1, y1, den0, den1;
    const double r2_0 = hit(0, x0, y0, den0), r2_1 = hit(nth - 1, x1, y1, den1);
    const bool one_sided = den0 * den1 > 0.0;
    {   // chord guess
        const double dx = x1 - x0, dy = y1 - y0;
        const double qa = fma(dx, dx, dy * dy), qb = fma(x0, dx, y0 * dy), qc = r2_0 - C.ff[3];
        const double disc = fma(qb, qb, -qa * qc);
        if (disc > 0.0 && qa > 0.0) {
            const double sq = sqrt(disc), scale = (double)(nth - 1) / qa;
            const double s_lo = (-qb - sq) * scale, s_hi = (-qb + sq) * scale;
            if (s_hi >= 0.0 && s_lo <= (double)(nth - 1)) {
                const int gl = s_lo > 0.0 ? (int)ceil(s_lo) : 0;
                const int gh = s_hi < (double)(nth - 1) ? (int)floor(s_hi) + 1 : nth;
                if (gl < gh && gl < nth && gh > 0) {
                    const bool ok_lo = (gl == 0 ? !(r2_0 > C.ff[3]) : (!rejected(gl) && rejected(gl - 1)));
                    const bool ok_hi = ok_lo && (gh == nth ? !(r2_1 > C.ff[3]) : (!rejected(gh - 1) && rejected(gh)));
                    if (ok_hi) { lo = gl; hi = gh; return one_sided; }
                }
            }
        }
    }
    int a = 0, b = nth - 1;
    while (a < b) {
        const int m = (a + b) >> 1;
        if (r2(m + 1) < r2(m)) a = m + 1; else b = m;
    }
    const int imin = a;
    lo = hi = 0;
    if (!rejected(imin)) {
        a = -1; b = imin;                                    // a rejected (or before the row), b accepted
        while (a + 1 < b) {
            const int m = (a + b) >> 1;
            if (rejected(m)) a = m; else b = m;
        }
        lo = b;
        a = imin; b = nth;                                   // a accepted, b rejected (or past the row)
        while (a + 1 < b) {
            const int m = (a + b) >> 1;
            if (rejected(m)) b = m; else a = m;
        }
        hi = b;
    }
    return one_sided;
}

// Single-precision prefilter of the same test: +1 inside, 0 outside, -1 too close to the rim to call in FP32
// (the caller then runs ff_aperture in FP64, so the accepted set is exactly the FP64 one).
SOLAT_HD int ff_aperture_coarse(const FfFrames& f, float dx, float dy, float dz)
{
    const float den = fmaf(f.apf[0], dx, fmaf(f.apf[1], dy, f.apf[2] * dz));
    const float t = div_f(f.apf[3], den);
    const float xa = fmaf(dx, t, f.apf[4]), ya = fmaf(dy, t, f.apf[5]);
    const float e = fmaf(xa, xa, ya * ya) - f.apf[6];
    if (!(fabsf(e) > f.apf[7])) return -1;
    return e > 0.0f ? 0 : 1;
}

template <class R>
struct FfRayT {
    vec3t<R> m1, m2, n2, foc, d_out;   // telescope-frame M1/M2 hits, M2 normal, focal hit, outgoing direction
    R opl;
    typename Pack<R>::I status;
};
using FfRay = FfRayT<double>;
// Remainder of the ray after ff_aperture.  STAGE1 = ray_mirror_pts brackets (saf:396,490), no focal
// plane; otherwise aperature_fields_from_panel_model brackets (saf:706,855,962).
template <bool STAGE1, bool PANEL, class R>
SOLAT_HD void ff_ray_trace(const Coef& C, const FfFrames& f, const vec3t<R>& d, R t_ap, const vec3t<R>& Pa, double foc_y,
                           const PanelErr* pe1, const PanelErr* pe2, const double* t2e_a, FfRayT<R>& r)
{
    using I = typename Pack<R>::I;
    R t1, t2;
    vec3t<R> H, N;
    const vec3t<R> none = mk<R>(R(0.0), R(0.0), R(0.0));
    // M1
    const vec3t<R> p1 = apply(f.t1, Pa), d1 = apply_lin(f.t1.a, d);
    I st = mirror_hit<1, PANEL, R>(C, p1, d1, 0.0, STAGE1 ? 5e4 : 1e4, pe1, p1, d1, t1, H, N);
    r.m1 = axpy(Pa, d, t1);
    const vec3t<R> dt2 = apply_lin(f.b1, reflect(d1, N));                       // saf:208-257
    // M2
    const vec3t<R> p2 = apply(f.t2, r.m1), d2 = apply_lin(f.t2.a, dt2);
    vec3t<R> pE = none, dE = none;
    if (PANEL) {   // error coordinates use the z-stretched point, saf:828-832
        Affine te;
        for (int i = 0; i < 9; i++) te.a[i] = t2e_a[i];
        te.b[0] = f.t2.b[0]; te.b[1] = f.t2.b[1]; te.b[2] = f.t2.b[2];
        pE = apply(te, r.m1); dE = apply_lin(te.a, dt2);
    }
    st |= mirror_hit<2, PANEL, R>(C, p2, d2, STAGE1 ? 0.0 : 1e3, STAGE1 ? 20e3 : 13e3, pe2, pE, dE, t2, H, N);
    r.m2 = axpy(r.m1, dt2, t2);
    r.opl = R(0.0);
    if (!STAGE1) {
        r.n2 = apply_lin(f.b2, N);
        const vec3t<R> dt3 = apply_lin(f.b2, reflect(d2, N));                   // saf:883-930
        const vec3t<R> fr = mk<R>(R(f.foc_row[0]), R(f.foc_row[1]), R(f.foc_row[2]));
        const R tf = (R(foc_y) - (dot(fr, r.m2) + R(f.foc_off))) * rcp_fast(dot(fr, dt3));   // saf:947-962
        st |= flag(out_of_range(tf, 1e3, 15e3), ST_BRACKET);
        r.foc = axpy(r.m2, dt3, tf);
        r.d_out = dt3;
        r.opl = ((t_ap + t1) + t2) + tf;                                        // saf:970
        st |= flag(not_finite(r.opl), ST_NAN);
    }
    r.status = st;
}

// The same ray for the fused map, where only the path length and the focal-plane (x, z) are needed: after the
// aperture point the ray never returns to the telescope frame (M1 local -> M2 local -> focal row), which saves a
// third of the frame arithmetic and, more importantly, keeps nothing but (point, direction, path sum) alive
// across each Newton solve.  Differs from ff_ray_trace only by rounding (1e-12 mm).
template <class R>
SOLAT_HD void ff_ray_lean(const Coef& C, const FfFrames& f, const vec3t<R>& d, R t_ap, const vec3t<R>& Pa, double foc_y,
                          R& opl, R& fx, R& fz, typename Pack<R>::I& status)
{
    using I = typename Pack<R>::I;
    R t1, t2;
    vec3t<R> H, N;
    const vec3t<R> none = mk<R>(R(0.0), R(0.0), R(0.0));
    const vec3t<R> p1 = apply(f.t1, Pa), d1 = apply_lin(f.t1.a, d);
    R dt1, dt2;
    mirror_hit<1, false, R>(C, p1, d1, 0.0, 1e4, nullptr, none, none, t1, H, N, &dt1);
    const vec3t<R> p2 = apply(f.t21, H), d2 = apply_lin(f.b12, reflect(d1, N));
    mirror_hit<2, false, R>(C, p2, d2, 1e3, 13e3, nullptr, none, none, t2, H, N, &dt2);
    const vec3t<R> r2 = reflect(d2, N);
    const R num = R(foc_y) - fma(R(f.foc_l[0]), H.x, fma(R(f.foc_l[1]), H.y, fma(R(f.foc_l[2]), H.z, R(f.foc_l[3]))));
    const R den = fma(R(f.foc_d[0]), r2.x, fma(R(f.foc_d[1]), r2.y, R(f.foc_d[2]) * r2.z));
    const R tf = num * rcp_fast(den);                                                        // saf:947-962
    fx = fma(tf, fma(R(f.fx_d[0]), r2.x, fma(R(f.fx_d[1]), r2.y, R(f.fx_d[2]) * r2.z)),
             fma(R(f.fx_l[0]), H.x, fma(R(f.fx_l[1]), H.y, fma(R(f.fx_l[2]), H.z, R(f.fx_l[3])))));
    fz = fma(tf, fma(R(f.fz_d[0]), r2.x, fma(R(f.fz_d[1]), r2.y, R(f.fz_d[2]) * r2.z)),
             fma(R(f.fz_l[0]), H.x, fma(R(f.fz_l[1]), H.y, fma(R(f.fz_l[2]), H.z, R(f.fz_l[3])))));
    opl = ((t_ap + t1) + t2) + tf;                                                           // saf:970
    // Status: nominal rays raise none of the six conditions, so they are tested together (integer compares on the high
    // words) and the individual bits are formed only on the rare ray that trips one.
    const auto odd = bor(bor(bor(out_of_range(t1, 0.0, 1e4), abs_gt(dt1, 1e-8)), bor(out_of_range(t2, 1e3, 13e3), abs_gt(dt2, 1e-8))),
                         bor(out_of_range(tf, 1e3, 15e3), not_finite(opl)));
    I st = I(0);
    if (any(odd)) {
        st = mirror_status(t1, dt1, 0.0, 1e4) | mirror_status(t2, dt2, 1e3, 13e3);
        st |= flag(out_of_range(tf, 1e3, 15e3), ST_BRACKET);
        st |= flag(not_finite(opl), ST_NAN);
    }
    status = st;
}

// Where on the focal plane will a ray land?  For one pixel the focal-plane position is a smooth function of the ray's
// launch direction cosines: an affine map (the beam is a defocused image of the aperture) plus aberrations of a few
// mm.  The fused map fits a CUBIC polynomial in the two direction cosines per pixel to a sample of exactly traced
// rays (worst error over a whole aperture 0.1-0.4 mm; the affine fit of round 1 left 2.5 mm), measures its worst
// error on the sample, and uses it ONLY to drop rays that the reference's projection would skip anyway (ffl:19:
// x > max(data_x[0, :]) or y > max(data_y[:, 0])) when they miss the high edge of the E_rec grid by more than
// 5 x that error + 2 mm (and rays certain to pile into cell (0, 0), of which one per row is kept, see
// ff_pred_row_keep); every other ray is traced in FP64 and binned exactly.  With the +-5 cm grid of ff_sim.py about
// half of the accepted rays land beyond the grid.
constexpr float FF_PRED_SCALE = 2.0e4f;   // launch direction cosines of the fan are ~5e-5
constexpr int FF_PRED_SAMPLES = 64;
constexpr int FF_PRED_NB = 10;            // basis 1, u, v, u^2, uv, v^2, u^3, u^2 v, u v^2, v^3
struct FfPred {
    float cx[FF_PRED_NB], cz[FF_PRED_NB];   // x - centre.x = sum cx[k] b_k(u, v) [mm], u = d.x * FF_PRED_SCALE, v = d.y * FF_PRED_SCALE
    float margin_cm;                        // (5 x worst sample error + 2 mm) / 10
    int on;
};
// Position of sample k in the pixel's dense list of n_list accepted rays: one per stratum of n_list / SAMPLES rays,
// at a golden-ratio offset inside it -- an even stride alone can put every sample at the same place of its phi row
// (a stride that is a multiple of the row length), which leaves the cubic unconstrained along the rows.
SOLAT_HD int ff_pred_sample_pos(int k, int n_list)
{
    const double g = (double)k * 0.6180339887498949;
    const double pos = ((double)k + (g - floor(g))) * ((double)n_list / (double)FF_PRED_SAMPLES);
    const int v = (int)pos;
    return v < 0 ? 0 : (v >= n_list ? n_list - 1 : v);
}
template <class T>
SOLAT_HD void ff_pred_basis(T u, T v, T* b)
{
    b[0] = T(1); b[1] = u; b[2] = v; b[3] = u * u; b[4] = u * v; b[5] = v * v;
    b[6] = b[3] * u; b[7] = b[3] * v; b[8] = u * b[5]; b[9] = b[5] * v;
}
// Least squares over the sample.  The normal equations live in an augmented 10 x 12 matrix A (row-major: the
// symmetric 10 x 10 matrix of basis products, then the right-hand sides for x and z); they are solved by Gaussian
// elimination without pivoting (the matrix is positive definite; u, v span [-1, 1]: well conditioned in FP64) in steps
// a warp can share: entry (a, b) <- sum over the samples, row i <- row i - (A[i][k] / A[k][k]) row k, back substitution.
constexpr int FF_PRED_LD = FF_PRED_NB + 2;
// samples: [n][FF_PRED_LD] = basis values b_0..b_9, X, Z of every sample ray.  Entry (a, c) of A, c < 12.
SOLAT_HD double ff_pred_normal_entry(const double* samples, int n, int a, int c)
{
    double acc = 0.0;
    for (int k = 0; k < n; k++) acc = fma(samples[k * FF_PRED_LD + a], samples[k * FF_PRED_LD + c], acc);
    return acc;
}
SOLAT_HD void ff_pred_eliminate(double* A, int k, int i)
{
    const double f = A[i * FF_PRED_LD + k] / A[k * FF_PRED_LD + k];
    for (int j = k; j < FF_PRED_LD; j++) A[i * FF_PRED_LD + j] = fma(-f, A[k * FF_PRED_LD + j], A[i * FF_PRED_LD + j]);
}
// A after the elimination (upper triangular), diag0 = the diagonal before it: a pivot that lost 12 digits means a rank
// deficient sample (e.g. all rays on one row): predictor off
SOLAT_HD void ff_pred_backsub(const double* A, const double* diag0, FfPred& p)
{
    bool ok = true;
    for (int k = 0; k < FF_PRED_NB; k++) {
        const double d = A[k * FF_PRED_LD + k];
        if (!(d > 1e-12 * diag0[k]) || !(d == d)) ok = false;
    }
    p.on = ok;
    p.margin_cm = 0.0f;
    for (int which = 0; which < 2; which++) {
        double y[FF_PRED_NB];
        float* o = which ? p.cz : p.cx;
        for (int i = FF_PRED_NB - 1; i >= 0; i--) {
            double v = A[i * FF_PRED_LD + FF_PRED_NB + which];
            for (int k = i + 1; k < FF_PRED_NB; k++) v = fma(-A[i * FF_PRED_LD + k], y[k], v);
            y[i] = ok ? v / A[i * FF_PRED_LD + i] : 0.0;
            o[i] = (float)y[i];
        }
    }
}
// the whole solve on one thread (host shadow; the kernel spreads the same steps over a warp)
SOLAT_HD void ff_pred_fit(const double* samples, int n, FfPred& p)
{
    double A[FF_PRED_NB * FF_PRED_LD], diag0[FF_PRED_NB];
    for (int a = 0; a < FF_PRED_NB; a++)
        for (int c = 0; c < FF_PRED_LD; c++) A[a * FF_PRED_LD + c] = ff_pred_normal_entry(samples, n, c < FF_PRED_NB && c < a ? c : a, c < FF_PRED_NB && c < a ? a : c);
    for (int a = 0; a < FF_PRED_NB; a++) diag0[a] = A[a * FF_PRED_LD + a];
    for (int k = 0; k < FF_PRED_NB - 1; k++)
        for (int i = k + 1; i < FF_PRED_NB; i++) ff_pred_eliminate(A, k, i);
    ff_pred_backsub(A, diag0, p);
}
SOLAT_HD void ff_pred_eval(const FfPred& p, float u, float v, float& x, float& z)
{
    float b[FF_PRED_NB];
    ff_pred_basis<float>(u, v, b);
    x = p.cx[0]; z = p.cz[0];
    for (int k = 1; k < FF_PRED_NB; k++) { x = fmaf(p.cx[k], b[k], x); z = fmaf(p.cz[k], b[k], z); }
}
// Along one phi row u = s cos(phi), v = s sin(phi) with s = sin(theta_i) * FF_PRED_SCALE: the prediction is the cubic
// a[0] + a[1] s + a[2] s^2 + a[3] s^3
SOLAT_HD void ff_pred_row_cubic(const float* c, float cpf, float spf, float* a)
{
    a[0] = c[0];
    a[1] = fmaf(c[1], cpf, c[2] * spf);
    a[2] = fmaf(c[3] * cpf, cpf, fmaf(c[4] * cpf, spf, c[5] * spf * spf));
    a[3] = fmaf(c[6] * cpf * cpf, cpf, fmaf(c[7] * cpf * cpf, spf, fmaf(c[8] * cpf * spf, spf, c[9] * spf * spf * spf)));
}
// Range of the cubic on [s0, s1] and whether it is strictly monotone there (its derivative keeps one sign at both
// ends and has no root inside).
SOLAT_HD bool ff_pred_row_range(const float* a, float s0, float s1, float& pmin, float& pmax)
{
    const float lo = fminf(s0, s1), hi = fmaxf(s0, s1);
    auto val = [&](float t) { return fmaf(fmaf(fmaf(a[3], t, a[2]), t, a[1]), t, a[0]); };
    auto der = [&](float t) { return fmaf(fmaf(3.0f * a[3], t, 2.0f * a[2]), t, a[1]); };
    const float v0 = val(lo), v1 = val(hi);
    pmin = fminf(v0, v1); pmax = fmaxf(v0, v1);
    bool mono = der(lo) * der(hi) > 0.0f;
    // roots of the derivative 3 a3 t^2 + 2 a2 t + a1
    float r[2];
    int nr = 0;
    if (a[3] != 0.0f) {
        const float disc = fmaf(a[2], a[2], -3.0f * a[3] * a[1]);
        if (disc >= 0.0f) {
            const float sq = sqrtf(disc);
            r[0] = (-a[2] - sq) / (3.0f * a[3]); r[1] = (-a[2] + sq) / (3.0f * a[3]);
            nr = 2;
        }
    } else if (a[2] != 0.0f) {
        r[0] = -a[1] / (2.0f * a[2]);
        nr = 1;
    }
    for (int k = 0; k < nr; k++)
        if (r[k] >= lo && r[k] <= hi) {
            mono = false;
            const float v = val(r[k]);
            pmin = fminf(pmin, v); pmax = fmaxf(pmax, v);
        }
    if (!(pmin == pmin) || !(pmax == pmax)) { mono = false; pmin = -3.0e38f; pmax = 3.0e38f; }
    return mono;
}
// worst sample error [mm] -> margin; a sample that is not finite or an error above 5 cm switches the predictor off
SOLAT_HD void ff_pred_set_margin(FfPred& p, float worst_mm)
{
    p.on = p.on && worst_mm < 50.0f;     // NaN compares false
    p.margin_cm = 0.1f * (5.0f * worst_mm + 2.0f);
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

// sin, cos for |x| < ~1e5 rad (near-field phases k L ~ 1e4): FP64 Cody-Waite reduction by pi/2 (fdlibm's 33-bit
// split) and the FP64 leading terms (r - r^3/6 + r^5/120, 1 - r^2/2 + r^4/24 - r^6/720); the remaining tails
// (|.| < 4e-5) run in FP32 on the FMA pipe.  Absolute error ~1e-11 (tests: test_fast_math_primitives).
SOLAT_HD void sincos_mid(double x, double& sn, double& cs)
{
    const double TWO_OVER_PI = 0.6366197723675814, MAGIC = 6755399441055744.0;        // 1.5 * 2^52
    const double P1 = 1.57079632673412561417e+00, P2 = 6.07710050630396597660e-11, P3 = 2.02226624879595063154e-21;
    const double nb = fma(x, TWO_OVER_PI, MAGIC);        // round-to-nearest integer lands in the low mantissa bits
    const double n = nb - MAGIC;
    double r = fma(-n, P1, x);
    r = fma(-n, P2, r);
    r = fma(-n, P3, r);
#if defined(__CUDA_ARCH__)
    const int q = __double2loint(nb);
#else
    long long bits;
    memcpy(&bits, &nb, 8);
    const int q = (int)(bits & 0xffffffff);
#endif
    const double r2 = r * r;
    const float rf = (float)r, r2f = rf * rf, r4f = r2f * r2f;
    // tails: sin: r^7 (-1/5040 + r^2/362880 - r^4/39916800);  cos: r^8 (1/40320 - r^2/3628800 + r^4/479001600)
    const float ts = rf * r2f * r4f * fmaf(r2f, fmaf(r2f, -2.5052108e-8f, 2.7557319e-6f), -1.9841270e-4f);
    const float tc = r4f * r4f * fmaf(r2f, fmaf(r2f, 2.0876757e-9f, -2.7557319e-7f), 2.4801587e-5f);
    const double s0 = fma(r * r2, fma(r2, 1.0 / 120.0, -1.0 / 6.0), r) + (double)ts;
    const double c0 = fma(r2, fma(r2, fma(r2, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0) + (double)tc;
    const double s1 = (q & 1) ? c0 : s0, c1 = (q & 1) ? s0 : c0;
    sn = (q & 2) ? -s1 : s1;
    cs = ((q + 1) & 2) ? -c1 : c1;
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
// source_to_lyot_model rows, rti:1346-1373.  rows[26]
SOLAT_HD void pack_nf_source_rows(const NfSourceRay& r, const vec3& p0, double n_si, double n_vac, double* rows)
{
    if (!r.valid) { for (int i = 0; i < 26; i++) rows[i] = 0.0; return; }     // rti:1086-1114
    // rti:1305-1339: Snell (si -> vac) applied to the vector pointing from the Lyot hit back to 2A
    const vec3 og = refract_flat(neg(r.d_out), 1.0, n_si / n_vac);
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

// np.searchsorted(a, v, 'left') from an interpolated first guess (single precision is enough for a guess), then an
// exact fix-up against the array itself: one or two loads on a uniform grid, still exact (only slower) on any
// sorted grid.
SOLAT_HD int searchsorted_guess(const double* a, int n, int stride, double v, double a0, float inv_step)
{
    const float g = ceilf((float)(v - a0) * inv_step);
    int i = g > 0.0f ? (g < (float)n ? (int)g : n) : 0;
    while (i > 0 && !(a[(long)(i - 1) * stride] < v)) --i;
    while (i < n && a[(long)i * stride] < v) ++i;
    return i;
}

// The same on an axis the host has found sorted and uniform to within a quarter of its step (axis_is_lean): the
// interpolated guess is then within one of the answer, and one comparison against each neighbour a[i-1], a[i] -- the
// staged axis carries -inf at a[-1] and +inf at a[n] -- settles it without a branch or a loop.
SOLAT_HD int searchsorted_lean(const double* a, float n, double v, double a0, float inv_step)
{
    const float g = ceilf((float)(v - a0) * inv_step);
    const int i = (int)fminf(fmaxf(g, 0.0f), n);
    const double lo = a[i - 1], hi = a[i];
    return i + (hi < v ? 1 : 0) - (lo >= v ? 1 : 0);
}
// Host-side condition for searchsorted_lean: strictly increasing and |a[i] - (a[0] + i step)| < step / 4, so that the
// single-precision guess (relative error 2e-7 on an index below 2^12) cannot miss by more than one.
inline double axis_deviation(const double* a, int n, long stride)   // max |a[i] - (a[0] + i step)| / step
{
    if (n < 2) return 1.0;
    const double a0 = a[0], step = (a[(long)(n - 1) * stride] - a0) / (n - 1);
    if (!(step > 0.0)) return 1.0;
    double d = 0.0;
    for (int i = 0; i < n; i++) d = fmax(d, fabs(a[(long)i * stride] - (a0 + i * step)) / step);
    return d;
}
inline bool axis_is_lean(const double* a, int n, long stride)
{
    if (n < 2 || n > 4096) return false;
    const double a0 = a[0], step = (a[(long)(n - 1) * stride] - a0) / (n - 1);
    if (!(step > 0.0) || !(step < 1e300)) return false;
    for (int i = 0; i < n; i++) {
        const double v = a[(long)i * stride];
        if (i && !(v > a[(long)(i - 1) * stride])) return false;
        if (!(fabs(v - (a0 + i * step)) < 0.25 * step)) return false;
    }
    return true;
}

// x / 10 correctly rounded without the division subroutine: q = RN(x * 0.1) is within 1 ulp, the FMA
// residual r = x - 10 q is exact, and RN(q + r * 0.1) is the correctly rounded quotient (Markstein).
SOLAT_HD double div10(double x)
{
    const double q = x * 0.1;
    return fma(fma(-q, 10.0, x), 0.1, q);
}

// far_field_latrt.py:14-26 for one ray: focal-plane offsets (mm) -> cell index iy*M+ix, or -1 when
// the reference skips the ray.  gx = data_x[0,:] (M), gy = data_y (M x M row-major).
struct FfGrid {
    const double* gx;      // data_x[0, :]
    const double* gy;      // data_y, (M, M) row-major; column ix is gy[ix * gy_cs + i * gy_rs]
    int M, gy_cs, gy_rs;   // (1, M) for the full array; (0, 1) when every column is the same and gy holds just one
    int y_uniform;         // host: all columns of data_y are identical (a kernel may then stage one column on chip)
    double gx_max, gy_max; // ffl:19: rays beyond these are skipped
    double gx_min, gy_min; // min(data_x[0, :]), min(data_y[:, 0]): rays below both pile into cell (0, 0)  (ffl:22-23)
    float inv_dx, inv_dy;  // (M-1)/(last-first) of gx and of gy[:,0]; 0 when not usable as a guess
    int lean;              // host: both axes pass axis_is_lean and all columns of data_y agree (searchsorted_lean applies)
    float dev_u;           // host: largest deviation of either axis from a[0] + i step, in steps (ff_map_interp_kernel)
};
SOLAT_HD int ff_cell(double fx_mm, double fz_mm, const FfGrid& G)
{
    const double xx = div10(fx_mm), yy = div10(fz_mm);                          // ffl:12-13: / 1e1
    if (G.gx_max < xx || G.gy_max < yy) return -1;
    if (not_finite(xx) || not_finite(yy)) return -1;       // a NaN ray never reaches a cell (flagged ST_NAN upstream)
    const int ix = searchsorted_guess(G.gx, G.M, 1, xx, G.gx[0], G.inv_dx);     // ffl:22
    if (ix >= G.M) return -1;                              // only possible on an unsorted grid
    const double* col = G.gy + ix * G.gy_cs;
    const int iy = searchsorted_guess(col, G.M, G.gy_rs, yy, col[0], G.inv_dy);      // ffl:23
    if (iy >= G.M) return -1;
    return iy * G.M + ix;
}

// np.searchsorted(a, v, 'left') on a uniform axis from u = (v - a[0]) / step alone, plus how sure the answer is.
// Returns ceil(u) -- the number of elements below v -- of u clamped to [-1/4, u_last + 3/4]: 0 at and below the first
// element (searchsorted returns 0 there), M = u_last + 1 beyond the last one (ffl:19 drops the ray); `sure`: u is more
// than tol away from every integer of [0, u_last], i.e. v more than eps away from every element (the clamp bounds sit
// a quarter step away from one, so a value beyond them is sure).  NaN is the caller's business.  Branch-free.  Single
// precision is enough: |u| < 2^9, tol ~ 3.5e-3, of which 5e-4 are slack for the roundings (lattice values and weights
// in single precision) and the axis' measured deviation from uniform (FfGrid::dev_u).
SOLAT_HD int ff_axis_cell(float u, float u_hi, float tol, bool& sure)
{
    const float uc = fminf(fmaxf(u, -0.25f), u_hi);
    const float r = rintf(uc), t = uc - r;
    sure = fabsf(t) > tol;
    return (int)r + (t > 0.0f ? 1 : 0);
}
// A ray's cell in one winner table from its predicted position: >= 0 the cell, surely; -1 surely dropped (beyond the
// last element of either axis); -2 within tol of an edge that matters.
SOLAT_HD int ff_cell_class(float ux, float uy, float u_hi, float tol_x, float tol_y, int M)
{
    bool sx, sy;
    const int cx = ff_axis_cell(ux, u_hi, tol_x, sx), cy = ff_axis_cell(uy, u_hi, tol_y, sy);
    if ((cx == M && sx) || (cy == M && sy)) return -1;
    return (sx && sy) ? cy * M + cx : -2;
}

// Last-write-wins table entry (ffl:28: a later ray overwrites an earlier one): the ray index in the top bits
// so that a 64-bit atomicMax keeps the highest index, and the ray's path length below it as an exact
// fixed-point offset from a per-plan base (unit = half an ulp of the base).  `shift` = bits of the offset: 44 for
// fans up to 2^20 - 1 rays (+-65 m at 1e8 mm), 40 for fans up to 2^24 - 1 rays (+-4 m); a path length outside the
// range is reported (ST_RANGE), never silently wrapped.
constexpr int FF_KEY_SHIFT = 44;
constexpr int FF_KEY_SHIFT_WIDE = 40;
constexpr int FF_MAX_RAYS = (1 << 20) - 1;
constexpr int FF_MAX_RAYS_WIDE = (1 << 24) - 1;
SOLAT_HD unsigned long long ff_key(int ray, double opl, double base, double inv_unit, bool& exact, double unit, int shift = FF_KEY_SHIFT)
{
    const double q = (opl - base) * inv_unit;
    const long long qi = (long long)rint(q), half = 1ll << (shift - 1);
    exact = (base + (double)qi * unit == opl) && (qi > -half) && (qi < half);
    return ((unsigned long long)(ray + 1) << shift) | (unsigned long long)(qi + half);
}
SOLAT_HD double ff_key_opl(unsigned long long key, double base, double unit, int shift = FF_KEY_SHIFT)
{
    const long long qi = (long long)(key & ((1ull << shift) - 1)) - (1ll << (shift - 1));
    return base + (double)qi * unit;
}

// Which rays of one phi row [lo, hi) does the map have to trace?  A ray is of use to the pixel itself (`pos`) unless
// the predictor puts it beyond the high edge of the E_rec grid (ffl:19 skips it) or below BOTH low edges: such a ray
// lands in cell (0, 0) for certain (ffl:22-23: searchsorted returns 0 below the first element), and of all the rays
// of the row that do, only the last one -- the highest ray index -- can hold the cell (ffl:28).  The same for the
// mirror pixel (`neg`: focal x negated, z shifted by dz_neg_cm, and its ray index grows with i as well).  Kept: the
// runs of useful rays plus those two corner candidates, as at most FF_ROW_K intervals (more: the whole row is kept).
// All thresholds carry the predictor's margin.  Along a row the predicted x and z are cubics in s = sin(theta_i)
// (ff_pred_row_cubic), dominated by their linear term; where a cubic is monotone over the row's accepted rays, or its
// whole range stays on one side of a threshold (ff_pred_row_range; otherwise the whole row is kept), the comparison
// changes its answer at most once along the row: the change points are found by bisection and the row is classified
// segment by segment, with the very same expressions.
constexpr int FF_ROW_K = 4;
// Pile cells.  A ray below a low edge of the E_rec grid lands in column 0 (x <= gx[0]) or row 0 of a winner table
// (ffl:22-23) and only the latest ray of each such cell counts (ffl:28): on a +-5 cm grid more than half of the listed
// rays are "pile" rays and all but a few hundred lose.  Before the map is traced, the set-up kernel traces a few
// hundred SCOUTS per work item -- per half cell of each pile line the latest ray the predictor places there -- and
// records, per pile cell, the latest scout that really landed in it (L, exact).  A pile ray whose every reachable cell
// (predicted position +- margin) already holds a later scout cannot win and is dropped from the list
// (ff_pred_row_keep with `pile`).  Exact by construction: a ray is dropped only if a later ray THAT STAYS LISTED sits
// in each cell it could reach (L only holds scouts that are the latest in a cell of their own line, and such a scout's
// row segment always keeps it).
struct FfPile {
    const int* L;          // [2 tables][2 lines][M]: (ray index in the table's own numbering) + 1 of the latest surviving scout in
                           // cell (c, 0) [line 0: column 0, c = iz] / (0, c) [line 1: row 0, c = ix]; 0 = none
    int nth, nph, j;       // fan, and the phi row the call is about
};
constexpr int FF_PILE_MIN = 20000;     // listed rays per work item from which the scouts pay (a 100 x 100 fan lists ~3000)
constexpr int FF_SCOUT_STEP = 4;       // every n-th listed ray of a row interval (from its last ray down) is offered as a scout
// half-cell slot [0, 2M) of a pile coordinate c [cm] on pile line `line` (0: the y axis, 1: the x axis).  Only a
// bucket for choosing scouts: arithmetic on a uniform axis (first-guess step known), binary search otherwise.
SOLAT_HD int ff_pile_slot(const FfGrid& G, int line, float c)
{
    const int M = G.M;
    const float inv = line ? G.inv_dx : G.inv_dy;
    if (inv > 0.0f) {
        const float h = floorf((c - (float)(line ? G.gx_min : G.gy_min)) * inv * 2.0f) + 1.0f;
        return h < 0.0f ? 0 : (h > (float)(2 * M - 1) ? 2 * M - 1 : (int)h);
    }
    const double* g = line ? G.gx : G.gy;
    const int gs = line ? 1 : G.gy_rs;
    int k = searchsorted_left(g, M, gs, (double)c);
    if (k > M - 1) k = M - 1;
    const bool upper = k > 0 && (double)c > 0.5 * (g[(long)(k - 1) * gs] + g[(long)k * gs]);
    return 2 * k + (upper ? 1 : 0);
}
SOLAT_HD int ff_pred_row_keep(const FfPred& p, const float* stf, float cpf, float spf, const FfGrid& G, bool need_pos, bool need_neg,
                              float dz_neg_cm, int lo, int hi, int* out_lo, int* out_cnt, const FfPile* pile = nullptr)
{
    for (int k = 0; k < FF_ROW_K; k++) { out_lo[k] = lo; out_cnt[k] = 0; }
    if (hi <= lo) return 0;
    float ax[4], az[4];
    ff_pred_row_cubic(p.cx, cpf, spf, ax);
    ff_pred_row_cubic(p.cz, cpf, spf, az);
    const float m = p.margin_cm;
    const float s_a = stf[lo] * FF_PRED_SCALE, s_b = stf[hi - 1] * FF_PRED_SCALE;
    float xmin, xmax, zmin, zmax;
    const bool x_mono = ff_pred_row_range(ax, s_a, s_b, xmin, xmax), z_mono = ff_pred_row_range(az, s_a, s_b, zmin, zmax);
    const float x_hi = (float)G.gx_max + m, z_hi = (float)G.gy_max + m, x_lo = (float)G.gx_min - m, z_lo = (float)G.gy_min - m;
    const float ez = fabsf(dz_neg_cm);
    {   // The bisections below need every comparison to change its answer at most once along the row: true where the
        // cubic is monotone, or where its whole range (0.1 mm slack) lies on one side of the threshold.  Otherwise the
        // row is kept whole.
        auto clear = [](bool mono, float pmin, float pmax, float t) { return mono || t > pmax + 0.01f || t < pmin - 0.01f; };
        const float x0 = 0.1f * xmin, x1 = 0.1f * xmax, z0 = 0.1f * zmin, z1 = 0.1f * zmax;     // cm
        const bool ok = clear(x_mono, x0, x1, x_hi) && clear(x_mono, x0, x1, -x_hi) && clear(x_mono, x0, x1, x_lo) &&
                        clear(x_mono, x0, x1, -x_lo) && clear(z_mono, z0, z1, z_hi + ez) && clear(z_mono, z0, z1, z_lo - ez);
        if (!ok) {
            out_cnt[0] = hi - lo;
            return 1;
        }
    }
    // the six comparisons of ray i (NaN predictions compare false everywhere: the ray counts as useful)
    auto test = [&](int i, int which) {
        const float s = stf[i] * FF_PRED_SCALE;
        const float x = 0.1f * fmaf(fmaf(fmaf(ax[3], s, ax[2]), s, ax[1]), s, ax[0]);
        const float z = 0.1f * fmaf(fmaf(fmaf(az[3], s, az[2]), s, az[1]), s, az[0]);
        switch (which) {
            case 0: return x > x_hi;
            case 1: return -x > x_hi;
            case 2: return z > z_hi + ez;
            case 3: return x < x_lo;
            case 4: return -x < x_lo;
            default: return z < z_lo - ez;
        }
    };
    // change points: first index whose answer differs from that of `lo` (hi if none)
    int cut[8];
    cut[0] = lo; cut[7] = hi;
    for (int w = 0; w < 6; w++) {
        const bool first = test(lo, w);
        int a = lo, b = hi;                       // answer(a) == first; b = hi or answer(b) != first
        if (test(hi - 1, w) == first) { cut[1 + w] = hi; continue; }
        b = hi - 1;
        while (a + 1 < b) {
            const int mid = (a + b) >> 1;
            if (test(mid, w) == first) a = mid; else b = mid;
        }
        cut[1 + w] = b;
    }
    for (int a = 1; a < 8; a++)                   // sort the eight cut points
        for (int b = a; b > 0 && cut[b] < cut[b - 1]; b--) { const int t = cut[b]; cut[b] = cut[b - 1]; cut[b - 1] = t; }
    // bounding sub-range [a, b) of the rays of the pile-only segment [i0, i1) that may still win a cell of pile line
    // `line` of table t (see FfPile): the cells the segment can reach are scanned for "open" ones (no scout later than
    // the segment's last ray), and the rays whose +-margin window touches the span of the open cells are kept
    auto shrink = [&](int i0, int i1, int t, int line, int& a, int& b) {
        const float eps = 1e-3f;
        auto coord = [&](int i) {
            const float s = stf[i] * FF_PRED_SCALE;
            const float x = 0.1f * fmaf(fmaf(fmaf(ax[3], s, ax[2]), s, ax[1]), s, ax[0]);
            const float z = 0.1f * fmaf(fmaf(fmaf(az[3], s, az[2]), s, az[1]), s, az[0]);
            return line == 0 ? (t ? z + dz_neg_cm : z) : (t ? -x : x);
        };
        const double* g = line == 0 ? G.gy : G.gx;
        const int gs = line == 0 ? G.gy_rs : 1, M = G.M;
        const float c0 = coord(i0), c1 = coord(i1 - 1);
        const float cmin = fminf(c0, c1) - m - eps, cmax = fmaxf(c0, c1) + m + eps;
        a = b = i0;
        if (!(cmin <= (float)g[(long)(M - 1) * gs]) || !(c0 == c0) || !(c1 == c1)) {
            if (!(c0 == c0) || !(c1 == c1)) { a = i0; b = i1; }                 // NaN prediction: keep
            return;                                                           // all beyond the last value: no cell at all
        }
        const int ka = searchsorted_left(g, M, gs, (double)cmin);
        int kb = searchsorted_left(g, M, gs, (double)cmax);
        if (kb > M - 1) kb = M - 1;
        const int* L = pile->L + (2 * t + line) * M;
        const int top = (t ? pile->nph - 1 - pile->j : pile->j) * pile->nth + i1;     // (index of the segment's last ray) + 1
        int k_lo = -1, k_hi = -1;
        for (int k = ka; k <= kb; k++)
            if (L[k] <= top) { if (k_lo < 0) k_lo = k; k_hi = k; }            // open: its scout is not later than the segment
        if (k_lo < 0) return;
        const float lo_edge = k_lo == 0 ? -3.0e38f : (float)g[(long)(k_lo - 1) * gs] - m - eps, hi_edge = (float)g[(long)k_hi * gs] + m + eps;
        // coord is monotone along the segment: the rays with lo_edge < coord <= hi_edge form one range
        const bool inc = c1 >= c0;
        auto first_true = [&](auto&& pred) {                                  // pred: false ... false true ... true on [i0, i1)
            int u = i0, v = i1;                                               // answer in [u, v]
            while (u < v) {
                const int mid = (u + v) >> 1;
                if (pred(mid)) v = mid; else u = mid + 1;
            }
            return u;
        };
        if (inc) {
            a = first_true([&](int i) { return coord(i) > lo_edge; });
            b = first_true([&](int i) { return coord(i) > hi_edge; });
        } else {
            a = first_true([&](int i) { return coord(i) <= hi_edge; });
            b = first_true([&](int i) { return coord(i) <= lo_edge; });
        }
        if (b < a) b = a;
    };
    int n = 0, run = -1, run_end = -1, cand_pos = -1, cand_neg = -1;
    bool overflow = false;
    auto flush_run = [&]() {
        if (run >= 0) {
            if (n < FF_ROW_K) { out_lo[n] = run; out_cnt[n] = run_end - run; }
            ++n; run = -1;
        }
    };
    for (int sgm = 0; sgm < 7; sgm++) {
        const int i0 = cut[sgm], i1 = cut[sgm + 1];
        if (i1 <= i0) continue;
        const bool beyond_z = test(i0, 2), below = test(i0, 5);
        const bool left_p = test(i0, 3), left_n = test(i0, 4);
        const bool pos_corner = need_pos && left_p && below, neg_corner = need_neg && left_n && below;
        const bool pos_use = need_pos && !test(i0, 0) && !beyond_z && !pos_corner;
        const bool neg_use = need_neg && !test(i0, 1) && !beyond_z && !neg_corner;
        if (pos_corner) cand_pos = i1 - 1;
        if (neg_corner) cand_neg = i1 - 1;
        int pa = i0, pb = (pos_use || neg_use) ? i1 : i0;                     // the piece of this segment that is kept
        if (pile && pb > pa) {
            const bool pos_in = pos_use && !left_p && !below, neg_in = neg_use && !left_n && !below;
            if (!pos_in && !neg_in) {                                         // pile rays only: keep what may still win
                int ua = i1, ub = i0;
                for (int t = 0; t < 2; t++) {
                    if (!(t ? neg_use : pos_use)) continue;
                    int a, b;
                    shrink(i0, i1, t, (t ? left_n : left_p) ? 0 : 1, a, b);
                    if (b > a) { ua = a < ua ? a : ua; ub = b > ub ? b : ub; }
                }
                pa = ua; pb = ub;
                if (pb <= pa) { pa = i0; pb = i0; }
            }
        }
        if (pb > pa) {
            if (run >= 0 && pa == run_end) run_end = pb;                      // joins the run
            else { flush_run(); run = pa; run_end = pb; }
        } else {
            flush_run();
        }
    }
    flush_run();
    // the corner candidates, unless a kept run holds them already
    for (int c = 0; c < 2; c++) {
        const int i = c ? cand_neg : cand_pos;
        if (i < 0 || (c == 1 && cand_neg == cand_pos)) continue;
        bool inside = false;
        for (int k = 0; k < n && k < FF_ROW_K; k++) inside = inside || (i >= out_lo[k] && i < out_lo[k] + out_cnt[k]);
        if (inside) continue;
        if (n < FF_ROW_K) { out_lo[n] = i; out_cnt[n] = 1; }
        ++n;
    }
    overflow = n > FF_ROW_K;
    if (overflow && pile)                      // the shrunk pieces do not fit: the row as it is without pile information
        return ff_pred_row_keep(p, stf, cpf, spf, G, need_pos, need_neg, dz_neg_cm, lo, hi, out_lo, out_cnt, nullptr);
    if (overflow) {
        out_lo[0] = lo; out_cnt[0] = hi - lo;
        for (int k = 1; k < FF_ROW_K; k++) { out_lo[k] = lo; out_cnt[k] = 0; }
        return 1;
    }
    // ascending order of the intervals (the walk's row cursor only moves one way): insertion sort of <= 4 entries
    for (int a = 1; a < n; a++)
        for (int b = a; b > 0 && out_lo[b] < out_lo[b - 1]; b--) {
            const int tl = out_lo[b], tc = out_cnt[b];
            out_lo[b] = out_lo[b - 1]; out_cnt[b] = out_cnt[b - 1];
            out_lo[b - 1] = tl; out_cnt[b - 1] = tc;
        }
    return n;
}


// ------------------------------------------------------------------------------------------
// regrid (ray_trace_int.py:1385-1440): one line of the separable interpolating bicubic spline.
// y: n samples at unit spacing (stride ystride); M, cp: scratch of n doubles each; writes n_new samples of the
// not-a-knot cubic spline at u_j = j (n-1)/(n_new-1).
// ------------------------------------------------------------------------------------------
SOLAT_HD void spline_line(const double* y, int ystride, int n, double* M, double* cp, int n_new, double* out, int ostride)
{
    // uniform unit spacing in old-index units; not-a-knot: M1 = d1, M(n-2) = d(n-2), d_i = y(i-1) - 2 y(i) + y(i+1)
    if (n < 4) {   // parabola through 3 points / straight line through 2 / constant
        for (int j = 0; j < n_new; j++) {
            const double u = n_new > 1 ? (double)j * (double)(n - 1) / (double)(n_new - 1) : 0.0;
            double v;
            if (n == 1) v = y[0];
            else if (n == 2) v = y[0] + (y[ystride] - y[0]) * u;
            else {
                const double y0 = y[0], y1 = y[ystride], y2 = y[2 * ystride];
                v = y0 * (u - 1) * (u - 2) * 0.5 - y1 * u * (u - 2) + y2 * u * (u - 1) * 0.5;
            }
            out[(size_t)j * ostride] = v;
        }
        return;
    }
#define SOLAT_SPL_Y(i) (y[(size_t)(i) * ystride])
#define SOLAT_SPL_D(i) (SOLAT_SPL_Y((i) - 1) - 2.0 * SOLAT_SPL_Y(i) + SOLAT_SPL_Y((i) + 1))
    M[1] = SOLAT_SPL_D(1);
    M[n - 2] = SOLAT_SPL_D(n - 2);
    if (n > 4) {
        // M(i-1) + 4 M(i) + M(i+1) = 6 d(i),  i = 2 .. n-3, with M(1) and M(n-2) known
        const int lo = 2, hi = n - 3;
        double bprev = 4.0, rprev = 6.0 * SOLAT_SPL_D(lo) - M[1];
        if (lo == hi) rprev -= M[n - 2];
        cp[lo] = 1.0 / bprev;             // c'(lo)
        M[lo] = rprev / bprev;            // d'(lo)
        for (int i = lo + 1; i <= hi; i++) {
            double r = 6.0 * SOLAT_SPL_D(i);
            if (i == hi) r -= M[n - 2];
            const double b = 4.0 - cp[i - 1];
            cp[i] = 1.0 / b;
            M[i] = (r - M[i - 1]) / b;
        }
        for (int i = hi - 1; i >= lo; i--) M[i] -= cp[i] * M[i + 1];
    }
    M[0] = 2.0 * M[1] - M[2];
    M[n - 1] = 2.0 * M[n - 2] - M[n - 3];
    for (int j = 0; j < n_new; j++) {
        const double u = n_new > 1 ? (double)j * (double)(n - 1) / (double)(n_new - 1) : 0.0;
        int i = (int)u;
        if (i > n - 2) i = n - 2;
        const double t = u - i, w = 1.0 - t;
        out[(size_t)j * ostride] = w * SOLAT_SPL_Y(i) + t * SOLAT_SPL_Y(i + 1) + ((w * w * w - w) * M[i] + (t * t * t - t) * M[i + 1]) / 6.0;
    }
#undef SOLAT_SPL_Y
#undef SOLAT_SPL_D
}

}  // namespace solat
