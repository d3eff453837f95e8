/*
 * solat_oracle.c  --  CPU restatement (plain C, FP64) of the reference's beam-simulation
 * hot path.  TEST INFRASTRUCTURE ONLY: it is the checker for the CUDA path and the
 * "port" CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product package never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
 * fixtures produced by the unmodified Python reference (oracle/gen_golden.py ->
 * tests/golden/ *.npz).  Far-field panel mask / panel errors come from the un-vendored
 * `pan_mod` (holosim-ml, unpinned); those are pinned only against the documented
 * stand-in (oracle/ref_shims/pan_mod.py) -- "parity unpinned" for the real pan_mod.
 *
 * The algorithm follows the reference structurally: one bracketed root-find per surface
 * per ray (same brackets and tolerances as scipy.optimize.brentq's defaults), the same
 * frame transforms, the same vector Snell formula, the same path-length bookkeeping.
 * Citations are file:line under /root/reference/.
 *
 * Third-party arithmetic restated here: scipy.optimize.brentq (requirements.txt pins
 * scipy==1.5.4; Brent's method as published in scipy/optimize/Zeros/brentq.c, defaults
 * xtol=2e-12, rtol=4*DBL_EPSILON, maxiter=100).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ brentq */
typedef double (*root_fn)(double t, void *ctx);

#define ORC_OK 0
#define ORC_ERR_BRACKET -1 /* reference: ValueError "f(a) and f(b) must have different signs" */
#define ORC_ERR_CONV -2

static double brentq(root_fn f, double xa, double xb, void *ctx, int *status)
{
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    const int maxiter = 100;
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0.;
    double sbis, delta, stry, dpre, dblk;
    int i;
    fpre = f(xpre, ctx);
    fcur = f(xcur, ctx);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) {
        *status = ORC_ERR_BRACKET;
        return NAN;
    }
    for (i = 0; i < maxiter; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double lim1 = fabs(spre), lim2 = 3 * fabs(sbis) - delta;
            if (2 * fabs(stry) < (lim1 < lim2 ? lim1 : lim2)) {
                spre = scur; scur = stry;
            } else {
                spre = sbis; scur = sbis;
            }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur, ctx);
    }
    *status = ORC_ERR_CONV;
    return xcur;
}

/* ------------------------------------------------------------------ pthread parallel-for */
typedef void (*par_body)(int i, void *ctx);
typedef struct { par_body body; void *ctx; int n; atomic_int next; } par_job;
static void *par_worker(void *v)
{
    par_job *j = (par_job *)v;
    for (;;) {
        int i = atomic_fetch_add(&j->next, 1);
        if (i >= j->n) break;
        j->body(i, j->ctx);
    }
    return NULL;
}
static int par_default_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
static void par_for(int n, int nthreads, par_body body, void *ctx)
{
    if (nthreads <= 0) nthreads = par_default_threads();
    if (nthreads > n) nthreads = n;
    par_job j;
    j.body = body; j.ctx = ctx; j.n = n;
    atomic_init(&j.next, 0);
    if (nthreads <= 1) { par_worker(&j); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++)
        if (pthread_create(&th[started], NULL, par_worker, &j) == 0) started++;
    par_worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------ small vectors */
typedef struct { double x, y, z; } v3;
static v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vscale(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
static double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 vcross(v3 a, v3 b)
{
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static v3 along(v3 p, v3 d, double t) { return V(p.x + d.x * t, p.y + d.y * t, p.z + d.z * t); }

/* ray_trace_int.py:20-30  snell_vec: mu*N x (-N x s) - N*sqrt(1 - mu^2 |N x s|^2) */
static v3 snell_vec(double n1, double n2, v3 N, v3 s)
{
    double mu = n1 / n2;
    v3 c1 = vcross(N, vcross(vneg(N), s));
    v3 c2 = vcross(N, s);
    double root = sqrt(1 - mu * mu * vdot(c2, c2)); /* NaN on total internal reflection */
    return vsub(vscale(c1, mu), vscale(N, root));
}

/* ------------------------------------------------------------------ near-field geometry
 * ot_geo.py:61-71 tilt angles and lens offsets; :81-100,193-212,297-338 sags;
 * :112-139,224-250,348-405 gradients; :162-183,273-294,474-495,522-531 frames.       */
typedef struct { double R, k, a1, a2, a3, a4; } asph_t;
static const asph_t ASPH_1A = {-1510.61, 12.0966, -5.876941e-5, 1.769198e-9, -1.310183e-15, 1.994766e-19};
static const asph_t ASPH_2A = {-871.25, -10.201256, -1.147640e-04, 1.103643e-10, -2.969608e-14, 5.980271e-19};
static const asph_t ASPH_3A = {6608.30, 414.142695, -2.981594e-06, 3.955202e-10, 3.180727e-16, 0};
static const asph_t ASPH_3B = {604.44, -3.177341, 2.366829e-05, -1.000417e-10, -6.565681e-15, 0};

#define HALF_PI 1.5707963267948966
static const double LENS1_Y = 975.08, LENS2_Y = 592.76, LYOT_Y = 453.62;
#define LENS3_Y (180.8 - ((2.5 + 1.513) - (35.16 - 32.19)))

typedef struct {
    double y0;          /* translation applied to y before the tilt */
    double th;          /* tilt; frame uses rotation by -th about x */
    const asph_t *asph; /* NULL for a flat */
} nf_surf;

enum { S_1B, S_1A, S_2B, S_2A, S_3A, S_3B, S_LYOT, S_COUNT };
static nf_surf nf_surface(int id)
{
    nf_surf s;
    s.asph = NULL;
    switch (id) {
    case S_1B: s.y0 = LENS1_Y; s.th = HALF_PI; break;                          /* th1_l1 */
    case S_1A: s.y0 = LENS1_Y - 19.32; s.th = -HALF_PI; s.asph = &ASPH_1A; break; /* th2_l1 */
    case S_2B: s.y0 = LENS2_Y; s.th = -HALF_PI; break;                         /* th1_l2 */
    case S_2A: s.y0 = LENS2_Y - 26.47; s.th = -HALF_PI; s.asph = &ASPH_2A; break; /* uses th2_l1, ot_geo.py:291 */
    case S_3A: s.y0 = LENS3_Y; s.th = -HALF_PI; s.asph = &ASPH_3A; break;      /* th1_l3 */
    case S_3B: s.y0 = LENS3_Y + 35.16; s.th = -HALF_PI; s.asph = &ASPH_3B; break; /* th2_l3 */
    default: s.y0 = LYOT_Y; s.th = -HALF_PI; break;                            /* th1_l3 */
    }
    return s;
}

static v3 tele_into(const nf_surf *s, v3 p)
{
    double y = p.y + s->y0, c = cos(-s->th), sn = sin(-s->th);
    return V(p.x, y * c - p.z * sn, y * sn + p.z * c);
}
static v3 rot_back(double th, v3 v)
{
    double c = cos(th), sn = sin(th);
    return V(v.x, v.y * c - v.z * sn, v.y * sn + v.z * c);
}
static double asph_sag(const asph_t *a, double x, double y)
{
    double c = 1 / a->R, r = sqrt(x * x + y * y), r2 = r * r;
    double z = (c * r2) / (1 + sqrt(1 - ((1 + a->k) * c * c * r2)));
    z += a->a1 * r2 + a->a2 * pow(r, 4) + a->a3 * pow(r, 6) + a->a4 * pow(r, 8);
    return z;
}
static void asph_grad(const asph_t *a, double x, double y, double *gx, double *gy)
{
    double c = 1 / a->R, r = sqrt(x * x + y * y), r2 = r * r;
    double S = sqrt(1 - ((1 + a->k) * c * c * r2));
    double c1 = a->a1 * 2 + a->a2 * 4 * r2 + a->a3 * 6 * pow(r, 4) + a->a4 * 8 * pow(r, 6);
    double c2 = (c * 2) / (1 + S);
    double c3 = (c * c * c * (a->k + 1) * r2) / (S * (1 + S) * (1 + S));
    *gx = x * (c1 + c2 + c3);
    *gy = y * (c1 + c2 + c3);
}

typedef struct { v3 p, d; const nf_surf *s; } nf_root_ctx;
static double nf_root(double t, void *vctx)
{
    nf_root_ctx *c = (nf_root_ctx *)vctx;
    v3 l = tele_into(c->s, along(c->p, c->d, t));
    double sag = 0;
    if (c->s->asph) {
        sag = asph_sag(c->s->asph, l.x, l.y);
        if (isnan(sag)) sag = 0; /* ray_trace_int.py:648-649 */
    }
    return l.z - sag;
}

/* One refracting interface.  Returns brentq parameter in *t, hit point in *hit, new
 * tele-frame direction in *dir_out, Euclidean leg length in *dist.  nsign multiplies the
 * normal handed to snell_vec; ssign the incoming tangent (toy trace, rti:465).            */
typedef struct { double t, dist; v3 hit, n_local, tan_local, og_local, og_tele; } nf_leg;
static int nf_interface(int surf_id, v3 p, v3 d, double lo, double hi, double n1, double n2,
                        double nsign, double ssign, nf_leg *leg)
{
    nf_surf s = nf_surface(surf_id);
    nf_root_ctx ctx = {p, d, &s};
    int st = ORC_OK;
    double t = brentq(nf_root, lo, hi, &ctx, &st);
    if (st != ORC_OK) return st;
    v3 hit = along(p, d, t);
    v3 hl = tele_into(&s, hit), pl = tele_into(&s, p);
    double gx = 0, gy = 0;
    if (s.asph) asph_grad(s.asph, hl.x, hl.y, &gx, &gy);
    v3 nt = V(-gx, -gy, 1);
    v3 n = vscale(nt, 1 / sqrt(vdot(nt, nt)));
    v3 vec = vsub(hl, pl);
    double dist = sqrt(vdot(vec, vec));
    v3 tan = vscale(vec, 1 / dist);
    v3 og = snell_vec(n1, n2, vscale(n, nsign), vscale(tan, ssign));
    leg->t = t; leg->dist = dist; leg->hit = hit; leg->n_local = n; leg->tan_local = tan;
    leg->og_local = og; leg->og_tele = rot_back(s.th, og);
    return ORC_OK;
}

static void linspace(double a, double b, int n, double *out)
{
    /* numpy.linspace: start + i*step, endpoint forced */
    double step = (n > 1) ? (b - a) / (n - 1) : 0;
    for (int i = 0; i < n; i++) out[i] = a + i * step;
    if (n > 1) out[n - 1] = b;
}
static v3 launch_dir(double th, double ph) { return V(sin(th) * cos(ph), sin(th) * sin(ph), cos(th)); }

/* ------------------------------------------------------------------ rx_to_lyot_model
 * ray_trace_int.py:579-885.  out is (20, n_pts) row-major.                              */
ORC_API int orc_nf_rx_trace(const double *p_rx, double th_fwhp, double n_si, double n_vac,
                            int n_scan, double *out)
{
    int n_pts = n_scan * n_scan, rc = ORC_OK;
    double *th = malloc(sizeof(double) * n_scan), *ph = malloc(sizeof(double) * n_scan);
    linspace(-HALF_PI - 0.28, -HALF_PI + 0.28, n_scan, th);
    linspace(HALF_PI - 0.28, HALF_PI + 0.28, n_scan, ph);
    double mth = 0, mph = 0;
    for (int i = 0; i < n_scan; i++) { mth += th[i]; mph += ph[i]; }
    mth /= n_scan; mph /= n_scan;
    double sig = th_fwhp / sqrt(8 * log(2.0));
    v3 p0 = V(p_rx[0], p_rx[1], p_rx[2]);
    nf_surf lyot = nf_surface(S_LYOT);
    for (int ii = 0; ii < n_pts && rc == ORC_OK; ii++) {
        double t_ = th[ii % n_scan], p_ = ph[ii / n_scan]; /* meshgrid(theta,phi): theta fast */
        nf_leg a, b, l;
        if ((rc = nf_interface(S_3A, p0, launch_dir(t_, p_), 40, 700, n_vac, n_si, 1, 1, &a))) break;
        if ((rc = nf_interface(S_3B, a.hit, a.og_tele, 5, 200, n_si, n_vac, 1, 1, &b))) break;
        /* Lyot plane, rti:769-846: plane hit only; no refraction is applied to the output dir */
        nf_root_ctx ctx = {b.hit, b.og_tele, &lyot};
        double tl = brentq(nf_root, 1, 2000, &ctx, &rc);
        if (rc) break;
        l.hit = along(b.hit, b.og_tele, tl);
        v3 vec = vsub(tele_into(&lyot, b.hit), tele_into(&lyot, l.hit));
        double dl = sqrt(vdot(vec, vec));
        v3 nt = rot_back(-HALF_PI, V(0, 0, 1)); /* th2_l3, rti:833-835 */
        v3 og = rot_back(-HALF_PI, b.og_local); /* rti:844-846 re-rotates the 3B output */
        double opl = (b.dist * n_si) + a.dist + dl;  /* rti:848 */
        double amp = exp(-0.5 * ((t_ - mth) * (t_ - mth) + (p_ - mph) * (p_ - mph)) / (sig * sig));
        double row[20] = {l.hit.x, l.hit.y, l.hit.z, opl, amp, nt.x, nt.y, nt.z, og.x, og.y, og.z,
                          b.hit.x, b.hit.y, b.hit.z, a.hit.x, a.hit.y, a.hit.z, p0.x, p0.y, p0.z};
        for (int r = 0; r < 20; r++) out[(size_t)r * n_pts + ii] = row[r];
    }
    free(th); free(ph);
    return rc;
}

/* ------------------------------------------------------------------ source_to_lyot_model
 * ray_trace_int.py:888-1382.  Traces one ray; row[26].  Returns 0, or <0 on bracket miss. */
static int nf_source_ray(v3 p0, v3 dir, double n_si, double n_vac, double *row)
{
    nf_leg l1b, l1a, l2b, l2a;
    int rc;
    memset(row, 0, 26 * sizeof(double));
    if ((rc = nf_interface(S_1B, p0, dir, 5, 14000, n_vac, n_si, 1, 1, &l1b))) return rc;
    if ((rc = nf_interface(S_1A, l1b.hit, l1b.og_tele, 1, 900, n_si, n_vac, -1, 1, &l1a))) return rc;
    if (sqrt(l1a.hit.x * l1a.hit.x + l1a.hit.z * l1a.hit.z) >= (392 / 2)) return ORC_OK; /* rti:1086 */
    if ((rc = nf_interface(S_2B, l1a.hit, l1a.og_tele, 5, 600, n_vac, n_si, -1, 1, &l2b))) return rc;
    if (sqrt(l2b.hit.x * l2b.hit.x + l2b.hit.z * l2b.hit.z) >= (352 / 2)) return ORC_OK; /* rti:1177 */
    if ((rc = nf_interface(S_2A, l2b.hit, l2b.og_tele, 0, 100, n_si, n_vac, -1, 1, &l2a))) return rc;
    nf_surf lyot = nf_surface(S_LYOT);
    nf_root_ctx ctx = {l2a.hit, l2a.og_tele, &lyot};
    double tl = brentq(nf_root, 10, 900, &ctx, &rc);
    if (rc) return rc;
    v3 hl = along(l2a.hit, l2a.og_tele, tl);
    /* rti:1305-1339: vector points from the Lyot hit back to 2A; Snell si->vac on it (NaN on TIR) */
    v3 vec = vsub(tele_into(&lyot, l2a.hit), tele_into(&lyot, hl));
    v3 tan = vscale(vec, 1 / sqrt(vdot(vec, vec)));
    v3 og = rot_back(-HALF_PI, snell_vec(n_si, n_vac, V(0, 0, 1), tan));
    v3 nt = rot_back(-HALF_PI, V(0, 0, 1));
    double opl = (l1b.t) + (l1a.t * n_si) + (l2b.t) + (l2a.t * n_si) + tl; /* rti:1341-1343 */
    double r[26] = {hl.x, hl.y, hl.z, opl, 1, nt.x, nt.y, nt.z, og.x, og.y, og.z,
                    l2b.hit.x, l2b.hit.y, l2b.hit.z, l2a.hit.x, l2a.hit.y, l2a.hit.z,
                    l1b.hit.x, l1b.hit.y, l1b.hit.z, l1a.hit.x, l1a.hit.y, l1a.hit.z, p0.x, p0.y, p0.z};
    memcpy(row, r, sizeof(r));
    return ORC_OK;
}

ORC_API int orc_nf_source_trace(const double *p_src, double n_si, double n_vac, int n_scan, double *out)
{
    int n_pts = n_scan * n_scan, rc = ORC_OK;
    double *ang = malloc(sizeof(double) * n_scan);
    linspace(HALF_PI - 0.22, HALF_PI + 0.22, n_scan, ang); /* rti:908-909, theta and phi share it */
    v3 p0 = V(p_src[0], p_src[1], p_src[2]);
    for (int ii = 0; ii < n_pts; ii++) {
        double row[26];
        int r = nf_source_ray(p0, launch_dir(ang[ii % n_scan], ang[ii / n_scan]), n_si, n_vac, row);
        if (r) { rc = r; break; }
        for (int k = 0; k < 26; k++) out[(size_t)k * n_pts + ii] = row[k];
    }
    free(ang);
    return rc;
}

/* ------------------------------------------------------------------ near-field map
 * scripts/holo_sim.py:157-185 per pixel, with the pixel-independent receiver field passed
 * in as (rx_opl, rx_amp_masked) in ray-index space (regrid, rti:1385-1440, is the identity
 * for equal grids; the Lyot mask of holo_sim.py:168-180 is folded into rx_amp_masked).
 * amp/phase: [nfreq][npix].                                                               */
typedef struct {
    const double *src_xz, *rx_opl, *rx_amp_masked, *k_per_m, *ang;
    double y_ap, n_si, n_vac, *amp, *phase;
    int npix, n_scan, nfreq;
    atomic_int rc;
} nf_map_ctx;
static void nf_map_pixel(int p, void *v)
{
    nf_map_ctx *m = (nf_map_ctx *)v;
    int n_scan = m->n_scan, n_pts = n_scan * n_scan, nfreq = m->nfreq;
    v3 p0 = V(m->src_xz[2 * p], m->y_ap, m->src_xz[2 * p + 1]);
    double *re = calloc(2 * nfreq, sizeof(double)), *im = re + nfreq;
    for (int ii = 0; ii < n_pts; ii++) {
        double row[26];
        int r = nf_source_ray(p0, launch_dir(m->ang[ii % n_scan], m->ang[ii / n_scan]), m->n_si, m->n_vac, row);
        if (r) { atomic_store(&m->rc, r); continue; }
        double w = row[4] * m->rx_amp_masked[ii];
        if (w == 0) continue;
        for (int f = 0; f < nfreq; f++) {
            double k = m->k_per_m[f];
            double a1 = m->rx_opl[ii] * k / 1e3, a2 = row[3] * k / 1e3; /* holo_sim.py:176-177 */
            double c1 = cos(a1), s1 = sin(a1), c2 = cos(a2), s2 = sin(a2);
            re[f] += w * (c1 * c2 - s1 * s2);
            im[f] += w * (s1 * c2 + c1 * s2);
        }
    }
    for (int f = 0; f < nfreq; f++) {
        m->amp[(size_t)f * m->npix + p] = hypot(re[f], im[f]);
        m->phase[(size_t)f * m->npix + p] = atan2(im[f], re[f]);
    }
    free(re);
}
ORC_API int orc_nf_map(const double *src_xz, int npix, double y_ap, double n_si, double n_vac,
                       int n_scan, const double *rx_opl, const double *rx_amp_masked,
                       const double *k_per_m, int nfreq, double *amp, double *phase, int nthreads)
{
    double *ang = malloc(sizeof(double) * n_scan);
    linspace(HALF_PI - 0.22, HALF_PI + 0.22, n_scan, ang);
    nf_map_ctx m = {src_xz, rx_opl, rx_amp_masked, k_per_m, ang, y_ap, n_si, n_vac, amp, phase, npix, n_scan, nfreq, 0};
    atomic_init(&m.rc, ORC_OK);
    par_for(npix, nthreads, nf_map_pixel, &m);
    free(ang);
    return atomic_load(&m.rc);
}

/* ------------------------------------------------------------------ toy FP->window trace
 * ray_trace_int.py:33-576 (`aperature_fields`), (32, n_pts).                             */
ORC_API int orc_nf_toy_trace(const double *p_rx, double y_ap, double th_fwhp, double n_si,
                             double n_vac, int n_scan, double *out)
{
    int n_pts = n_scan * n_scan, rc = ORC_OK;
    double *th = malloc(sizeof(double) * n_scan), *ph = malloc(sizeof(double) * n_scan);
    linspace(-HALF_PI - 0.25, -HALF_PI + 0.25, n_scan, th);
    linspace(HALF_PI - 0.25, HALF_PI + 0.25, n_scan, ph);
    double sig = th_fwhp / sqrt(8 * log(2.0));
    v3 p0 = V(p_rx[0], p_rx[1], p_rx[2]);
    for (int ii = 0; ii < n_pts; ii++) {
        nf_leg a3, b3, a2, b2, a1, b1;
        v3 d0 = launch_dir(th[ii % n_scan], ph[ii / n_scan]);
        if ((rc = nf_interface(S_3A, p0, d0, 2, 1600, n_vac, n_si, 1, 1, &a3))) break;
        if ((rc = nf_interface(S_3B, a3.hit, a3.og_tele, 5, 200, n_si, n_vac, 1, 1, &b3))) break;
        if ((rc = nf_interface(S_2A, b3.hit, b3.og_tele, 10, 2000, n_vac, n_si, 1, 1, &a2))) break;
        if ((rc = nf_interface(S_2B, a2.hit, a2.og_tele, 5, 400, n_si, n_vac, 1, 1, &b2))) break;
        if ((rc = nf_interface(S_1A, b2.hit, b2.og_tele, 5, 500, n_vac, n_si, 1, 1, &a1))) break;
        if ((rc = nf_interface(S_1B, a1.hit, a1.og_tele, 5, 200, n_si, n_vac, 1, -1, &b1))) break;
        v3 og = b1.og_tele;
        double d_ap = fabs((y_ap - b1.hit.y) / og.y); /* rti:489 */
        double opl = a3.dist + b3.dist * n_si + a2.dist + b2.dist * n_si + a1.dist + b1.dist * n_si + d_ap;
        v3 pos = vsub(b1.hit, vscale(og, d_ap)); /* rti:500 */
        v3 t0 = rot_back(-HALF_PI, a3.tan_local); /* tan_rx_m3a_t, rti:147-154 */
        double de_ve = atan(t0.x / (-t0.y));
        double de_ho = atan(t0.z / sqrt(t0.x * t0.x + t0.y * t0.y));
        double amp = exp(-0.5 * (de_ho * de_ho + de_ve * de_ve) / (sig * sig));
        double row[32] = {pos.x, pos.y, pos.z, b1.hit.x, b1.hit.y, b1.hit.z, a1.hit.x, a1.hit.y, a1.hit.z,
                          b2.hit.x, b2.hit.y, b2.hit.z, a2.hit.x, a2.hit.y, a2.hit.z,
                          b3.hit.x, b3.hit.y, b3.hit.z, a3.hit.x, a3.hit.y, a3.hit.z, p0.x, p0.y, p0.z,
                          b1.n_local.x, b1.n_local.y, b1.n_local.z, og.x, og.y, og.z, opl, amp};
        for (int r = 0; r < 32; r++) out[(size_t)r * n_pts + ii] = row[r];
    }
    free(th); free(ph);
    return rc;
}

/* =================================================================== far field (LAT)
 * latrt_geo.py:7-294, solat_apert_field.py:307-567, :570-1009, far_field_latrt.py:9-32,
 * scripts/ff_sim.py:204-221.                                                            */
typedef struct {
    double th_1, th2;      /* latrt_geo.py:13-15 */
    double rotc[3];        /* mm, latrt_geo.py:31-33 (x1e3) */
    double tow[3];         /* mm, source position */
    double rx_y_mm;        /* y_ap = tele_geo.rx_y*1e3, saf:582 */
} ff_geo;

static const double A1C[7][7] = {
    {0, 0, -57.74022, 1.5373825, 1.154294, -0.441762, 0.0906601},
    {0, 0, 0, 0, 0, 0, 0},
    {-72.17349, 1.8691899, 2.8859421, -1.026471, 0.2610568, 0, 0},
    {0, 0, 0, 0, 0, 0, 0},
    {1.8083973, -0.603195, 0.2177414, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0},
    {0.0394559, 0, 0, 0, 0, 0, 0}};
static const double A2C[8][8] = {
    {0, 0, 103.90461, 6.6513025, 2.8405781, -0.7819705, -0.0400483, 0.0896645},
    {0, 0, 0, 0, 0, 0, 0, 0},
    {115.44758, 7.3024355, 5.7640389, -1.578144, -0.0354326, 0.2781226, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0},
    {2.9130983, -0.8104051, -0.0185283, 0.2626023, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0},
    {-0.0250794, 0.0709672, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0}};
#define R_N 3000.0
#define Y_CENT_M1 (-7201.003729431267) /* saf:19 */

/* latrt_geo.py:77-150.  Zero coefficients are skipped and integer powers are built by
 * repeated multiplication (the reference evaluates all 49/64 terms with float pow).    */
static void poly_eval(const double *a, int n, double x, double y, double *z, double *gx, double *gy)
{
    double xp[8], yp[8], u = x / R_N, v = y / R_N;
    xp[0] = yp[0] = 1;
    for (int i = 1; i < n; i++) { xp[i] = xp[i - 1] * u; yp[i] = yp[i - 1] * v; }
    double s = 0, sx = 0, sy = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double c = a[i * n + j];
            if (c == 0) continue;
            s += c * xp[i] * yp[j];
            if (i > 0) sx += c * (i / R_N) * xp[i - 1] * yp[j];
            if (j > 0) sy += c * xp[i] * (j / R_N) * yp[j - 1];
        }
    if (z) *z = s;
    if (gx) { *gx = sx; *gy = sy; }
}

/* latrt_geo.py:272-294 */
static v3 rotate_az_el(const ff_geo *g, v3 p, double el, double az)
{
    double x = p.x - g->rotc[0], y = p.y - g->rotc[1], z = p.z - g->rotc[2];
    double yt = cos(el) * y - sin(el) * z, zt = sin(el) * y + cos(el) * z;
    double xo = cos(az) * x + sin(az) * zt, zo = -sin(az) * x + cos(az) * zt;
    return V(xo + g->rotc[0], yt + g->rotc[1], zo + g->rotc[2]);
}
/* latrt_geo.py:183-199 */
static v3 tele_into_m1(const ff_geo *g, v3 p, double el, double az)
{
    v3 q = rotate_az_el(g, p, el, az);
    double th1 = g->th_1;
    q.z += 3600; q.y += 7200;
    double yt = q.y * cos(-th1) - q.z * sin(-th1), zt = q.y * sin(-th1) + q.z * cos(-th1);
    return V(q.x * cos(M_PI) + zt * sin(M_PI), yt, -q.x * sin(M_PI) + zt * cos(M_PI));
}
/* latrt_geo.py:202-218 */
static v3 tele_into_m2(const ff_geo *g, v3 p, double el, double az)
{
    v3 q = rotate_az_el(g, p, el, az);
    double th2 = g->th2;
    q.y += 4800 + 7200;
    double yt = q.y * cos(-th2) - q.z * sin(-th2), zt = q.y * sin(-th2) + q.z * cos(-th2);
    return V(q.x * cos(-M_PI) - yt * sin(-M_PI), q.x * sin(-M_PI) + yt * cos(-M_PI), zt);
}
/* latrt_geo.py:253-269 */
static v3 tele_into_foc(const ff_geo *g, v3 p, double el, double az)
{
    v3 q = rotate_az_el(g, p, el, az);
    return V(q.x, q.y * cos(-M_PI / 2) - q.z * sin(-M_PI / 2), q.y * sin(-M_PI / 2) + q.z * cos(-M_PI / 2));
}

typedef struct {
    const ff_geo *g; v3 p, d; double el, az;
    int which; /* 0 aperture, 1 M1, 2 M2, 3 focal */
    const double *pan; /* NULL or pointer to 10 panel rows [row,col,a,b,c,d,e,f,x0,y0] */
} ff_root_ctx;

static double ff_root(double t, void *vctx)
{
    ff_root_ctx *c = (ff_root_ctx *)vctx;
    v3 P = along(c->p, c->d, t);
    if (c->which == 0) { /* saf:359-370 + latrt_geo.py:237-245 */
        v3 q = rotate_az_el(c->g, P, -c->el, -c->az);
        return (q.z - 4e3) - 0;
    }
    if (c->which == 3) { /* saf:947-960 */
        v3 q = tele_into_foc(c->g, P, c->el, c->az);
        return q.z - (0 - c->g->rx_y_mm);
    }
    double zs, zerr = 0;
    if (c->which == 1) {
        v3 l = tele_into_m1(c->g, P, -c->el, -c->az);
        poly_eval(&A1C[0][0], 7, l.x, l.y, &zs, NULL, NULL);
        if (c->pan) { /* saf:676-703 */
            const double *q = c->pan;
            double xt = l.x * cos(M_PI) + l.z * sin(M_PI), yt = l.y;
            double xpc = xt - q[8], ypc = yt - q[9];
            zerr = q[2] + q[3] * xpc + q[4] * ypc + q[5] * (xpc * xpc + ypc * ypc) + q[6] * (xpc * ypc);
        }
        return l.z - (zs + zerr);
    }
    v3 l = tele_into_m2(c->g, P, -c->el, -c->az);
    poly_eval(&A2C[0][0], 8, l.x, l.y, &zs, NULL, NULL);
    if (c->pan) { /* saf:820-852 incl. the z /= cos(atan(1/3)) tweak of the error coordinates */
        const double *q = c->pan;
        v3 P2 = P;
        if (c->g->tow[2] != 0) P2.z /= cos(atan(1.0 / 3.0));
        v3 le = tele_into_m2(c->g, P2, -c->el, -c->az);
        double xt = le.x * cos(M_PI) + le.z * sin(M_PI), yt = le.y;
        double xpc = xt - q[8], ypc = yt - q[9];
        zerr = q[2] + q[3] * xpc + q[4] * ypc + q[5] * (xpc * xpc + ypc * ypc) + q[6] * (xpc * ypc);
    }
    return l.z - (zs + zerr);
}

/* direction from a mirror frame back to the telescope frame, saf:215-257 (M1) / :885-930 (M2),
 * including the reference's index slips and its (+el,+az) back-rotation.                 */
static v3 m1_vec_to_tele(const ff_geo *g, v3 v, double el, double az, int slip)
{
    double th1 = g->th_1;
    double xt = slip ? (v.x * cos(M_PI) - v.y * sin(M_PI)) : (v.x * cos(M_PI) + v.z * sin(M_PI));
    double yt = v.y, zt = -v.z * sin(M_PI) + v.z * cos(M_PI);
    v3 r = V(xt, yt * cos(th1) - zt * sin(th1), yt * sin(th1) + zt * cos(th1));
    double y2 = cos(el) * r.y - sin(el) * r.z, z2 = sin(el) * r.y + cos(el) * r.z;
    return V(cos(az) * r.x + sin(az) * z2, y2, -sin(az) * r.x + cos(az) * z2);
}
static v3 m2_vec_to_tele(const ff_geo *g, v3 v, double el, double az)
{
    double th2 = g->th2;
    double xt = v.x * cos(M_PI) - v.y * sin(M_PI), yt = v.x * sin(M_PI) + v.y * cos(M_PI), zt = v.z;
    v3 r = V(xt, yt * cos(th2) - zt * sin(th2), yt * sin(th2) + zt * cos(th2));
    double y2 = cos(el) * r.y - sin(el) * r.z, z2 = sin(el) * r.y + cos(el) * r.z;
    return V(cos(az) * r.x + sin(az) * z2, y2, -sin(az) * r.x + cos(az) * z2);
}

/* saf:323-339: focal-plane centre from a NaN-masked 100x100 grid pushed through foc_into_tele */
static void ff_focal_centre(const ff_geo *g, double el, double az, double *c3)
{
    double ax[100];
    linspace(-4000, 4000, 100, ax);
    double sx = 0, sy = 0, sz = 0;
    int nx = 0, ny = 0;
    for (int i = 0; i < 100; i++) {
        /* row 50: Y = ax[50], X = ax[i];   column 50: X = ax[50], Y = ax[i] */
        if (sqrt(ax[i] * ax[i] + ax[50] * ax[50]) < 3600) {
            v3 f = V(ax[i], ax[50] * cos(M_PI / 2) - 0 * sin(M_PI / 2), ax[50] * sin(M_PI / 2) + 0 * cos(M_PI / 2));
            sx += rotate_az_el(g, f, el, az).x; nx++;
            v3 h = V(ax[50], ax[i] * cos(M_PI / 2) - 0 * sin(M_PI / 2), ax[i] * sin(M_PI / 2) + 0 * cos(M_PI / 2));
            v3 r = rotate_az_el(g, h, el, az);
            sy += r.y; sz += r.z; ny++;
        }
    }
    c3[0] = sx / nx; c3[1] = sy / ny; c3[2] = sz / ny;
}

/* one far-field ray.  stage 1 (ray_mirror_pts) when pan1==pan2==NULL && !full; returns
 * 1 if the ray produced output, 0 if rejected, <0 on bracket error.                      */
typedef struct { v3 m1, m2, n2_t, foc, og; double opl; } ff_ray_out;
static int ff_ray(const ff_geo *g, v3 dir, double el, double az, int full, const double *pan1,
                  const double *pan2, ff_ray_out *o)
{
    int rc = ORC_OK;
    v3 src = V(g->tow[0], g->tow[1], g->tow[2]);
    ff_root_ctx c = {g, src, dir, el, az, 0, NULL};
    double t_ap = brentq(ff_root, 1e4, 1e9, &c, &rc);
    if (rc) return rc;
    v3 Pa = along(src, dir, t_ap);
    if (Pa.x * Pa.x + (Pa.y - Y_CENT_M1) * (Pa.y - Y_CENT_M1) > 3e3 * 3e3) return 0; /* saf:378,664 */
    c.p = Pa; c.which = 1; c.pan = pan1;
    double t_m1 = full ? brentq(ff_root, 0, 1e4, &c, &rc) : brentq(ff_root, 0, 5e4, &c, &rc); /* saf:706 / :396 */
    if (rc) return rc;
    v3 Pm1 = along(Pa, dir, t_m1);
    v3 l1 = tele_into_m1(g, Pm1, -el, -az), ls = tele_into_m1(g, src, -el, -az);
    double gx, gy;
    poly_eval(&A1C[0][0], 7, l1.x, l1.y, NULL, &gx, &gy);
    v3 nt = V(-gx, -gy, 1), N = vscale(nt, 1 / sqrt(vdot(nt, nt)));
    v3 vec = vsub(l1, ls);
    v3 tan = vscale(vec, 1 / sqrt(vdot(vec, vec)));
    v3 og = vsub(tan, vscale(N, 2 * vdot(tan, N))); /* saf:208,733 */
    v3 d1 = m1_vec_to_tele(g, og, el, az, 1);
    c.p = Pm1; c.d = d1; c.which = 2; c.pan = pan2;
    double t_m2 = full ? brentq(ff_root, 1e3, 13e3, &c, &rc) : brentq(ff_root, 0, 20e3, &c, &rc); /* saf:855 / :490 */
    if (rc) return rc;
    v3 Pm2 = along(Pm1, d1, t_m2);
    o->m1 = Pm1; o->m2 = Pm2;
    if (!full) return 1;
    v3 l2 = tele_into_m2(g, Pm2, -el, -az), lp = tele_into_m2(g, Pm1, -el, -az);
    poly_eval(&A2C[0][0], 8, l2.x, l2.y, NULL, &gx, &gy);
    nt = V(-gx, -gy, 1); N = vscale(nt, 1 / sqrt(vdot(nt, nt)));
    vec = vsub(l2, lp);
    tan = vscale(vec, 1 / sqrt(vdot(vec, vec)));
    og = vsub(tan, vscale(N, 2 * vdot(tan, N))); /* saf:883 */
    o->n2_t = m2_vec_to_tele(g, N, el, az);
    v3 d2 = m2_vec_to_tele(g, og, el, az);
    c.p = Pm2; c.d = d2; c.which = 3; c.pan = NULL;
    double t_foc = brentq(ff_root, 1e3, 15e3, &c, &rc); /* saf:962 */
    if (rc) return rc;
    o->foc = along(Pm2, d2, t_foc);
    o->og = d2;
    o->opl = t_ap + t_m1 + t_m2 + t_foc; /* saf:970 */
    return 1;
}

/* geo: [th_1, th2, rotc xyz (mm), tow xyz (mm), rx_y_mm] = 9 doubles */
static ff_geo ff_unpack(const double *gv)
{
    ff_geo g;
    g.th_1 = gv[0]; g.th2 = gv[1];
    for (int i = 0; i < 3; i++) { g.rotc[i] = gv[2 + i]; g.tow[i] = gv[5 + i]; }
    g.rx_y_mm = gv[8];
    return g;
}

/* solat_apert_field.py:307-567.  out (9, n*n). */
ORC_API int orc_ff_ray_mirror_pts(const double *geo, const double *theta, const double *phi, int n,
                                  double el, double az, double *out)
{
    ff_geo g = ff_unpack(geo);
    int n_pts = n * n, rc = ORC_OK;
    memset(out, 0, sizeof(double) * 9 * n_pts);
    for (int ii = 0; ii < n_pts; ii++) {
        ff_ray_out o;
        int r = ff_ray(&g, launch_dir(theta[ii % n], phi[ii / n]), el, az, 0, NULL, NULL, &o);
        if (r < 0) { rc = r; break; }
        if (r == 0) continue;
        out[0 * n_pts + ii] = o.m2.x; out[1 * n_pts + ii] = o.m2.y; out[2 * n_pts + ii] = o.m2.z;
        out[3 * n_pts + ii] = o.m1.x; out[4 * n_pts + ii] = o.m1.y; out[5 * n_pts + ii] = o.m1.z;
    }
    double c3[3];
    ff_focal_centre(&g, el, az, c3);
    out[6 * n_pts] = c3[0]; out[7 * n_pts] = c3[1]; out[8 * n_pts] = c3[2];
    return rc;
}

/* solat_apert_field.py:570-1009.  pan_idx1/2[n*n]: index into the panel tables of the
 * first panel whose (row,col) matches the ray's identify_panel ids, or -1 (host-side
 * np.where, saf:626,775-778).  pan1/pan2: (10, np1)/(10, np2) row-major.  out (17, n*n). */
ORC_API int orc_ff_apert_fields(const double *geo, const double *theta, const double *phi, int n,
                                const double *focal_centre, const int *pan_idx1, const int *pan_idx2,
                                const double *pan1, int np1, const double *pan2, int np2,
                                double el, double az, double *out)
{
    ff_geo g = ff_unpack(geo);
    int n_pts = n * n, rc = ORC_OK;
    memset(out, 0, sizeof(double) * 17 * n_pts);
    for (int ii = 0; ii < n_pts; ii++) {
        if (pan_idx1[ii] < 0) continue;
        double q1[10], q2[10];
        for (int r = 0; r < 10; r++) q1[r] = pan1[r * np1 + pan_idx1[ii]];
        int have2 = pan_idx2[ii] >= 0;
        if (have2) for (int r = 0; r < 10; r++) q2[r] = pan2[r * np2 + pan_idx2[ii]];
        ff_ray_out o;
        /* a ray whose M2 panel is missing produces no output (saf:779): trace needs q2 only at M2 */
        if (!have2) continue;
        int r = ff_ray(&g, launch_dir(theta[ii % n], phi[ii / n]), el, az, 1, q1, q2, &o);
        if (r < 0) { rc = r; break; }
        if (r == 0) continue;
        double row[17] = {o.m2.x, o.m2.y, o.m2.z, o.m1.x, o.m1.y, o.m1.z, o.n2_t.x, o.n2_t.y, o.n2_t.z,
                          o.foc.x - focal_centre[0], o.foc.y - focal_centre[1], o.foc.z - focal_centre[2],
                          o.og.x, o.og.y, o.og.z, o.opl, 1};
        for (int k = 0; k < 17; k++) out[(size_t)k * n_pts + ii] = row[k];
    }
    return rc;
}

static int searchsorted_left(const double *a, int n, int stride, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (a[(size_t)mid * stride] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
static double max_strided(const double *a, int n, int stride)
{
    double m = a[0];
    for (int i = 1; i < n; i++) if (a[(size_t)i * stride] > m) m = a[(size_t)i * stride];
    return m;
}

/* far_field_latrt.py:9-32.  sim (17, n_pts); data_x, data_y (M, M); out complex (M, M) as
 * interleaved re,im.  Returns number of cells written (distinct).                        */
ORC_API int orc_ff_project(const double *sim, int n_pts, const double *data_x, const double *data_y,
                           int M, double k, double *out_reim)
{
    memset(out_reim, 0, sizeof(double) * 2 * M * M);
    double xmax = max_strided(data_x, M, 1), ymax = max_strided(data_y, M, M);
    for (int ii = 0; ii < n_pts; ii++) {
        double L = sim[(size_t)15 * n_pts + ii];
        if (L == 0) continue;
        double xx = sim[(size_t)9 * n_pts + ii] / 1e1, yy = sim[(size_t)11 * n_pts + ii] / 1e1;
        if (xmax < xx || ymax < yy) continue;
        int ix = searchsorted_left(data_x, M, 1, xx);
        int iy = searchsorted_left(data_y + ix, M, M, yy);
        double a = k * L / 1e3;
        out_reim[2 * ((size_t)iy * M + ix)] = cos(a);
        out_reim[2 * ((size_t)iy * M + ix) + 1] = sin(a);
    }
    int cnt = 0;
    for (int i = 0; i < M * M; i++) if (out_reim[2 * i] != 0 || out_reim[2 * i + 1] != 0) cnt++;
    return cnt;
}

/* scripts/ff_sim.py:204-221 over a list of (az, el) pixels, one-panel/zero-error panel model
 * (pan_mod stand-in) unless pan tables are given.  two_pass!=0 also runs the ray_mirror_pts
 * stage per ray exactly as the reference does (its outputs only feed identify_panel and
 * the focal centre).  amp, phase: [npix].                                                 */
typedef struct {
    ff_geo g;
    const double *azel, *theta, *phi, *data_x, *data_y, *e_rec;
    double k, xmax, ymax, *amp, *phase;
    int n, M, two_pass;
    atomic_int rc;
} ff_map_ctx;
static void ff_map_pixel(int p, void *v)
{
    ff_map_ctx *m = (ff_map_ctx *)v;
    const ff_geo *g = &m->g;
    int n = m->n, n_pts = n * n, M = m->M;
    double zero_panel[10] = {1, 1, 0, 0, 0, 0, 0, 0, 0, 0};
    double az = m->azel[2 * p], el = m->azel[2 * p + 1], c3[3];
    double *cell = calloc((size_t)2 * M * M, sizeof(double));
    ff_focal_centre(g, el, az, c3);
    for (int ii = 0; ii < n_pts; ii++) {
        ff_ray_out o;
        v3 d = launch_dir(m->theta[ii % n], m->phi[ii / n]);
        if (m->two_pass) {
            int r0 = ff_ray(g, d, el, az, 0, NULL, NULL, &o);
            if (r0 < 0) atomic_store(&m->rc, r0);
        }
        int r = ff_ray(g, d, el, az, 1, zero_panel, zero_panel, &o);
        if (r < 0) { atomic_store(&m->rc, r); continue; }
        if (r == 0 || o.opl == 0) continue;
        double xx = (o.foc.x - c3[0]) / 1e1, yy = (o.foc.z - c3[2]) / 1e1;
        if (m->xmax < xx || m->ymax < yy) continue;
        int ix = searchsorted_left(m->data_x, M, 1, xx);
        int iy = searchsorted_left(m->data_y + ix, M, M, yy);
        double a = m->k * o.opl / 1e3;
        cell[2 * ((size_t)iy * M + ix)] = cos(a);
        cell[2 * ((size_t)iy * M + ix) + 1] = sin(a);
    }
    double re = 0, im = 0;
    for (int i = 0; i < M * M; i++) {
        double cr = cell[2 * i], ci = cell[2 * i + 1], er = m->e_rec[2 * i], ei = m->e_rec[2 * i + 1];
        re += cr * er - ci * ei;
        im += cr * ei + ci * er;
    }
    re /= (double)M * M; im /= (double)M * M; /* Npts = len(out_new) = M, ff_sim.py:216-217 */
    m->amp[p] = hypot(re, im);
    m->phase[p] = atan2(im, re);
    free(cell);
}
ORC_API int orc_ff_map(const double *geo, const double *azel, int npix, const double *theta,
                       const double *phi, int n, const double *data_x, const double *data_y,
                       const double *e_rec_reim, int M, double k, int two_pass, double *amp,
                       double *phase, int nthreads)
{
    ff_map_ctx m;
    m.g = ff_unpack(geo);
    m.azel = azel; m.theta = theta; m.phi = phi; m.data_x = data_x; m.data_y = data_y; m.e_rec = e_rec_reim;
    m.k = k; m.xmax = max_strided(data_x, M, 1); m.ymax = max_strided(data_y, M, M);
    m.amp = amp; m.phase = phase; m.n = n; m.M = M; m.two_pass = two_pass;
    atomic_init(&m.rc, ORC_OK);
    par_for(npix, nthreads, ff_map_pixel, &m);
    return atomic_load(&m.rc);
}

ORC_API int orc_num_threads(void) { return par_default_threads(); }
