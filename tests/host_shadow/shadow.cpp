// Host shadow of the device math: compiles solat_optics_b200/csrc/solat_math.cuh with g++ so that the
// EXACT per-ray source the kernels run can be checked against the golden fixtures in a container that
// has no GPU.  Test infrastructure only (built into tests/_build/, never shipped, never loaded by the
// product package).  The loops here stand in for the CUDA grid; they carry no arithmetic of their own.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../solat_optics_b200/csrc/solat_math.cuh"

using namespace solat;
#define API extern "C" __attribute__((visibility("default")))

API int shadow_nf_source_rows(const double* p, const double* th, int nth, const double* ph, int nph, double n_si,
                              double n_vac, double* out, int* status)
{
    const int n = nth * nph;
    int agg = 0;
    for (int r = 0; r < n; r++) {
        const int i = r % nth, j = r / nth;
        NfSourceRay ray;
        vec3 p0 = mk(p[0], p[1], p[2]);
        nf_source_ray(K_COEF, p0, launch_dir(sin(th[i]), cos(th[i]), sin(ph[j]), cos(ph[j])), n_si, n_vac, ray);
        double rows[26];
        pack_nf_source_rows(ray, p0, n_si, n_vac, rows);
        for (int k = 0; k < 26; k++) out[(size_t)k * n + r] = rows[k];
        if (status) status[r] = ray.status;
        agg |= ray.status;
    }
    return agg;
}
API int shadow_nf_rx_rows(const double* p, const double* th, int nth, const double* ph, int nph, double th_mean,
                          double ph_mean, double th_fwhp, double n_si, double n_vac, double* out)
{
    const int n = nth * nph;
    int agg = 0;
    const double sig = th_fwhp / sqrt(8 * log(2.0));
    for (int r = 0; r < n; r++) {
        const int i = r % nth, j = r / nth;
        NfRxRay ray;
        vec3 p0 = mk(p[0], p[1], p[2]);
        nf_rx_ray(K_COEF, p0, launch_dir(sin(th[i]), cos(th[i]), sin(ph[j]), cos(ph[j])), n_si, n_vac, ray);
        double rows[20];
        pack_nf_rx_rows(ray, p0, th[i] - th_mean, ph[j] - ph_mean, 0.5 / (sig * sig), rows);
        for (int k = 0; k < 20; k++) out[(size_t)k * n + r] = rows[k];
        agg |= ray.status;
    }
    return agg;
}
API int shadow_nf_toy_rows(const double* p, const double* th, int nth, const double* ph, int nph, double y_ap,
                           double th_fwhp, double n_si, double n_vac, double* out)
{
    const int n = nth * nph;
    int agg = 0;
    const double sig = th_fwhp / sqrt(8 * log(2.0));
    for (int r = 0; r < n; r++) {
        const int i = r % nth, j = r / nth;
        NfToyRay ray;
        vec3 p0 = mk(p[0], p[1], p[2]);
        nf_toy_ray(K_COEF, p0, launch_dir(sin(th[i]), cos(th[i]), sin(ph[j]), cos(ph[j])), y_ap, n_si, n_vac, ray);
        double rows[32];
        pack_nf_toy_rows(ray, p0, 0.5 / (sig * sig), rows);
        for (int k = 0; k < 32; k++) out[(size_t)k * n + r] = rows[k];
        agg |= ray.status;
    }
    return agg;
}
// one near-field pixel: sum over rays of w * exp(i k (L_rx + L_so)/1e3)
API int shadow_nf_pixel(const double* p, const double* th, int nth, const double* ph, int nph, double n_si,
                        double n_vac, const double* rx_opl, const double* rx_w, double k, double* re_im)
{
    const int n = nth * nph;
    double re = 0, im = 0;
    int agg = 0;
    for (int r = 0; r < n; r++) {
        if (rx_w[r] == 0.0) continue;
        const int i = r % nth, j = r / nth;
        NfSourceRay ray;
        nf_source_ray(K_COEF, mk(p[0], p[1], p[2]), launch_dir(sin(th[i]), cos(th[i]), sin(ph[j]), cos(ph[j])), n_si, n_vac, ray);
        agg |= ray.status;
        if (!ray.valid) continue;
        // as in nf_map_kernel: receiver phasor (libm, host side) x exp(i k L_source) from sincos_mid
        const double ar = rx_opl[r] * k / 1e3;
        double sn, cs;
        sincos_mid(ray.opl * (k / 1e3), sn, cs);
        re += rx_w[r] * (cos(ar) * cs - sin(ar) * sn);
        im += rx_w[r] * (cos(ar) * sn + sin(ar) * cs);
    }
    re_im[0] = re; re_im[1] = im;
    return agg;
}

// the same pixel with the two-rays-per-thread packs the fused map kernels use (d2 lock-step tracing)
API int shadow_nf_pixel2(const double* p, const double* th, int nth, const double* ph, int nph, double n_si,
                         double n_vac, const double* rx_opl, const double* rx_w, double k, double* re_im)
{
    const int n = nth * nph;
    std::vector<int> act;
    for (int r = 0; r < n; r++) if (rx_w[r] != 0.0) act.push_back(r);
    double re = 0, im = 0;
    int agg = 0;
    const vec3t<d2> p0 = mk<d2>(d2(p[0]), d2(p[1]), d2(p[2]));
    for (size_t a = 0; a < act.size(); a += 2) {
        const bool two = a + 1 < act.size();
        const int r0 = act[a], r1 = two ? act[a + 1] : r0;
        const vec3 u = launch_dir(sin(th[r0 % nth]), cos(th[r0 % nth]), sin(ph[r0 / nth]), cos(ph[r0 / nth]));
        const vec3 v = launch_dir(sin(th[r1 % nth]), cos(th[r1 % nth]), sin(ph[r1 / nth]), cos(ph[r1 / nth]));
        NfSourceRayT<d2> ray;
        nf_source_ray(K_COEF, p0, mk<d2>(d2(u.x, v.x), d2(u.y, v.y), d2(u.z, v.z)), n_si, n_vac, ray);
        agg |= ray.status.a | (two ? ray.status.b : 0);
        const double w0 = ray.valid.a ? rx_w[r0] : 0.0, w1 = (two && ray.valid.b) ? rx_w[r1] : 0.0;
        const double a0 = (rx_opl[r0] + ray.opl.a) * (k / 1e3), a1 = (rx_opl[r1] + ray.opl.b) * (k / 1e3);
        re += w0 * cos(a0) + w1 * cos(a1);
        im += w0 * sin(a0) + w1 * sin(a1);
    }
    re_im[0] = re; re_im[1] = im;
    return agg;
}

static FfGrid make_grid(const double* gx_full, const double* gy, int M)
{
    FfGrid G;
    G.gx = gx_full; G.gy = gy; G.M = M; G.gy_cs = 1; G.gy_rs = M; G.y_uniform = 0;
    G.gx_min = gx_full[0]; G.gy_min = gy[0];
    for (int i = 1; i < M; i++) { G.gx_min = fmin(G.gx_min, gx_full[i]); G.gy_min = fmin(G.gy_min, gy[(size_t)i * M]); }
    G.gx_max = gx_full[0]; G.gy_max = gy[0];
    for (int i = 0; i < M; i++) { G.gx_max = fmax(G.gx_max, gx_full[i]); G.gy_max = fmax(G.gy_max, gy[(size_t)i * M]); }
    const double sx = gx_full[M - 1] - gx_full[0], sy = gy[(size_t)(M - 1) * M] - gy[0];
    G.inv_dx = (M > 1 && sx > 0) ? (float)((M - 1) / sx) : 0.0f;
    G.inv_dy = (M > 1 && sy > 0) ? (float)((M - 1) / sy) : 0.0f;
    return G;
}

static FfGeom unpack(const double* g)
{
    FfGeom q;
    q.th_1 = g[0]; q.th2 = g[1];
    for (int i = 0; i < 3; i++) { q.rotc[i] = g[2 + i]; q.src[i] = g[5 + i]; }
    q.foc_y = g[8];
    return q;
}
API int shadow_ff_mirror_rows(const double* geo, const double* th, int nth, const double* ph, int nph, double el,
                              double az, double* out)
{
    FfGeom g = unpack(geo);
    FfFrames f;
    ff_build_frames(K_COEF, g, el, az, f);
    const int n = nth * nph;
    int agg = 0;
    memset(out, 0, sizeof(double) * 9 * n);
    for (int r = 0; r < n; r++) {
        const int i = r % nth, j = r / nth;
        vec3 d = launch_dir(sin(th[i]), cos(th[i]), sin(ph[j]), cos(ph[j])), Pa;
        double t_ap;
        if (!ff_aperture(K_COEF, g, f, d, t_ap, Pa)) continue;
        FfRay ray;
        ff_ray_trace<true, false>(K_COEF, f, d, t_ap, Pa, g.foc_y, nullptr, nullptr, nullptr, ray);
        agg |= ray.status;
        const double v[6] = {ray.m2.x, ray.m2.y, ray.m2.z, ray.m1.x, ray.m1.y, ray.m1.z};
        for (int k = 0; k < 6; k++) out[(size_t)k * n + r] = v[k];
    }
    out[6 * (size_t)n] = f.centre[0]; out[7 * (size_t)n] = f.centre[1]; out[8 * (size_t)n] = f.centre[2];
    return agg;
}
API int shadow_ff_field_rows(const double* geo, const double* th, int nth, const double* ph, int nph,
                             const double* centre, const int* pidx1, const int* pidx2, const double* pan1, int np1,
                             const double* pan2, int np2, double el, double az, double* out)
{
    FfGeom g = unpack(geo);
    FfFrames f;
    ff_build_frames(K_COEF, g, el, az, f);
    double t2e[9];
    for (int i = 0; i < 9; i++) t2e[i] = f.t2.a[i];
    if (g.src[2] != 0) for (int i = 0; i < 3; i++) t2e[3 * i + 2] /= cos(atan(1.0 / 3.0));
    const int n = nth * nph;
    int agg = 0;
    memset(out, 0, sizeof(double) * 17 * n);
    for (int r = 0; r < n; r++) {
        if (pidx1[r] < 0 || pidx2[r] < 0) continue;
        const int i = r % nth, j = r / nth;
        vec3 d = launch_dir(sin(th[i]), cos(th[i]), sin(ph[j]), cos(ph[j])), Pa;
        double t_ap;
        if (!ff_aperture(K_COEF, g, f, d, t_ap, Pa)) continue;
        PanelErr e1, e2;
        const double* q = pan1 + pidx1[r];
        e1.a = q[2 * np1]; e1.b = q[3 * np1]; e1.c = q[4 * np1]; e1.d = q[5 * np1]; e1.e = q[6 * np1]; e1.x0 = q[8 * np1]; e1.y0 = q[9 * np1];
        q = pan2 + pidx2[r];
        e2.a = q[2 * np2]; e2.b = q[3 * np2]; e2.c = q[4 * np2]; e2.d = q[5 * np2]; e2.e = q[6 * np2]; e2.x0 = q[8 * np2]; e2.y0 = q[9 * np2];
        FfRay ray;
        ff_ray_trace<false, true>(K_COEF, f, d, t_ap, Pa, g.foc_y, &e1, &e2, t2e, ray);
        agg |= ray.status;
        double rows[17];
        pack_ff_rows(ray, centre, rows);
        for (int k = 0; k < 17; k++) out[(size_t)k * n + r] = rows[k];
    }
    return agg;
}
// far-field pixel traced two rays at a time (d2 packs), as the fused map kernel does
API int shadow_ff_pixel2(const double* geo, const double* th, int nth, const double* ph, int nph, double el, double az,
                         const double* gx_full, const double* gy, const double* e_rec, int M, double k, double* re_im,
                         int* n_cells)
{
    FfGeom g = unpack(geo);
    FfFrames f;
    ff_build_frames(K_COEF, g, el, az, f);
    const int n = nth * nph;
    int agg = 0;
    const FfGrid G = make_grid(gx_full, gy, M);
    std::vector<int> win(M * M, -1), queue;
    std::vector<double> opl(M * M, 0.0);
    for (int r = 0; r < n; r++) {
        vec3 d = launch_dir(sin(th[r % nth]), cos(th[r % nth]), sin(ph[r / nth]), cos(ph[r / nth])), Pa;
        double t_ap;
        const bool exact = ff_aperture(K_COEF, g, f, d, t_ap, Pa);
        // the kernel's FP32 prefilter must agree with the FP64 test wherever it commits to an answer
        const float stf = (float)sin(th[r % nth]);
        const int v = ff_aperture_coarse(f, stf * (float)cos(ph[r / nth]), stf * (float)sin(ph[r / nth]), (float)cos(th[r % nth]));
        if (v >= 0 && (v > 0) != exact) agg |= 16;
        if (exact) queue.push_back(r);
    }
    for (size_t a = 0; a < queue.size(); a += 2) {
        const bool two = a + 1 < queue.size();
        const int r0 = queue[a], r1 = two ? queue[a + 1] : r0;
        const vec3 u = launch_dir(sin(th[r0 % nth]), cos(th[r0 % nth]), sin(ph[r0 / nth]), cos(ph[r0 / nth]));
        const vec3 v = launch_dir(sin(th[r1 % nth]), cos(th[r1 % nth]), sin(ph[r1 / nth]), cos(ph[r1 / nth]));
        const vec3t<d2> d = mk<d2>(d2(u.x, v.x), d2(u.y, v.y), d2(u.z, v.z));
        vec3t<d2> Pa;
        d2 t_ap;
        ff_aperture(K_COEF, g, f, d, t_ap, Pa);
        d2 opl2, fx2, fz2;
        i2 st2;
        ff_ray_lean<d2>(K_COEF, f, d, t_ap, Pa, g.foc_y, opl2, fx2, fz2, st2);
        agg |= st2.a | (two ? st2.b : 0);
        const int rr[2] = {r0, r1};
        const double oo[2] = {opl2.a, opl2.b}, fx[2] = {fx2.a, fx2.b}, fz[2] = {fz2.a, fz2.b};
        for (int l = 0; l < (two ? 2 : 1); l++) {
            if (oo[l] == 0.0) continue;
            const int c = ff_cell(fx[l] - f.centre[0], fz[l] - f.centre[2], G);
            if (c < 0) continue;
            // through the packed 64-bit key of the fused kernel (exact fixed-point path length)
            bool exact;
            const double base = 1.0002e8, unit = ldexp(1.0, ilogb(base) - 53);
            const unsigned long long key = ff_key(rr[l], oo[l], base, 1.0 / unit, exact, unit);
            if (!exact) agg |= 8;
            if (rr[l] > win[c]) { win[c] = rr[l]; opl[c] = ff_key_opl(key, base, unit); }
        }
    }
    double re = 0, im = 0;
    int cnt = 0;
    for (int c = 0; c < M * M; c++) {
        if (win[c] < 0) continue;
        cnt++;
        double cs, sn;
        cis_large(k * opl[c] / 1e3, cs, sn);
        re += cs * e_rec[2 * c] - sn * e_rec[2 * c + 1];
        im += cs * e_rec[2 * c + 1] + sn * e_rec[2 * c];
    }
    re_im[0] = re / ((double)M * M); re_im[1] = im / ((double)M * M);
    if (n_cells) *n_cells = cnt;
    return agg;
}

// accepted rays of every phi row: interval search (what the fused map runs) next to the per-ray test; returns the
// number of rows whose interval differs from the brute-force accepted set (-1: some accepted set is not one
// interval), and how many rows asked for the per-ray walk.
API int shadow_ff_row_intervals(const double* geo, const double* th, int nth, const double* ph, int nph, double el, double az,
                                int* lo_hi, int* n_fallback)
{
    FfGeom g = unpack(geo);
    FfFrames f;
    ff_build_frames(K_COEF, g, el, az, f);
    std::vector<double> st(nth), ct(nth);
    for (int i = 0; i < nth; i++) { st[i] = sin(th[i]); ct[i] = cos(th[i]); }
    int bad = 0;
    *n_fallback = 0;
    for (int j = 0; j < nph; j++) {
        int lo, hi;
        const bool ok = ff_row_interval(K_COEF, g, f, st.data(), ct.data(), nth, sin(ph[j]), cos(ph[j]), lo, hi);
        lo_hi[2 * j] = lo; lo_hi[2 * j + 1] = hi;
        if (!ok) { ++*n_fallback; continue; }
        int first = -1, last = -1, count = 0;
        for (int i = 0; i < nth; i++) {
            vec3 d = launch_dir(st[i], ct[i], sin(ph[j]), cos(ph[j])), Pa;
            double t_ap;
            if (!ff_aperture(K_COEF, g, f, d, t_ap, Pa)) continue;
            if (first < 0) first = i;
            last = i; ++count;
        }
        if (count && last - first + 1 != count) return -1;
        if (count == 0 ? hi != lo : (lo != first || hi != last + 1)) ++bad;
    }
    return bad;
}

// the cubic fit exactly as ff_prep_kernel forms it (same entries, same elimination steps)
static void pred_fit(const float* us, const float* vs, const double* Xs, const double* Zs, FfPred& p)
{
    std::vector<double> samples((size_t)FF_PRED_SAMPLES * FF_PRED_LD);
    for (int k = 0; k < FF_PRED_SAMPLES; k++) {
        double* row = samples.data() + (size_t)k * FF_PRED_LD;
        ff_pred_basis<double>((double)us[k], (double)vs[k], row);
        row[FF_PRED_NB] = Xs[k]; row[FF_PRED_NB + 1] = Zs[k];
    }
    ff_pred_fit(samples.data(), FF_PRED_SAMPLES, p);
}

// The fused map's focal-position predictor, as the kernel builds it: FF_PRED_SAMPLES rays spread over the pixel's
// dense list of accepted rays are traced exactly, the affine map is fitted and its worst sample error sets the
// margin.  out4: worst predictor error over ALL accepted rays [mm], worst sample error [mm], rays the predictor
// drops for a grid with high edges gx_max, gy_max [cm], rays it drops although the exact test (ffl:19) keeps them.
API int shadow_ff_pred_check(const double* geo, const double* th, int nth, const double* ph, int nph, double el, double az,
                             double gx_max, double gy_max, double* out4)
{
    FfGeom g = unpack(geo);
    FfFrames f;
    ff_build_frames(K_COEF, g, el, az, f);
    // uniform 51-point E_rec axes [-g_max, g_max]
    const int M = 51;
    std::vector<double> ax(M), ay(M);
    for (int i = 0; i < M; i++) { ax[i] = -gx_max + 2.0 * gx_max * i / (M - 1); ay[i] = -gy_max + 2.0 * gy_max * i / (M - 1); }
    FfGrid G;
    G.gx = ax.data(); G.gy = ay.data(); G.M = M; G.gy_cs = 0; G.gy_rs = 1; G.y_uniform = 1;
    G.gx_max = ax[M - 1]; G.gy_max = ay[M - 1]; G.gx_min = ax[0]; G.gy_min = ay[0];
    G.inv_dx = gx_max > 0 ? (float)((M - 1) / (2.0 * gx_max)) : 0.0f; G.inv_dy = gy_max > 0 ? (float)((M - 1) / (2.0 * gy_max)) : 0.0f;
    std::vector<double> st(nth), ct(nth);
    for (int i = 0; i < nth; i++) { st[i] = sin(th[i]); ct[i] = cos(th[i]); }
    std::vector<int> rays;
    for (int j = 0; j < nph; j++) {
        int lo, hi;
        ff_row_interval(K_COEF, g, f, st.data(), ct.data(), nth, sin(ph[j]), cos(ph[j]), lo, hi);
        for (int i = lo; i < hi; i++) rays.push_back(j * nth + i);
    }
    const int n_acc = (int)rays.size();
    auto trace = [&](int r, float& u, float& v, double& X, double& Z) {
        const int i = r % nth, j = r / nth;
        vec3 d = launch_dir(st[i], ct[i], sin(ph[j]), cos(ph[j])), Pa;
        double t_ap, opl, fx, fz;
        int rst;
        ff_aperture(K_COEF, g, f, d, t_ap, Pa);
        ff_ray_lean<double>(K_COEF, f, d, t_ap, Pa, g.foc_y, opl, fx, fz, rst);
        u = (float)st[i] * (float)cos(ph[j]) * FF_PRED_SCALE; v = (float)st[i] * (float)sin(ph[j]) * FF_PRED_SCALE;
        X = fx - f.centre[0]; Z = fz - f.centre[2];
    };
    FfPred p;
    p.on = 0;
    double worst_sample = 0.0;
    if (n_acc >= 512) {
        float us[FF_PRED_SAMPLES], vs[FF_PRED_SAMPLES];
        double Xs[FF_PRED_SAMPLES], Zs[FF_PRED_SAMPLES];
        for (int k = 0; k < FF_PRED_SAMPLES; k++) {
            const int r = rays[ff_pred_sample_pos(k, n_acc)];
            trace(r, us[k], vs[k], Xs[k], Zs[k]);
        }
        pred_fit(us, vs, Xs, Zs, p);
        for (int k = 0; k < FF_PRED_SAMPLES; k++) {
            float x, z;
            ff_pred_eval(p, us[k], vs[k], x, z);
            worst_sample = fmax(worst_sample, fmax(fabs(x - Xs[k]), fabs(z - Zs[k])));
        }
        ff_pred_set_margin(p, (float)worst_sample);
    }
    double worst = 0.0;
    long dropped = 0, wrong = 0;
    if (p.on)
        for (int r : rays) {
            float u, v, x, z;
            double X, Z;
            trace(r, u, v, X, Z);
            ff_pred_eval(p, u, v, x, z);
            worst = fmax(worst, fmax(fabs(x - X), fabs(z - Z)));
            // rays the predictor calls "beyond the high edge" / "below both low edges" must be so
            const float xc = 0.1f * x, zc = 0.1f * z;
            if (xc > (float)gx_max + p.margin_cm || zc > (float)gy_max + p.margin_cm) {
                ++dropped;
                if (!(div10(X) > gx_max || div10(Z) > gy_max)) ++wrong;
            }
            if (gx_max > 0 && xc < (float)G.gx_min - p.margin_cm && zc < (float)G.gy_min - p.margin_cm && ff_cell(X, Z, G) != 0) ++wrong;
        }
    out4[0] = worst; out4[1] = worst_sample; out4[2] = (double)dropped; out4[3] = (double)wrong;
    return p.on ? n_acc : -n_acc;
}

// The ray selection of the fused map (row intervals, predictor, ff_pred_row_keep) against tracing every accepted ray:
// winner tables (ray index per cell) of the pixel and of its mirror pixel from the kept rays only must equal the
// tables from all rays.  E_rec axes: M points on [g_lo, g_hi] cm.  out: accepted rays, kept rays, cells that differ.
API int shadow_ff_keep_check(const double* geo, const double* th, int nth, const double* ph, int nph, double el, double az, int M,
                             double g_lo, double g_hi, int need_pos, int need_neg, double* out3, int pile_min)
{
    FfGeom g = unpack(geo);
    FfFrames f;
    ff_build_frames(K_COEF, g, el, az, f);
    std::vector<double> ax(M);
    for (int i = 0; i < M; i++) ax[i] = g_lo + (g_hi - g_lo) * i / (M - 1);
    FfGrid G;
    G.gx = ax.data(); G.gy = ax.data(); G.M = M; G.gy_cs = 0; G.gy_rs = 1; G.y_uniform = 1;
    G.gx_max = G.gy_max = ax[M - 1]; G.gx_min = G.gy_min = ax[0];
    G.inv_dx = G.inv_dy = (float)((M - 1) / (g_hi - g_lo));
    std::vector<double> st(nth), ct(nth);
    std::vector<float> stf(nth);
    for (int i = 0; i < nth; i++) { st[i] = sin(th[i]); ct[i] = cos(th[i]); stf[i] = (float)st[i]; }
    std::vector<int> lo(nph), hi(nph), rays;
    for (int j = 0; j < nph; j++) {
        ff_row_interval(K_COEF, g, f, st.data(), ct.data(), nth, sin(ph[j]), cos(ph[j]), lo[j], hi[j]);
        for (int i = lo[j]; i < hi[j]; i++) rays.push_back(j * nth + i);
    }
    const int n_acc = (int)rays.size();
    struct Hit { double X, Z, Zm; };
    auto trace = [&](int r) {
        const int i = r % nth, j = r / nth;
        vec3 d = launch_dir(st[i], ct[i], sin(ph[j]), cos(ph[j])), Pa;
        double t_ap, opl, fx, fz;
        int rst;
        ff_aperture(K_COEF, g, f, d, t_ap, Pa);
        ff_ray_lean<double>(K_COEF, f, d, t_ap, Pa, g.foc_y, opl, fx, fz, rst);
        Hit h = {fx - f.centre[0], fz - f.centre[2], fz - f.centre_mz};
        return h;
    };
    // predictor exactly as ff_prep_kernel builds it
    FfPred p;
    p.on = 0;
    if (n_acc >= 512) {
        float us[FF_PRED_SAMPLES], vs[FF_PRED_SAMPLES];
        double Xs[FF_PRED_SAMPLES], Zs[FF_PRED_SAMPLES];
        for (int k = 0; k < FF_PRED_SAMPLES; k++) {
            const int r = rays[ff_pred_sample_pos(k, n_acc)], i = r % nth, j = r / nth;
            const Hit h = trace(r);
            us[k] = stf[i] * FF_PRED_SCALE * (float)cos(ph[j]); vs[k] = stf[i] * FF_PRED_SCALE * (float)sin(ph[j]);
            Xs[k] = h.X; Zs[k] = h.Z;
        }
        pred_fit(us, vs, Xs, Zs, p);
        double worst = 0.0;
        for (int k = 0; k < FF_PRED_SAMPLES; k++) {
            float x, z;
            ff_pred_eval(p, us[k], vs[k], x, z);
            worst = fmax(worst, fmax(fabs(x - Xs[k]), fabs(z - Zs[k])));
        }
        ff_pred_set_margin(p, (float)worst);
    }
    std::vector<char> keep(nth * (size_t)nph, 0);
    long n_kept = 0;
    const float dz_m = 0.1f * (float)(f.centre[2] - f.centre_mz);
    std::vector<int> KLO((size_t)nph * FF_ROW_K), KCNT((size_t)nph * FF_ROW_K);
    auto keep_pass = [&](const FfPile* pile_in) {
        n_kept = 0;
        for (int j = 0; j < nph; j++) {
            int* klo = &KLO[(size_t)j * FF_ROW_K];
            int* kcnt = &KCNT[(size_t)j * FF_ROW_K];
            if (p.on) {
                FfPile pl;
                if (pile_in) { pl = *pile_in; pl.j = j; }
                ff_pred_row_keep(p, stf.data(), (float)cos(ph[j]), (float)sin(ph[j]), G, need_pos != 0, need_neg != 0, dz_m, lo[j], hi[j], klo,
                                 kcnt, pile_in ? &pl : nullptr);
            } else {
                for (int k = 0; k < FF_ROW_K; k++) { klo[k] = lo[j]; kcnt[k] = 0; }
                kcnt[0] = hi[j] - lo[j];
            }
            for (int k = 0; k < FF_ROW_K; k++) n_kept += kcnt[k];
        }
    };
    keep_pass(nullptr);
    std::vector<int> L2v(4 * (size_t)M, 0);
    if (p.on && pile_min >= 0 && n_kept >= pile_min) {
        // the scout stage exactly as ff_prep_kernel runs it
        std::vector<int> scout(8 * (size_t)M, 0), L1(4 * (size_t)M, 0), sc(16 * (size_t)M, -1);
        const float mg = p.margin_cm;
        const float x_lo = (float)G.gx_min - mg, z_lo = (float)G.gy_min - mg, x_hi = (float)G.gx_max + mg, z_hi = (float)G.gy_max + mg;
        auto pile_line = [&](float X, float Z, float& c) {
            if (X > x_hi || Z > z_hi) return -1;
            const bool left = X < x_lo, below = Z < z_lo;
            if (left == below) return -1;
            c = left ? Z : X;
            return left ? 0 : 1;
        };
        for (int j = 0; j < nph; j++) {
            float ax[4], az[4];
            ff_pred_row_cubic(p.cx, (float)cos(ph[j]), (float)sin(ph[j]), ax);
            ff_pred_row_cubic(p.cz, (float)cos(ph[j]), (float)sin(ph[j]), az);
            for (int k = 0; k < FF_ROW_K; k++) {
                const int l0 = KLO[(size_t)j * FF_ROW_K + k], cnt = KCNT[(size_t)j * FF_ROW_K + k];
                for (int i = l0 + cnt - 1; i >= l0; i -= FF_SCOUT_STEP) {
                    const float sf = stf[i] * FF_PRED_SCALE;
                    const float X = 0.1f * fmaf(fmaf(fmaf(ax[3], sf, ax[2]), sf, ax[1]), sf, ax[0]);
                    const float Z = 0.1f * fmaf(fmaf(fmaf(az[3], sf, az[2]), sf, az[1]), sf, az[0]);
                    float c;
                    if (need_pos) {
                        const int line = pile_line(X, Z, c);
                        if (line >= 0) {
                            int& s = scout[(size_t)line * 2 * M + ff_pile_slot(G, line, c)];
                            s = std::max(s, j * nth + i + 1);
                        }
                    }
                    if (need_neg) {
                        const int line = pile_line(-X, Z + dz_m, c);
                        if (line >= 0) {
                            int& s = scout[(size_t)(2 + line) * 2 * M + ff_pile_slot(G, line, c)];
                            s = std::max(s, (nph - 1 - j) * nth + i + 1);
                        }
                    }
                }
            }
        }
        auto scout_ray = [&](int e, int& i, int& j) {
            const int sidx = scout[e] - 1;
            j = sidx / nth; i = sidx - j * nth;
            if (e >= 4 * M) j = nph - 1 - j;
        };
        auto landings = [&](int cp, int cn, int i, int j, std::vector<int>& L) {
            if (cp >= 0) {
                if (cp % M == 0) L[0 * M + cp / M] = std::max(L[0 * M + cp / M], j * nth + i + 1);
                if (cp / M == 0) L[1 * M + cp % M] = std::max(L[1 * M + cp % M], j * nth + i + 1);
            }
            if (cn >= 0) {
                if (cn % M == 0) L[2 * M + cn / M] = std::max(L[2 * M + cn / M], (nph - 1 - j) * nth + i + 1);
                if (cn / M == 0) L[3 * M + cn % M] = std::max(L[3 * M + cn % M], (nph - 1 - j) * nth + i + 1);
            }
        };
        for (int e = 0; e < 8 * M; e++) {
            if (scout[e] <= 0) continue;
            int i, j;
            scout_ray(e, i, j);
            const Hit h = trace(j * nth + i);
            if (need_pos) sc[e] = ff_cell(h.X, h.Z, G);
            if (need_neg) sc[8 * M + e] = ff_cell(-h.X, h.Zm, G);
            landings(sc[e], sc[8 * M + e], i, j, L1);
        }
        for (int e = 0; e < 8 * M; e++) {
            if (scout[e] <= 0) continue;
            const int t = e / (4 * M), line = (e / (2 * M)) & 1, cell = sc[(size_t)t * 8 * M + e];
            if (cell < 0 || (line == 0 ? cell % M != 0 : cell / M != 0)) continue;
            if (L1[(size_t)(2 * t + line) * M + (line == 0 ? cell / M : cell % M)] != scout[e]) continue;
            int i, j;
            scout_ray(e, i, j);
            landings(sc[e], sc[8 * M + e], i, j, L2v);
        }
        FfPile pile;
        pile.L = L2v.data(); pile.nth = nth; pile.nph = nph; pile.j = 0;
        keep_pass(&pile);
    }
    n_kept = 0;
    for (int j = 0; j < nph; j++)
        for (int k = 0; k < FF_ROW_K; k++) {
            const int kl = KLO[(size_t)j * FF_ROW_K + k], kc = KCNT[(size_t)j * FF_ROW_K + k];
            for (int i = kl; i < kl + kc; i++) {
                if (i < lo[j] || i >= hi[j] || keep[(size_t)j * nth + i]) return -1;      // outside the row or twice
                keep[(size_t)j * nth + i] = 1; ++n_kept;
            }
        }
    std::vector<int> wp_all(M * M, -1), wn_all(M * M, -1), wp_k(M * M, -1), wn_k(M * M, -1);
    for (int r : rays) {
        const int i = r % nth, j = r / nth, rm = (nph - 1 - j) * nth + i;
        const Hit h = trace(r);
        const int cp = need_pos ? ff_cell(h.X, h.Z, G) : -1, cn = need_neg ? ff_cell(-h.X, h.Zm, G) : -1;
        if (cp >= 0) { wp_all[cp] = std::max(wp_all[cp], r); if (keep[r]) wp_k[cp] = std::max(wp_k[cp], r); }
        if (cn >= 0) { wn_all[cn] = std::max(wn_all[cn], rm); if (keep[r]) wn_k[cn] = std::max(wn_k[cn], rm); }
    }
    long diff = 0;
    for (int c = 0; c < M * M; c++) diff += (wp_all[c] != wp_k[c]) + (wn_all[c] != wn_k[c]);
    out3[0] = n_acc; out3[1] = (double)n_kept; out3[2] = (double)diff;
    return p.on;
}

// one far-field pixel, zero-error single-panel model (fused-map semantics)
API int shadow_ff_pixel(const double* geo, const double* th, int nth, const double* ph, int nph, double el, double az,
                        const double* gx_full, const double* gy, const double* e_rec, int M, double k, double* re_im,
                        int* n_cells)
{
    FfGeom g = unpack(geo);
    FfFrames f;
    ff_build_frames(K_COEF, g, el, az, f);
    const int n = nth * nph;
    int agg = 0;
    const FfGrid G = make_grid(gx_full, gy, M);
    std::vector<int> win(M * M, -1);
    std::vector<double> opl(M * M, 0.0);
    for (int r = 0; r < n; r++) {
        const int i = r % nth, j = r / nth;
        vec3 d = launch_dir(sin(th[i]), cos(th[i]), sin(ph[j]), cos(ph[j])), Pa;
        double t_ap;
        if (!ff_aperture(K_COEF, g, f, d, t_ap, Pa)) continue;
        double ropl, rfx, rfz;
        int rst;
        ff_ray_lean<double>(K_COEF, f, d, t_ap, Pa, g.foc_y, ropl, rfx, rfz, rst);
        agg |= rst;
        if (ropl == 0.0) continue;
        const int c = ff_cell(rfx - f.centre[0], rfz - f.centre[2], G);
        if (c < 0) continue;
        if (r > win[c]) { win[c] = r; opl[c] = ropl; }
    }
    double re = 0, im = 0;
    int cnt = 0;
    for (int c = 0; c < M * M; c++) {
        if (win[c] < 0) continue;
        cnt++;
        double cs, sn;
        cis_large(k * opl[c] / 1e3, cs, sn);
        re += cs * e_rec[2 * c] - sn * e_rec[2 * c + 1];
        im += cs * e_rec[2 * c + 1] + sn * e_rec[2 * c];
    }
    re_im[0] = re / ((double)M * M); re_im[1] = im / ((double)M * M);
    if (n_cells) *n_cells = cnt;
    return agg;
}

// regrid: the separable not-a-knot bicubic resample the CUDA kernels run (rows, then columns)
API void shadow_regrid(const double* field, int n_old, int n_new, double* out)
{
    std::vector<double> tmp((size_t)n_old * n_new), M(2 * (size_t)n_old);
    for (int r = 0; r < n_old; r++)
        spline_line(field + (size_t)r * n_old, 1, n_old, M.data(), M.data() + n_old, n_new, tmp.data() + (size_t)r * n_new, 1);
    for (int c = 0; c < n_new; c++)
        spline_line(tmp.data() + c, n_new, n_old, M.data(), M.data() + n_old, n_new, out + c, n_new);
}

// searchsorted_lean / axis_is_lean (the lean binning of ff_map_kernel): the axis is staged with its two sentinels
// exactly as the kernel does it.  out[k] = index of v[k]; returns axis_is_lean(a).
API int shadow_searchsorted_lean(const double* a, int n, const double* v, int nv, int* out)
{
    const bool lean = axis_is_lean(a, n, 1);
    std::vector<double> s((size_t)n + 2);
    s[0] = -INFINITY; s[(size_t)n + 1] = INFINITY;
    for (int i = 0; i < n; i++) s[(size_t)i + 1] = a[i];
    const double sx = a[n - 1] - a[0];
    const float inv = (n > 1 && sx > 0) ? (float)((n - 1) / sx) : 0.0f;
    for (int k = 0; k < nv; k++) out[k] = searchsorted_lean(s.data() + 1, (float)n, v[k], a[0], inv);
    return lean ? 1 : 0;
}

// ff_cell_class (ff_map_interp_kernel): class of predicted positions (ux, uy) in cell units on an M x M table
API void shadow_cell_class(const float* ux, const float* uy, int n, int M, float tol_x, float tol_y, int* out)
{
    for (int k = 0; k < n; k++) out[k] = ff_cell_class(ux[k], uy[k], (float)(M - 1) + 0.75f, tol_x, tol_y, M);
}
