// Experiment: how accurate is each stage of the mirror Newton iteration on the default far-field fan?
// g++ -O2 -std=c++17 -o /tmp/np tools/probes/newton_probe.cpp && /tmp/np
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../solat_optics_b200/csrc/solat_math.cuh"
using namespace solat;

template <int M>
static void probe(const vec3& P, const vec3& D, double* stats)
{
    double t = -P.z * rcp_step(D.z);
    const double t0 = t;
    {
        const float xf = (float)fma(D.x, t, P.x), yf = (float)fma(D.y, t, P.y);
        float zf, zxf, zyf;
        mirror_sag_f<M, float>(xf, yf, zf, zxf, zyf);
        const float fpf = (float)D.z - fmaf(zxf, (float)D.x, zyf * (float)D.y);
        t += (double)(zf / fpf);
    }
    const double t1 = t;
    double ts[6];
    double z, zx, zy;
    for (int it = 0; it < 6; it++) {
        const double x = fma(D.x, t, P.x), y = fma(D.y, t, P.y);
        mirror_sag<M, double>(K_COEF, x, y, z, zx, zy);
        const double f = fma(D.z, t, P.z) - z, fp = D.z - fma(zx, D.x, zy * D.y);
        t -= f / fp;
        ts[it] = t;
    }
    const double tc = ts[5];
    stats[0] = fmax(stats[0], fabs(t0 - tc));
    stats[1] = fmax(stats[1], fabs(t1 - tc));
    stats[2] = fmax(stats[2], fabs(ts[0] - tc));
    stats[3] = fmax(stats[3], fabs(ts[1] - tc));
}

int main()
{
    FfGeom g;
    g.th_1 = atan(0.5); g.th2 = -M_PI / 2 - atan(1.0 / 3.0);
    g.rotc[0] = 0; g.rotc[1] = -7200; g.rotc[2] = 0;
    g.src[0] = 0; g.src[1] = -7200; g.src[2] = 1e8;
    g.foc_y = 1320;
    const int n = 100;
    for (double az : {0.0, 1.4544e-3, 3e-3}) {
        FfFrames f;
        ff_build_frames(K_COEF, g, 7.2e-5 + az, az, f);
        double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) {
                const double th = M_PI - 5e-5 + 1e-4 * i / (n - 1), ph = M_PI * j / (n - 1);
                vec3 d = launch_dir(sin(th), cos(th), sin(ph), cos(ph)), Pa;
                double t_ap;
                if (!ff_aperture(K_COEF, g, f, d, t_ap, Pa)) continue;
                const vec3 p1 = apply(f.t1, Pa), d1 = apply_lin(f.t1.a, d);
                probe<1>(p1, d1, s1);
                double t1;
                vec3 H, N;
                const vec3 none = mk(0.0, 0.0, 0.0);
                mirror_hit<1, false, double>(K_COEF, p1, d1, 0.0, 1e4, nullptr, none, none, t1, H, N);
                const vec3 p2 = apply(f.t21, H), d2 = apply_lin(f.b12, reflect(d1, N));
                probe<2>(p2, d2, s2);
            }
        printf("az %.4g  M1: plane %.3g coarse %.3g upd1 %.3g upd2 %.3g | M2: plane %.3g coarse %.3g upd1 %.3g upd2 %.3g\n", az,
               s1[0], s1[1], s1[2], s1[3], s2[0], s2[1], s2[2], s2[3]);
    }
    return 0;
}
