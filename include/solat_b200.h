/*
 * solat_b200.h -- C ABI of libsolat_b200.so, the B200 (sm_100a) implementation of the
 * solat-optics beam-simulation hot path.
 *
 * The reference (McMahonCosmologyGroup/solat-optics) is pure Python and has no FFI layer;
 * its boundary for this path is the Python call surface listed below.  Each entry point
 * names the reference interface it replaces (file:line under the reference tree).  The
 * Python package solat_optics_b200 binds these with ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; all floating point is IEEE double; lengths in mm, angles in rad,
 *    wavenumbers in rad/m exactly as in the reference.
 *  - "host" entry points take HOST pointers (pageable or pinned), do the H2D/D2H copies
 *    themselves and return after the result is in the output buffer.
 *  - "device" entry points take DEVICE pointers, enqueue on the given CUDA stream
 *    (a cudaStream_t passed as void*; NULL = legacy default stream) and return immediately.
 *  - every function returns 0 on success and a negative solat_err on failure;
 *    solat_last_error() gives a thread-local message.
 *  - per-ray status bits (also OR-reduced per map): 1 = root outside the reference's brentq
 *    bracket (the reference raises ValueError there), 2 = Newton not converged, 4 = NaN path,
 *    8 = far-field path length outside the fixed-point range of the last-write-wins key.
 *  - no global mutable state except the per-thread error string and a launch counter.
 */
#ifndef SOLAT_B200_H
#define SOLAT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SOLAT_ABI_VERSION 1
#if defined(__GNUC__)
#define SOLAT_API __attribute__((visibility("default")))
#else
#define SOLAT_API
#endif

enum solat_err {
    SOLAT_OK = 0,
    SOLAT_ERR_ARG = -1,     /* bad argument */
    SOLAT_ERR_CUDA = -2,    /* CUDA runtime error, see solat_last_error() */
    SOLAT_ERR_NODEV = -3    /* no usable sm_100 device */
};
enum solat_ray_status { SOLAT_ST_BRACKET = 1, SOLAT_ST_NOCONV = 2, SOLAT_ST_NAN = 4, SOLAT_ST_RANGE = 8 };

typedef struct solat_ctx solat_ctx;         /* one per device; owns scratch buffers */
typedef struct solat_nf_plan solat_nf_plan; /* near-field map plan (tables resident in HBM) */
typedef struct solat_ff_plan solat_ff_plan; /* far-field map plan */

/* ot_geo.LatGeo (solat_optics/ot_geo.py:12-58): the attributes the tracers read */
typedef struct {
    double n_si, n_vac; /* refractive indices (3.4, 1.0) */
    double th_fwhp;     /* feed full width at half power [rad] */
    double y_ap;        /* source plane / aperture y [mm] */
} solat_nf_geom;

/* latrt_geo.initialize_telescope_geometry (solat_optics/latrt_geo.py:7-49) */
typedef struct {
    double th_1, th2;   /* mirror tilts; th2 = -pi/2 - atan(1/3) */
    double rotc_mm[3];  /* (x,y,z)_rotc * 1e3 */
    double src_mm[3];   /* (x,y,z)_tow * 1e3 */
    double foc_y_mm;    /* rx_y * 1e3 (solat_apert_field.py:582) */
} solat_ff_geom;

SOLAT_API int solat_abi_version(void);
SOLAT_API const char *solat_last_error(void);
SOLAT_API int solat_device_count(void);
SOLAT_API int solat_ctx_create(int device, solat_ctx **ctx);
SOLAT_API void solat_ctx_destroy(solat_ctx *ctx);
SOLAT_API int solat_ctx_info(const solat_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, long long *hbm_bytes);
/* number of kernels this library has launched in the calling process so far */
SOLAT_API long long solat_launch_count(void);
/* DFMA-issue microbenchmark: measured FP64 FMA throughput of the device in TFLOP/s (roofline denominator) */
SOLAT_API int solat_fp64_peak(solat_ctx *ctx, int repeats, double *tflops);
/* accuracy probe of the device fast-math primitives on x[n] (host pointers).  out10 is (10, n): reciprocal with
 * 0..3 Newton steps, reciprocal square root with 0..1 polish steps, cos and sin of the large-argument reducer,
 * cos and sin of the mid-range (|x| < 1e5) routine */
SOLAT_API int solat_selftest_math(solat_ctx *ctx, const double *x, int n, double *out10);

/* ---------------- near field: reference-shaped per-ray arrays (host pointers) ----------------
 * theta[nth], phi[nph] are the 1-D launch axes; ray ii uses theta[ii % nth], phi[ii / nth]
 * (np.meshgrid(theta, phi) raveled, ray_trace_int.py:596-598).  out is (rows, nth*nph) row-major.
 * nph == 0 selects PAIRED mode for the per-ray array functions (not the maps): phi then has nth
 * entries and ray ii uses (theta[ii], phi[ii]) -- what trace_rays(ii) of the reference classes needs.
 * status may be NULL; otherwise int[number of rays].                                             */

/* replaces ray_trace_int.rx_to_lyot_model(p_rx, tele_geo, plotOpts).output(), rti:579-885; 20 rows */
SOLAT_API int solat_nf_rx_trace(solat_ctx *ctx, const solat_nf_geom *g, const double p_rx[3], const double *theta, int nth,
                      const double *phi, int nph, double *out20, int *status);
/* replaces ray_trace_int.source_to_lyot_model(p_source, ...).output(), rti:888-1382; 26 rows */
SOLAT_API int solat_nf_source_trace(solat_ctx *ctx, const solat_nf_geom *g, const double p_src[3], const double *theta, int nth,
                          const double *phi, int nph, double *out26, int *status);
/* replaces ray_trace_int.aperature_fields(P_rx, ...).output(), rti:33-576; 32 rows */
SOLAT_API int solat_nf_toy_trace(solat_ctx *ctx, const solat_nf_geom *g, const double p_rx[3], const double *theta, int nth,
                       const double *phi, int nph, double *out32, int *status);

/* replaces the four interp2d(kind="cubic") resamples of ray_trace_int.regrid, rti:1403-1425: interpolating bicubic
 * (not-a-knot) spline of field (n_old, n_old), given on linspace(-70, 70, n_old) in both axes, evaluated on
 * linspace(-70, 70, n_new); out is (n_new, n_new).  Host pointers.                                          */
SOLAT_API int solat_regrid_bicubic(solat_ctx *ctx, const double *field, int n_old, int n_new, double *out);

/* ---------------- near field: fused holography map --------------------------------------------
 * replaces the worker body scripts/holo_sim.py:157-185 for a whole list of source positions:
 * per pixel  sum_rays  rx_w * exp(i k (rx_opl + L_source)/1e3)  ->  amp = |.|, phase = atan2.
 * rx_opl / rx_w [nth*nph]: receiver path length and (feed taper x Lyot mask) in ray-index space,
 * i.e. rows 3 and 4 of the regridded rx field (rti:1385-1440, holo_sim.py:168-180); rays with
 * rx_w == 0 are culled before tracing.  k_per_m[nfreq]: several wavenumbers share one trace.
 * amp/phase are [nfreq][npix].  The _multi form takes nrx receiver tables, rx_opl / rx_w [nrx][nth*nph]
 * (receiver-offset sweeps share the source trace too); amp/phase are then [nrx][nfreq][npix].     */
SOLAT_API int solat_nf_plan_create(solat_ctx *ctx, const solat_nf_geom *g, const double *theta, int nth, const double *phi,
                         int nph, const double *rx_opl, const double *rx_w, const double *k_per_m, int nfreq,
                         solat_nf_plan **plan);
SOLAT_API int solat_nf_plan_create_multi(solat_ctx *ctx, const solat_nf_geom *g, const double *theta, int nth,
                                         const double *phi, int nph, const double *rx_opl, const double *rx_w, int nrx,
                                         const double *k_per_m, int nfreq, solat_nf_plan **plan);
SOLAT_API void solat_nf_plan_destroy(solat_nf_plan *plan);
SOLAT_API int solat_nf_map_host(solat_nf_plan *plan, const double *src_xz, int npix, double *amp, double *phase, int *status);
SOLAT_API int solat_nf_map_device(solat_nf_plan *plan, const double *d_src_xz, int npix, double *d_amp, double *d_phase,
                        void *stream);
/* Device-pointer map that also reports its ray status on the device: d_status is an 8-byte device slot, the OR of the
 * launch's ray status bits as an int in its low word (the multi-GPU path gathers it with the results).        */
SOLAT_API int solat_nf_map_device_st(solat_nf_plan *plan, const double *d_src_xz, int npix, double *d_amp, double *d_phase,
                           void *d_status, void *stream);
/* Receiver fields of nrx receiver positions p_rx[nrx][3] in one launch: rows 3 (path length) and 4 (feed-taper
 * amplitude) of rx_to_lyot_model(...).output() (ray_trace_int.py:848-853), opl/amp [nrx][nth*nph] -- the receiver side
 * of the per-pixel integral (scripts/holo_sim.py:157-176) for a receiver-offset sweep.                          */
SOLAT_API int solat_nf_rx_fields(solat_ctx *ctx, const solat_nf_geom *g, const double *p_rx, int nrx, const double *theta, int nth,
                       const double *phi, int nph, double *opl, double *amp, int *status);
/* (near-field maps share a trace between the source positions (+-x, +-z) the same way when the fan is symmetric
 * in phi and theta about pi/2: a position on the negative side of an axis is computed as a mirror image)
 * traced rays per pixel after culling (for FLOP accounting) */
SOLAT_API int solat_nf_plan_active_rays(const solat_nf_plan *plan);
/* number of output channels nrx * nfreq */
SOLAT_API int solat_nf_plan_channels(const solat_nf_plan *plan);
/* OR of the ray status bits of all device-side maps run on this plan since the last call (synchronises) */
SOLAT_API int solat_nf_plan_status(solat_nf_plan *plan, int *status);

/* ---------------- far field: reference-shaped per-ray arrays (host pointers) ------------------ */

/* replaces solat_apert_field.ray_mirror_pts(tele_geo, theta, phi, el, az), saf:307-567; out (9, nth*nph) */
SOLAT_API int solat_ff_ray_mirror_pts(solat_ctx *ctx, const solat_ff_geom *g, const double *theta, int nth, const double *phi,
                            int nph, double el, double az, double *out9, int *status);
/* replaces solat_apert_field.aperature_fields_from_panel_model(...), saf:570-1009; out (17, nth*nph).
 * focal_centre = rxmirror[6:9,0].  pan_idx1/2[nth*nph]: column of pan1/pan2 of the first panel whose
 * (row,col) matches the ray's identify_panel ids, or -1 (the host side of saf:609-628,775-779; panel
 * identification itself belongs to the external pan_mod).  pan1 (10,np1), pan2 (10,np2) row-major
 * [row,col,a,b,c,d,e,f,x0,y0] (saf:667-674).                                                      */
SOLAT_API int solat_ff_apert_fields(solat_ctx *ctx, const solat_ff_geom *g, const double *theta, int nth, const double *phi,
                          int nph, const double focal_centre[3], const int *pan_idx1, const int *pan_idx2,
                          const double *pan1, int np1, const double *pan2, int np2, double el, double az,
                          double *out17, int *status);
/* replaces far_field_latrt.project_to_data(sim, data_x, data_y, tele_geo), ffl:9-32.
 * sim (17, n_pts); data_x, data_y (M, M); out_reim (M, M) complex128 as interleaved re,im.       */
SOLAT_API int solat_ff_project(solat_ctx *ctx, const double *sim17, int n_pts, const double *data_x, const double *data_y, int M,
                     double k, double *out_reim);

/* ---------------- far field: fused beam map ----------------------------------------------------
 * replaces the worker body scripts/ff_sim.py:204-221 for a list of (az, el) pixels, with the
 * single-panel / zero-error panel model (adjuster offsets are identically zero in ff_sim.py:61-62):
 * per pixel  sum_cells  exp(i k L_winner/1e3) * e_rec[cell] / M^2 , winner = highest ray index
 * landing in the cell (last write wins, ffl:28).  azel is [npix][2] = (az, el).
 * How the kernels spend less than the reference: rays the projection would skip
 * (ffl:19) are not traced, and when the telescope, the fan (phi symmetric about pi/2) and therefore the
 * physics are symmetric under x -> -x, the pixels (az, el) and (-az, el) of one call share one trace;
 * a pixel's value does not depend on which other pixels are in the call.  Square uniform fans of 400^2
 * rays and more go through the interpolated-cell kernel: a lattice of rays and the winner of every cell
 * are traced exactly, every other listed ray only has its cell worked out from interpolated focal
 * positions (maps bit-identical to the exact walk; every winner is checked against its prediction and an
 * item that fails is redone exactly).  A ray that is never traced reports no status; the environment
 * switch SOLAT_FF_PREFILTER=0 traces and checks every accepted ray.  Calls on one plan must be
 * stream-ordered by the caller.                                                                   */
SOLAT_API int solat_ff_plan_create(solat_ctx *ctx, const solat_ff_geom *g, const double *theta, int nth, const double *phi,
                         int nph, const double *data_x, const double *data_y, const double *e_rec_reim, int M,
                         double k, solat_ff_plan **plan);
SOLAT_API void solat_ff_plan_destroy(solat_ff_plan *plan);
SOLAT_API int solat_ff_map_host(solat_ff_plan *plan, const double *azel, int npix, double *amp, double *phase, int *status);
SOLAT_API int solat_ff_map_device(solat_ff_plan *plan, const double *d_azel, int npix, double *d_amp, double *d_phase,
                        void *stream);
/* as solat_nf_map_device_st */
SOLAT_API int solat_ff_map_device_st(solat_ff_plan *plan, const double *d_azel, int npix, double *d_amp, double *d_phase,
                           void *d_status, void *stream);
SOLAT_API int solat_ff_plan_status(solat_ff_plan *plan, int *status);

/* Far-field map with panel models (ff_sim.py:61-70,204-221 with non-zero adjuster offsets), two steps around the
 * host's pan_mod.identify_panel (external holosim-ml; its panel layout is not part of the reference):
 * 1. solat_ff_map_mirror_xy_host: for every pixel the coordinates the reference hands to identify_panel
 *    (saf:596-612): xy[(pix*4 + k)*n_rays + r], k = 0,1: M2 hit x, z; k = 2,3: M1 hit x, y (telescope frame,
 *    error-free mirrors = rows 0,2,3,4 of ray_mirror_pts' output; 0 where the ray misses the aperture).
 * 2. solat_ff_map_panels_host: the fused map where ray r of pixel pix sits on column pan_idx1[pix*n_rays + r] of
 *    pan1 and pan_idx2[...] of pan2 ((10, np) row-major [row,col,a,b,c,d,e,f,x0,y0], saf:667-674); a negative
 *    index drops the ray (saf:626-628, :805-809).                                                             */
SOLAT_API int solat_ff_map_mirror_xy_host(solat_ff_plan *plan, const double *azel, int npix, double *xy, int *status);
SOLAT_API int solat_ff_map_panels_host(solat_ff_plan *plan, const double *azel, int npix, const short *pan_idx1,
                             const short *pan_idx2, const double *pan1, int np1, const double *pan2, int np2,
                             double *amp, double *phase, int *status);

/* The same map with the panels identified ON THE DEVICE (no per-ray round trip through the host).  A panel layout is
 * what an identify_panel that loops over rectangular panels computes (the external pan_mod's is not part of the
 * reference; call sites saf:596-617): panel k of a mirror covers [x0, x1) x [y0, y1) in the coordinates the reference
 * hands to identify_panel -- M1: (x, y - y_cent_m1), M2: (x, z) of the ideal-mirror hit, telescope frame, mm -- a later
 * panel overrides an earlier one, a point in no panel sits on no panel.  model_col[k] = column of the mirror's panel
 * model whose (row, col) ids are panel k's (first match, saf:626,805), -1 = the panel is not in the model: rays there
 * are dropped (saf:626-628,805-809).  Per ray the kernel runs the reference's two passes (ff_sim.py:208-209): ideal
 * trace -> panels -> trace with the panels' error polynomials.                                                  */
typedef struct {
    const double *rect;    /* [n][4]: x0, x1, y0, y1 */
    const int *model_col;  /* [n] */
    int n;
} solat_panel_layout;
SOLAT_API int solat_ff_map_panel_layout_host(solat_ff_plan *plan, const double *azel, int npix, const solat_panel_layout *m1,
                                   const solat_panel_layout *m2, const double *pan1, int np1, const double *pan2, int np2,
                                   double *amp, double *phase, int *status);

#ifdef __cplusplus
}
#endif
#endif /* SOLAT_B200_H */
