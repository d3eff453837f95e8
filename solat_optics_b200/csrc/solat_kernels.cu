// solat_kernels.cu -- sm_100a kernels and the C ABI (include/solat_b200.h) of libsolat_b200.so.
//
// Kernels (all FP64 SIMT; the path has no dense contraction, so no tensor cores -- DESIGN.md):
//   nf_rows_kernel<KIND>   thread-per-ray dump of the reference's .output() arrays (rx / source / toy)
//   nf_rx_fields_kernel    path length and taper of every ray for a list of receiver positions (offset sweeps)
//   nf_first_leg_kernel,   per-plan tables of a near-field map: the pixel-independent first leg of every active ray,
//   nf_phasor_kernel,      the receiver phasors per (receiver table, wavenumber), and for a mirror-symmetric plan the
//   nf_phasor_mirror_kernel  phasors of every ray's three mirror rays next to its own
//   nf_map_kernel<NF,NT,MODE>  fused source trace + phase + aperture integral: one CTA per (work item, ray tile), a
//                          work item being a source position, the mirror pair (+-x, z) or the quad (+-x, +-z); rays
//                          culled by the receiver weight; warp-shuffle + shared-memory block reduce
//   nf_phase_kernel,       sweeps over more than 16 receiver tables: exp(i k L) of every (quad item, ray), traced once,
//   nf_contract_kernel     and its complex matrix product with all the phasor tables (register-tiled FP64, 64 x 64 per CTA)
//   nf_finalize_kernel     fixed-order sum of tile partials -> amp, phase
//   ff_rows_kernel<STAGE>  thread-per-ray dump of ray_mirror_pts / aperature_fields_from_panel_model
//   ff_project_*           last-write-wins scatter (atomicMax on the ray index) + phasor fill
//   ff_pair_*_kernel       pixel list -> work items (a pixel or the mirror pair (+az, el), (-az, el)); also used for
//                          the near-field source positions (a second stage there pairs the pairs into quads)
//   ff_prep_kernel         per-item set-up of the fused far-field map: frames, the rays that have to be traced as up
//                          to four index intervals per phi row (aperture, focal-position predictor)
//   ff_map_kernel<PANEL,T,B,LEAN>  fused far-field map, the exact walk: one CTA per work item; every warp traces its
//                          share of the item's dense ray list with full lanes; winner tables (ray index, OPL) in shared
//                          memory; winners-only phasor sums (LEAN: the binning on shared-memory addresses only)
//   ff_map_interp_kernel   fused far-field map of large fans: a lattice of rays traced exactly, the cell of every
//                          listed ray from its interpolated focal position, exact traces for the winner of every cell
//                          (each one a probe of its prediction); flagged items are redone by ff_map_kernel
//   ff_mirror_xy_kernel    mirror coordinates of every (pixel, ray) for the host's identify_panel
//   fp64_peak_kernel       DFMA issue-rate microbenchmark (roofline denominator)
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/solat_b200.h"
#include "solat_math.cuh"

using namespace solat;

// launch-shape tunables (tools/variant_bench.py A/B-tests them on the B200)
#ifndef SOLAT_FF_T
#define SOLAT_FF_T 256      // threads per far-field CTA while three CTAs fit on an SM (E_rec grids up to ~94 x 94)
#endif
#ifndef SOLAT_FF_MINB
#define SOLAT_FF_MINB 3     // resident far-field CTAs per SM the register allocation is capped for
#endif
#ifndef SOLAT_NF_T
#define SOLAT_NF_T 256      // threads per near-field CTA
#endif
#ifndef SOLAT_NF_MINB
#define SOLAT_NF_MINB 3
#endif

// ------------------------------------------------------------------------------------------------
// host-side plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* what, const char* detail = "")
{
    g_err = std::string(what) + (detail[0] ? ": " : "") + detail;
    return code;
}
#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return fail(SOLAT_ERR_CUDA, #call, cudaGetErrorString(e_)); \
    } while (0)
#define LAUNCHED()                                                                \
    do {                                                                          \
        g_launches.fetch_add(1, std::memory_order_relaxed);                       \
        CK(cudaGetLastError());                                                   \
    } while (0)

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync on one private stream per device, pool
// kept warm): a plan is created and destroyed per map by the Python layer, and plain cudaMalloc / cudaFree cost more
// than the map kernels themselves.  Allocation synchronises the private stream, so the pointer is valid on every
// stream; a buffer is only released after the work that used it was synchronised (host entries synchronise their
// stream before returning, plan destruction synchronises the device).
static cudaStream_t pool_stream()
{
    static cudaStream_t streams[64] = {};
    static std::atomic<bool> ready[64] = {};
    static std::mutex mu;
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (!ready[d].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(mu);
        if (!ready[d].load(std::memory_order_relaxed)) {
            cudaStreamCreateWithFlags(&streams[d], cudaStreamNonBlocking);
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            ready[d].store(true, std::memory_order_release);
        }
    }
    return streams[d];
}

// device buffer with RAII, bound to the current device
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int reserve(size_t count)
    {
        if (count <= n) return 0;
        cudaStream_t ps = pool_stream();
        if (p) {
            cudaDeviceSynchronize();          // growing a buffer earlier launches may still read
            cudaFreeAsync(p, ps);
        }
        p = nullptr; n = 0;
        CK(cudaMallocAsync(&p, count * sizeof(T), ps));
        CK(cudaStreamSynchronize(ps));
        n = count;
        return 0;
    }
    int upload(const T* h, size_t count, cudaStream_t s)
    {
        int rc = reserve(count);
        if (rc) return rc;
        CK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
        return 0;
    }
    ~DevBuf() { if (p) cudaFreeAsync(p, pool_stream()); }
};

// pinned host staging buffer (one D2H per host-entry call instead of three)
struct PinnedBuf {
    double* p = nullptr;
    size_t n = 0;
    int reserve(size_t count)
    {
        if (count <= n) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        CK(cudaHostAlloc(&p, count * sizeof(double), cudaHostAllocDefault));
        n = count;
        return 0;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

struct solat_ctx {
    int device;
    cudaDeviceProp prop;
    cudaStream_t stream;   // used by the host entry points
    // staging of the host-buffer map entries (calls on one context are synchronous, one at a time): pixel list in,
    // [amp | phase | status word] out, and its pinned mirror for a single D2H
    DevBuf<double> stage_in, stage_out;
    PinnedBuf pin;
    std::mutex stage_mu;   // host-buffer map entries of one context run one at a time
};

struct AngleTabs {
    const double *st, *ct, *sp, *cp;   // sin/cos of theta[nth], phi[nph]
    const float *stf, *ctf, *spf, *cpf; // single-precision copies (far-field aperture prefilter)
    const double2 *tsc, *psc;          // (sin, cos) of theta[nth] / phi[nph] interleaved: one 16-byte load per axis in the hot loop
    int nth, nph;                      // nph == 0: paired mode, ray ii uses theta[ii], phi[ii] (phi has nth entries)
    unsigned magic;                    // ceil(2^32 / nth): r / nth == __umulhi(r, magic) (+1 fix-up) for r < 2^32 / nth
    __host__ __device__ int rays() const { return nph == 0 ? nth : nth * nph; }
};
struct AngleTabsOwner {
    DevBuf<double> buf;
    DevBuf<float> fbuf;
    DevBuf<double2> pbuf;
    AngleTabs tabs;
    int init(const double* theta, int nth, const double* phi, int nph_arg, cudaStream_t s)
    {
        const int nph = nph_arg == 0 ? nth : nph_arg;
        std::vector<double> h(2 * (size_t)nth + 2 * (size_t)nph);
        for (int i = 0; i < nth; i++) { h[i] = sin(theta[i]); h[nth + i] = cos(theta[i]); }
        for (int j = 0; j < nph; j++) { h[2 * nth + j] = sin(phi[j]); h[2 * nth + nph + j] = cos(phi[j]); }
        std::vector<float> hf(h.size());
        for (size_t i = 0; i < h.size(); i++) hf[i] = (float)h[i];
        std::vector<double2> hp((size_t)nth + nph);
        for (int i = 0; i < nth; i++) hp[i] = make_double2(h[i], h[nth + i]);
        for (int j = 0; j < nph; j++) hp[nth + j] = make_double2(h[2 * nth + j], h[2 * nth + nph + j]);
        int rc = buf.upload(h.data(), h.size(), s);
        if (!rc) rc = fbuf.upload(hf.data(), hf.size(), s);
        if (!rc) rc = pbuf.upload(hp.data(), hp.size(), s);
        if (rc) return rc;
        CK(cudaStreamSynchronize(s));   // h, hf go out of scope
        tabs.stf = fbuf.p; tabs.ctf = fbuf.p + nth; tabs.spf = fbuf.p + 2 * nth; tabs.cpf = fbuf.p + 2 * nth + nph;
        tabs.tsc = pbuf.p; tabs.psc = pbuf.p + nth;
        tabs.st = buf.p; tabs.ct = buf.p + nth; tabs.sp = buf.p + 2 * nth; tabs.cp = buf.p + 2 * nth + nph;
        tabs.nth = nth; tabs.nph = nph_arg;
        const unsigned long long rays = (unsigned long long)nth * (unsigned long long)nph;
        tabs.magic = (rays * (unsigned long long)nth < 0x100000000ull && nth > 1)
                         ? (unsigned)((0x100000000ull + (unsigned long long)nth - 1) / (unsigned long long)nth) : 0u;
        return 0;
    }
};

__device__ __forceinline__ void ray_ij(const AngleTabs& a, int r, int& i, int& j)
{
    if (a.magic) {
        j = (int)__umulhi((unsigned)r, a.magic);
        i = r - j * a.nth;
        if (i < 0) { i += a.nth; --j; }
    } else {
        j = r / a.nth;
        i = r - j * a.nth;
    }
    if (a.nph == 0) { i = r; j = r; }
}

__device__ __forceinline__ vec3 ray_dir(const AngleTabs& a, int r)
{
    // meshgrid(theta, phi) raveled: theta is the fast index.  r / nth by multiply-high: the estimate is at most
    // one too large while r * nth < 2^32 (the compare fixes it up); larger fans take the plain division.
    int j, i;
    if (a.magic) {
        j = (int)__umulhi((unsigned)r, a.magic);
        i = r - j * a.nth;
        if (i < 0) { i += a.nth; --j; }
    } else {
        j = r / a.nth;
        i = r - j * a.nth;
    }
    if (a.nph == 0) { i = r; j = r; }
    return launch_dir(__ldg(a.st + i), __ldg(a.ct + i), __ldg(a.sp + j), __ldg(a.cp + j));
}

__device__ __forceinline__ vec3t<d2> ray_dir2(const AngleTabs& a, int r0, int r1)
{
    const vec3 u = ray_dir(a, r0), v = ray_dir(a, r1);
    return mk<d2>(d2(u.x, v.x), d2(u.y, v.y), d2(u.z, v.z));
}

// ------------------------------------------------------------------------------------------------
// near field: per-ray dumps
// ------------------------------------------------------------------------------------------------
enum { NF_RX = 0, NF_SOURCE = 1, NF_TOY = 2 };
struct NfRowsParams {
    AngleTabs ang;
    const double *theta, *phi;   // raw angles (rx taper needs them)
    double p[3];
    double n_si, n_vac, y_ap, inv_2sig2, th_mean, ph_mean;
    double* out;
    int* status;
};

template <int KIND>
__global__ void __launch_bounds__(128) nf_rows_kernel(NfRowsParams q)
{
    const int n = q.ang.rays();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const vec3 p0 = mk(q.p[0], q.p[1], q.p[2]);
    const vec3 d = ray_dir(q.ang, r);
    constexpr int ROWS = KIND == NF_RX ? 20 : (KIND == NF_SOURCE ? 26 : 32);
    double rows[ROWS];
    int st;
    if (KIND == NF_RX) {
        NfRxRay ray;
        nf_rx_ray(K_COEF, p0, d, q.n_si, q.n_vac, ray);
        int j = r / q.ang.nth, i = r - j * q.ang.nth;
        if (q.ang.nph == 0) { i = r; j = r; }
        pack_nf_rx_rows(ray, p0, q.theta[i] - q.th_mean, q.phi[j] - q.ph_mean, q.inv_2sig2, rows);
        st = ray.status;
    } else if (KIND == NF_SOURCE) {
        NfSourceRay ray;
        nf_source_ray(K_COEF, p0, d, q.n_si, q.n_vac, ray);
        pack_nf_source_rows(ray, p0, q.n_si, q.n_vac, rows);
        st = ray.status;
    } else {
        NfToyRay ray;
        nf_toy_ray(K_COEF, p0, d, q.y_ap, q.n_si, q.n_vac, ray);
        pack_nf_toy_rows(ray, p0, q.inv_2sig2, rows);
        st = ray.status;
    }
#pragma unroll
    for (int k = 0; k < ROWS; k++) q.out[(size_t)k * n + r] = rows[k];
    if (q.status) q.status[r] = st;
}


// Receiver fields of many receiver positions at once: path length and feed-taper amplitude of every ray (rows 3 and 4
// of rx_to_lyot_model.output(), rti:848-853) -- what the per-pixel integral needs from the receiver side.
// out_opl / out_amp [nrx][n_rays]; grid (ray blocks, receiver).
__global__ void __launch_bounds__(128) nf_rx_fields_kernel(NfRowsParams q, const double* p_rx, double* out_opl, double* out_amp, int* status)
{
    const int n = q.ang.rays();
    const int r = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (r >= n) return;
    const vec3 p0 = mk(__ldg(p_rx + 3 * k), __ldg(p_rx + 3 * k + 1), __ldg(p_rx + 3 * k + 2));
    NfRxRay ray;
    nf_rx_ray(K_COEF, p0, ray_dir(q.ang, r), q.n_si, q.n_vac, ray);
    const int j = r / q.ang.nth, i = r - j * q.ang.nth;
    double rows[20];
    pack_nf_rx_rows(ray, p0, q.theta[i] - q.th_mean, q.phi[j] - q.ph_mean, q.inv_2sig2, rows);
    out_opl[(size_t)k * n + r] = rows[3];
    out_amp[(size_t)k * n + r] = rows[4];
    if (ray.status) atomicOr(status, ray.status);
}

// Mirror pairs.  The LAT is symmetric under x -> -x (only even powers of x in both mirrors, source and rotation centre
// at x = 0) and so is the reference's launch fan (phi = linspace(0, pi)): ray (i, j) of pixel (az, el) and ray
// (i, nph-1-j) of pixel (-az, el) are mirror images -- same path length, focal x of opposite sign, same focal z.  A map
// with an az grid symmetric about 0 (ff_sim.py:121-124) therefore needs every ray traced once for TWO pixels: the
// work item is the canonical pixel (|az|, el); the traced ray is binned into the winner table of (+|az|, el) as ray
// j*nth+i at (x, z) and into that of (-|az|, el) as ray (nph-1-j)*nth+i at (-x, z).  A pixel with az < 0 is ALWAYS
// computed as the mirror of its canonical pixel, whether or not its partner is in the list, so a pixel's value never
// depends on how the list is sharded.  Pairing runs on the device (open-addressing table keyed by (|az|, el)).
struct FfItem {
    double az, el;       // canonical pixel: az >= 0
    int pos, neg;        // output index of (+az, el) and of (-az, el); -1 = not asked for
};
struct FfPairSlot { int owner, pos, neg; };
// Partners are matched with an absolute tolerance of 1e-17 rad on |az|: np.linspace(-a, a, n) is symmetric only to a
// few 1e-19 rad.  The item takes the az of its (+az) pixel; tracing the mirror pixel at an az that is off by 1e-17 rad
// moves a focal position by 1e-13 mm and a path length by 3e-14 mm.  Buckets of 2e-17 rad: a partner is looked up in
// its own bucket and both neighbours.
// (The near-field map pairs source positions (x, z), (-x, z) the same way, with 1e-11 mm.)
struct PairTol {
    double tol, scale, limit;    // match tolerance, buckets per unit (= 1 / (2 tol)), largest |first coordinate| that is bucketed
};
constexpr PairTol FF_PAIR_TOL = {1e-17, 5e16, 1e2};
constexpr PairTol NF_PAIR_TOL = {1e-11, 5e10, 1e8};
__device__ __forceinline__ long long ff_pair_bucket(double abs_az, const PairTol& t) { return (long long)(abs_az * t.scale); }
__device__ __forceinline__ unsigned ff_pair_hash(long long bucket, double e)
{
    unsigned long long x = (unsigned long long)bucket * 0x9E3779B97F4A7C15ull;
    x ^= (unsigned long long)__double_as_longlong(e) * 0xC2B2AE3D27D4EB4Full;
    x ^= x >> 29;
    return (unsigned)(x * 0xBF58476D1CE4E5B9ull >> 32);
}
// phase 1: every pixel with az >= 0 claims a slot.  ints[0] = number of items, ints[1] = length of the overflow list
// (POW2: the table has a power-of-two number of slots, h & (nslots - 1) instead of the integer division)
template <bool POW2 = false>
__device__ __forceinline__ void ff_pair_insert(int p, const double* azel, FfPairSlot* slots, unsigned nslots, int* over, int* ints,
                                               const PairTol& tol)
{
    const double az = azel[2 * p], el = azel[2 * p + 1];
    if (!(az == az) || !(el == el)) { over[atomicAdd(ints + 1, 1)] = p; return; }   // NaN pointing: an item of its own
    if (az < 0.0) return;
    if (!(az < tol.limit)) { over[atomicAdd(ints + 1, 1)] = p; return; }             // beyond the bucket range: an item of its own
    unsigned h = POW2 ? ff_pair_hash(ff_pair_bucket(az, tol), el) & (nslots - 1) : ff_pair_hash(ff_pair_bucket(az, tol), el) % nslots;
    for (;;) {
        const int o = atomicCAS(&slots[h].owner, -1, p);
        if (o == -1) { slots[h].pos = p; return; }
        if (fabs(azel[2 * o]) == az && azel[2 * o + 1] == el) {                          // the same pixel twice in the list
            over[atomicAdd(ints + 1, 1)] = p;
            return;
        }
        h = h + 1 == nslots ? 0 : h + 1;
    }
}
// phase 2: every pixel with az < 0 looks for the slot of its mirror pixel
template <bool POW2 = false>
__device__ __forceinline__ void ff_pair_attach(int p, const double* azel, FfPairSlot* slots, unsigned nslots, int* over, int* ints,
                                               const PairTol& tol)
{
    const double az = azel[2 * p], el = azel[2 * p + 1];
    if (!(az < 0.0) || !(el == el)) return;
    const double a = -az;
    for (int dk = 0; dk < 3 && a < tol.limit; dk++) {
        const long long bucket = ff_pair_bucket(a, tol) + (dk == 0 ? 0 : (dk == 1 ? -1 : 1));
        unsigned h = POW2 ? ff_pair_hash(bucket, el) & (nslots - 1) : ff_pair_hash(bucket, el) % nslots;
        for (;;) {
            const int o = slots[h].owner;
            if (o == -1) break;
            if (azel[2 * o + 1] == el && fabs(fabs(azel[2 * o]) - a) <= tol.tol) {
                if (atomicCAS(&slots[h].neg, -1, p) == -1) return;               // paired
                dk = 3;                                                           // partner taken (duplicate): an item of its own
                break;
            }
            h = h + 1 == nslots ? 0 : h + 1;
        }
    }
    over[atomicAdd(ints + 1, 1)] = p;
}
// phase 3: slots and overflow list -> items; the table is left empty for the next call and the other call's counters
// (ints_next) are zeroed
__device__ __forceinline__ void ff_pair_emit(unsigned h, const double* azel, FfPairSlot* slots, unsigned nslots, const int* over, int* ints,
                                             int* ints_next, FfItem* items)
{
    if (h < nslots && slots[h].owner >= 0) {
        FfItem it;
        it.az = fabs(azel[2 * slots[h].owner]); it.el = azel[2 * slots[h].owner + 1]; it.pos = slots[h].pos; it.neg = slots[h].neg;
        items[atomicAdd(ints, 1)] = it;
        slots[h].owner = slots[h].pos = slots[h].neg = -1;
    }
    if (h < (unsigned)ints[1]) {
        const int p = over[h];
        const double az = azel[2 * p];
        FfItem it;
        it.az = fabs(az); it.el = azel[2 * p + 1]; it.pos = az < 0.0 ? -1 : p; it.neg = az < 0.0 ? p : -1;
        items[atomicAdd(ints, 1)] = it;
    }
    if (ints_next && h < 2) ints_next[h] = 0;
}
__global__ void ff_pair_insert_kernel(const double* azel, int npix, FfPairSlot* slots, unsigned nslots, int* over, int* ints, PairTol tol)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npix) ff_pair_insert(p, azel, slots, nslots, over, ints, tol);
}
__global__ void ff_pair_attach_kernel(const double* azel, int npix, FfPairSlot* slots, unsigned nslots, int* over, int* ints, PairTol tol)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npix) ff_pair_attach(p, azel, slots, nslots, over, ints, tol);
}
__global__ void ff_pair_emit_kernel(const double* azel, FfPairSlot* slots, unsigned nslots, const int* over, int* ints, int* ints_next,
                                    FfItem* items)
{
    ff_pair_emit(blockIdx.x * blockDim.x + threadIdx.x, azel, slots, nslots, over, ints, ints_next, items);
}
// all three phases in one CTA with the table in shared memory (pixel lists up to FF_PAIR_ONE_CTA: one short launch
// instead of three); ints[0] = number of items
constexpr int FF_PAIR_ONE_CTA = 4096;
static unsigned ff_pair_slots(int npix) { unsigned n = 2; while (2 * n < 3u * (unsigned)npix) n *= 2; return n; }   // power of two, load factor <= 2/3
static size_t ff_pair_smem(int npix) { return (size_t)ff_pair_slots(npix) * sizeof(FfPairSlot) + ((size_t)npix + 2) * sizeof(int) + 8 + 2 * (size_t)npix * sizeof(double); }
// Second stage (near-field quads, items2 != nullptr): the items of the first stage are paired once more on their
// second coordinate -- item k -> (z, |x|) -- and the items that come out hold |z|, |x| and the indices of the
// first-stage items of (+z) and (-z); ints[2] = their number.  The pixel list is staged in shared memory (every phase
// reads it several times, and a lone CTA has nothing to hide global latency behind).
__global__ void __launch_bounds__(1024) ff_pair_kernel(const double* azel_g, int npix, unsigned nslots, int* ints, FfItem* items, PairTol tol,
                                                        FfItem* items2 = nullptr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FfPairSlot* slots = reinterpret_cast<FfPairSlot*>(smem_raw);
    int* cnt = reinterpret_cast<int*>(slots + nslots);      // [0] items, [1] overflow length
    int* over = cnt + 2;
    double* azel = reinterpret_cast<double*>((reinterpret_cast<size_t>(over + npix) + 7) / 8 * 8);   // [npix][2]
    for (int stage = 0; stage < (items2 ? 2 : 1); stage++) {
        for (unsigned h = threadIdx.x; h < nslots; h += blockDim.x) slots[h].owner = slots[h].pos = slots[h].neg = -1;
        if (stage) {
            npix = cnt[0];
            __syncthreads();                                // (cnt[0] is cleared below)
            for (int p = threadIdx.x; p < npix; p += blockDim.x) { azel[2 * p] = items[p].el; azel[2 * p + 1] = items[p].az; }   // this CTA's own writes
            items = items2;
        } else {
            for (int p = threadIdx.x; p < 2 * npix; p += blockDim.x) azel[p] = __ldg(azel_g + p);
        }
        __syncthreads();
        if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
        __syncthreads();
        for (int p = threadIdx.x; p < npix; p += blockDim.x) ff_pair_insert<true>(p, azel, slots, nslots, over, cnt, tol);
        __syncthreads();
        for (int p = threadIdx.x; p < npix; p += blockDim.x) ff_pair_attach<true>(p, azel, slots, nslots, over, cnt, tol);
        __syncthreads();
        for (unsigned h = threadIdx.x; h < nslots; h += blockDim.x) ff_pair_emit(h, azel, slots, nslots, over, cnt, nullptr, items);
        __syncthreads();
        if (threadIdx.x == 0) ints[2 * stage] = cnt[0];
    }
}


// ------------------------------------------------------------------------------------------------
// near field: fused map
// ------------------------------------------------------------------------------------------------
constexpr int NF_MAXF = 4;    // wavenumbers per launch
constexpr int NF_MAXC = 8;    // channels (wavenumber x receiver table) per launch
constexpr int NF_T = SOLAT_NF_T;
constexpr int NF_MAX_SUMS = 16;   // complex accumulators per thread (outputs x channels) a launch may carry
struct NfMapParams {
    AngleTabs ang;
    const double* src_xz;      // [npix][2]
    const int* act_idx;        // [n_act] ray indices with a non-zero receiver weight in any table
    const double2* phasor;     // ([4])[nrx][nfreq][n_act]  w_rx exp(i k L_rx / 1e3)  (holo_sim.py:173-176)
    const double2* first_leg;  // [n_act][4]  pixel-independent start of every active ray (nf_source_first_leg)
    // Mirror images (see nf_map_kernel).  sym bit 0: sources (x, z), (-x, z) are mirror images, bit 1: (x, z), (x, -z).
    // A symmetric plan (sym != 0) stores phasor[((m * nrx + table) * nfreq + f) * n_act + a], m = 0 the ray itself,
    // 1 / 2 / 3 its mirror ray in x / z / both.
    // MODE 1: items[k] = (|x|, z, pixel of +x, pixel of -x), counters[0] of them.
    // MODE 2: items2[k] = (|z|, |x|, index in `items` of the +z pair, of the -z pair), counters2[0] of them.
    const FfItem* items;
    const FfItem* items2;
    const int* counters;
    const int* counters2;
    int sym;
    double y_ap, n_si, n_vac;
    double kk[NF_MAXF];        // k/1e3 [rad/mm] of the wavenumbers of this launch
    int n_act, tiles, tile_len, nfreq_total, nch_total, f0, j0;
    double2* partial;          // [npix][tiles][nch_total], channel = table * nfreq + wavenumber
    int* status;
};

// CTA-wide sum of NC complex accumulators: warp shuffles, then one shared-memory hop.  dst[c * stride]
template <int NC>
__device__ __forceinline__ void block_reduce_store(double* re, double* im, double2* dst, int stride = 1)
{
    __shared__ double sh[2 * NF_MAXC][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int f = 0; f < NC; f++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            re[f] += __shfl_down_sync(0xffffffffu, re[f], o);
            im[f] += __shfl_down_sync(0xffffffffu, im[f], o);
        }
        if (lane == 0) { sh[2 * f][warp] = re[f]; sh[2 * f + 1][warp] = im[f]; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int f = 0; f < NC; f++) {
            double a = lane < nwarp ? sh[2 * f][lane] : 0.0, b = lane < nwarp ? sh[2 * f + 1][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_down_sync(0xffffffffu, a, o);
                b += __shfl_down_sync(0xffffffffu, b, o);
            }
            if (lane == 0) dst[f * stride] = make_double2(a, b);
        }
    }
}

// One CTA per (source position, ray tile); thread per ray.  The source ray is traced ONCE and feeds
// NF wavenumbers x NT receiver tables: exp(i k L_source) is formed per wavenumber and multiplied into the
// precomputed receiver phasor of every table (frequency sweeps of one band and receiver-offset sweeps share
// the geometry: BASELINE configs 3 and 4).
// Mirror images.  The optics tube is a body of revolution about y.  With a launch fan symmetric in phi about pi/2, ray
// (i, j) from the source (x, z) and ray (i, nph-1-j) from (-x, z) are mirror images with the same path length; with a
// fan symmetric in theta about pi/2 the same holds for (x, -z) and ray (nth-1-i, j).  A pixel on the negative side of a
// symmetric axis is ALWAYS computed as the mirror image of its canonical source (|x|, |z|) -- rays traced in the
// canonical order, receiver phasors looked up at the mirror ray -- so its value does not depend on what else is in
// the list or on how the list is sharded.
// MODE 0: one output pixel per CTA row.  MODE 1: the work item is (|x|, z) and every traced ray feeds the two pixels
// (+-x, z).  MODE 2: the work item is (|x|, |z|) and feeds the four pixels (+-x, +-z).  (Two / four sets of NF * NT
// accumulators: beyond about 16 complex sums they live in local memory, a few loads and stores per traced ray, far
// cheaper than tracing the ray again.)
template <int NF, int NT, int MODE>
__global__ void __launch_bounds__(NF_T, SOLAT_NF_MINB) nf_map_kernel(NfMapParams q)
{
    constexpr int NC = NF * NT;
    constexpr int NO = MODE == 2 ? 4 : (MODE == 1 ? 2 : 1);
    const Coef& sC = K_COEF;   // constant bank; a shared-memory copy measured 2-4 % slower (profiles/r01_variants.md)
    const int pix = blockIdx.x / q.tiles, tile = blockIdx.x - pix * q.tiles;
    int out[NO];
    int msel = 0;                                           // MODE 0: which mirror image this pixel is
    vec3 p0;
    if (MODE == 2) {
        if (pix >= __ldg(q.counters2)) return;             // one CTA row per pixel is launched, items are fewer
        const FfItem it = q.items2[pix];
        p0 = mk(it.el, q.y_ap, it.az);
#pragma unroll
        for (int o = 0; o < NO; o++) out[o] = -1;
        if (it.pos >= 0) { out[0] = q.items[it.pos].pos; out[1] = q.items[it.pos].neg; }
        if (it.neg >= 0) { out[2 % NO] = q.items[it.neg].pos; out[3 % NO] = q.items[it.neg].neg; }
    } else if (MODE == 1) {
        if (pix >= __ldg(q.counters)) return;
        const FfItem it = q.items[pix];
        // the z mirror of a pair item is taken as a whole: (|x|, -z) -> trace (|x|, |z|), both pixels through the z map
        msel = ((q.sym & 2) && it.el < 0.0) ? 2 : 0;
        p0 = mk(it.az, q.y_ap, msel ? -it.el : it.el);
        out[0] = it.pos; out[1 % NO] = it.neg;
    } else {
        const double x = __ldg(q.src_xz + 2 * pix), z = __ldg(q.src_xz + 2 * pix + 1);
        msel = (((q.sym & 1) && x < 0.0) ? 1 : 0) + (((q.sym & 2) && z < 0.0) ? 2 : 0);
        p0 = mk((msel & 1) ? -x : x, q.y_ap, (msel & 2) ? -z : z);
        out[0] = pix;
    }
    // which mirror ray's receiver phasor every output takes: MODE 0 {msel}, MODE 1 {msel, msel | 1}, MODE 2 {0, 1, 2, 3};
    // a symmetric plan holds one copy of the tables per mirror image, gathered through the mirror map once
    size_t mo[NO];
#pragma unroll
    for (int o = 0; o < NO; o++) mo[o] = (size_t)(MODE == 2 ? o : (msel | o)) * q.nch_total * q.n_act;
    double re[NO * NC], im[NO * NC];
#pragma unroll
    for (int c = 0; c < NO * NC; c++) re[c] = im[c] = 0.0;
    int st = 0;
    const int q_end = min((tile + 1) * q.tile_len, q.n_act);
    for (int a = tile * q.tile_len + threadIdx.x; a < q_end; a += NF_T) {
        // the leg to the flat front of lens 1 does not depend on the source position: two 32-byte loads instead of a
        // reciprocal, a square root and the launch direction
        const double2 l0 = __ldg(q.first_leg + 4 * a), l1 = __ldg(q.first_leg + 4 * a + 1), l2 = __ldg(q.first_leg + 4 * a + 2),
                      l3 = __ldg(q.first_leg + 4 * a + 3);          // {t1, d0.x}, {hit y, d0.z}, {d.x, d.y}, {d.z, status}
        NfSourceRay ray;
        ray.p1b = mk(fma(l0.y, l0.x, p0.x), l1.x, fma(l1.y, l0.x, p0.z));
        nf_source_ray_from_1b(sC, mk(l2.x, l2.y, l3.x), l0.x, (int)l3.y, q.n_si, q.n_vac, ray);
        st |= ray.status;
        if (ray.valid) {
#pragma unroll
            for (int f = 0; f < NF; f++) {
                double sn, cs;
                sincos_mid(ray.opl * q.kk[f], sn, cs);                            // holo_sim.py:177
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const double2* tab = q.phasor + ((size_t)(q.j0 + t) * q.nfreq_total + (q.f0 + f)) * q.n_act + a;
#pragma unroll
                    for (int o = 0; o < NO; o++) {
                        if (MODE != 0 && out[o] < 0) continue;                      // CTA-uniform
                        const double2 ph = __ldg(tab + mo[o]);
                        const int c = o * NC + t * NF + f;
                        re[c] = fma(ph.x, cs, fma(-ph.y, sn, re[c]));
                        im[c] = fma(ph.x, sn, fma(ph.y, cs, im[c]));
                    }
                }
            }
        }
    }
    if (st) atomicOr(q.status, st);
    // accumulators are ordered [output][t][f]; channel (j0 + t) * nfreq + f0 + f
    __shared__ double2 tmp[4 * NF_MAXC];
#pragma unroll
    for (int o = 0; o < NO; o++) {
        block_reduce_store<NC>(re + o * NC, im + o * NC, tmp + o * NC);
        __syncthreads();                                    // tmp is complete, the reduction scratch free again
    }
    if (threadIdx.x < NO * NC) {
        const int o = threadIdx.x / NC, c = threadIdx.x - o * NC, t = c / NF, f = c - t * NF;
        int dst = out[0];
#pragma unroll
        for (int k = 1; k < NO; k++) if (o == k) dst = out[k];
        if (dst >= 0)
            q.partial[((size_t)dst * q.tiles + tile) * q.nch_total + (size_t)(q.j0 + t) * q.nfreq_total + q.f0 + f] = tmp[threadIdx.x];
    }
}

// pixel-independent first leg of every active source ray of a near-field plan
__global__ void nf_first_leg_kernel(AngleTabs ang, const int* act_idx, int n_act, double y_ap, double n_si, double n_vac, double2* tab)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_act) return;
    double a4[4], b4[4];
    nf_source_first_leg(K_COEF, ray_dir(ang, act_idx[a]), y_ap, n_si, n_vac, a4, b4);
    tab[4 * a] = make_double2(a4[0], a4[1]); tab[4 * a + 1] = make_double2(a4[2], a4[3]);
    tab[4 * a + 2] = make_double2(b4[0], b4[1]); tab[4 * a + 3] = make_double2(b4[2], b4[3]);
}

// receiver phasor tables of a near-field plan: ph[(j * nfreq + f) * n_act + a] = w exp(i k_f L_rx / 1e3)
__global__ void nf_phasor_kernel(const double2* opl_w, const double* k_per_m, int nrx, int nfreq, int n_act, double2* ph)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nrx * nfreq * n_act) return;
    const int a = (int)(idx % n_act), f = (int)((idx / n_act) % nfreq), j = (int)(idx / ((long long)n_act * nfreq));
    const double2 v = opl_w[(size_t)j * n_act + a];
    double sn, cs;
    sincos(v.x * k_per_m[f] / 1e3, &sn, &cs);
    ph[idx] = make_double2(v.y * cs, v.y * sn);
}

// symmetric plans: one more copy of the phasor tables per mirror image, out[m][channel][a] = ph[channel][mirror_m(a)]
// amap[(m - 1) * n_act + a], m = 1 (x), 2 (z), 3 (both) = active index of the mirror ray of active ray a
__global__ void nf_phasor_mirror_kernel(const double2* ph, const int* amap, int n_act, long long per_image, double2* out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (m * channels + channel) * n_act + a
    if (idx >= 4 * per_image) return;
    const int m = (int)(idx / per_image), a = (int)(idx % n_act);
    const long long ch = (idx - m * per_image) / n_act;
    out[idx] = ph[ch * n_act + (m ? amap[(size_t)(m - 1) * n_act + a] : a)];
}

// ------------------------------------------------------------------------------------------------
// near field, sweeps over many receiver tables: trace once, contract against every table
// ------------------------------------------------------------------------------------------------
// With more than NF_MAX_SUMS channels the fused kernel has to re-trace (and re-read the phasor tables, 16 bytes per
// output pixel, ray and table: the 25-table map of BASELINE config 4 moved 7.6 GB out of L2).  Here the quad items
// are traced ONCE per wavenumber into E[item][a] = exp(i k L) (0 for a ray that does not arrive), and the maps are the
// complex matrix product  out[item][m * nrx + j] = sum_a E[item][a] * phasor[m][j][f][a]  -- m the mirror image, j the
// receiver table -- computed by a register-tiled FP64 kernel (64 x 64 outputs per CTA, 4 x 4 per thread; every table
// entry read from L2 once per 64 items).  The ray range is cut into `splits` pieces that depend on the ray count only,
// nf_finalize_kernel adds them up in order.
__global__ void __launch_bounds__(NF_T, SOLAT_NF_MINB) nf_phase_kernel(NfMapParams q, int row0, int rows, double2* E, int ldE)
{
    const Coef& sC = K_COEF;
    const int r = blockIdx.x / q.tiles, tile = blockIdx.x - r * q.tiles;
    const int item = row0 + r;
    if (r >= rows || item >= __ldg(q.counters2)) return;
    const FfItem it = q.items2[item];
    const vec3 p0 = mk(it.el, q.y_ap, it.az);
    int st = 0;
    const int q_end = min((tile + 1) * q.tile_len, q.n_act);
    for (int a = tile * q.tile_len + threadIdx.x; a < q_end; a += NF_T) {
        const double2 l0 = __ldg(q.first_leg + 4 * a), l1 = __ldg(q.first_leg + 4 * a + 1), l2 = __ldg(q.first_leg + 4 * a + 2),
                      l3 = __ldg(q.first_leg + 4 * a + 3);
        NfSourceRay ray;
        ray.p1b = mk(fma(l0.y, l0.x, p0.x), l1.x, fma(l1.y, l0.x, p0.z));
        nf_source_ray_from_1b(sC, mk(l2.x, l2.y, l3.x), l0.x, (int)l3.y, q.n_si, q.n_vac, ray);
        st |= ray.status;
        double sn = 0.0, cs = 0.0;
        if (ray.valid) sincos_mid(ray.opl * q.kk[0], sn, cs);                     // holo_sim.py:177
        E[(size_t)r * ldE + a] = make_double2(cs, sn);
    }
    if (st) atomicOr(q.status, st);
}

constexpr int NF_GM = 64, NF_GN = 64, NF_GK = 8;
struct NfContractParams {
    const double2* E;          // [rows][ldE]
    const double2* phasor;     // [4][nrx][nfreq][n_act]
    const FfItem* items;       // pair items
    const FfItem* items2;      // quad items
    const int* counters2;
    int row0, rows, ldE, n_act, nrx, nfreq, f, nch_total, splits, k_len, pix0;
    double2* partial;          // [npix][splits][nch_total]
};
__global__ void __launch_bounds__(256, 2) nf_contract_kernel(NfContractParams q)
{
    __shared__ double2 As[NF_GK][NF_GM + 1], Bs[NF_GK][NF_GN + 1];
    const int count = min(__ldg(q.counters2) - q.row0, q.rows);       // rows of E that hold an item
    const int r0 = blockIdx.x * NF_GM;
    if (r0 >= count) return;
    // launch order: all splits of the first (full) column tile, then those of the next ones -- the last tile is
    // usually the narrowest (dead column groups are skipped), and the cheap CTAs fill the tail of the last wave
    const int lin = blockIdx.y + gridDim.y * blockIdx.z, ytile = lin / gridDim.z;
    const int N = 4 * q.nrx, c0 = ytile * NF_GN, split = lin - ytile * gridDim.z;
    const int k_begin = split * q.k_len, k_end = min(k_begin + q.k_len, q.n_act);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // staging: thread -> rows sr and sr + 32 of both tiles, ray k0 + sk
    const int sr = threadIdx.x >> 3, sk = threadIdx.x & 7;
    const bool a_on0 = r0 + sr < count, a_on1 = r0 + sr + 32 < count, b_on0 = c0 + sr < N, b_on1 = c0 + sr + 32 < N;
    const double2* pa0 = q.E + (size_t)(r0 + sr) * q.ldE + sk;
    const double2* pa1 = pa0 + (size_t)32 * q.ldE;
    const double2* pb0 = q.phasor + ((size_t)(c0 + sr) * q.nfreq + q.f) * q.n_act + sk;     // row m * nrx + j
    const double2* pb1 = pb0 + (size_t)32 * q.nfreq * q.n_act;
    const double2 zero = make_double2(0.0, 0.0);
    double re[4][4], im[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) re[i][j] = im[i][j] = 0.0;
    // the columns of this tile come in four groups of 16 (one per j): groups past the last column are skipped
    const int groups = min(4, (N - c0 + 15) >> 4);
    double2 va0, va1, vb0, vb1;
    {
        const bool k_on = k_begin + sk < k_end;
        va0 = k_on && a_on0 ? pa0[k_begin] : zero; va1 = k_on && a_on1 ? pa1[k_begin] : zero;
        vb0 = k_on && b_on0 ? __ldg(pb0 + k_begin) : zero; vb1 = k_on && b_on1 ? __ldg(pb1 + k_begin) : zero;
    }
    for (int k0 = k_begin; k0 < k_end; k0 += NF_GK) {
        As[sk][sr] = va0; As[sk][sr + 32] = va1; Bs[sk][sr] = vb0; Bs[sk][sr + 32] = vb1;
        __syncthreads();
        {   // the next stretch of rays is on its way while this one is multiplied
            const int kn = k0 + NF_GK;
            const bool k_on = kn + sk < k_end;
            va0 = k_on && a_on0 ? pa0[kn] : zero; va1 = k_on && a_on1 ? pa1[kn] : zero;
            vb0 = k_on && b_on0 ? __ldg(pb0 + kn) : zero; vb1 = k_on && b_on1 ? __ldg(pb1 + kn) : zero;
        }
#pragma unroll
        for (int kk = 0; kk < NF_GK; kk++) {
            double2 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty + 16 * i]; b[i] = Bs[kk][tx + 16 * i]; }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j >= groups) break;                                            // CTA-uniform
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    re[i][j] = fma(b[j].x, a[i].x, fma(-b[j].y, a[i].y, re[i][j]));
                    im[i][j] = fma(b[j].x, a[i].y, fma(b[j].y, a[i].x, im[i][j]));
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = r0 + ty + 16 * i;
        if (r >= count) continue;
        const FfItem it = q.items2[q.row0 + r];
        int o0 = -1, o1 = -1, o2 = -1, o3 = -1;
        if (it.pos >= 0) { o0 = q.items[it.pos].pos; o1 = q.items[it.pos].neg; }
        if (it.neg >= 0) { o2 = q.items[it.neg].pos; o3 = q.items[it.neg].neg; }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c0 + tx + 16 * j;
            if (c >= N) continue;
            const int m = c / q.nrx, t = c - m * q.nrx;
            const int pixel = m == 0 ? o0 : (m == 1 ? o1 : (m == 2 ? o2 : o3));
            if (pixel >= 0)
                q.partial[((size_t)(q.pix0 + pixel) * q.splits + split) * q.nch_total + (size_t)t * q.nfreq + q.f] = make_double2(re[i][j], im[i][j]);
        }
    }
}

// The same contraction on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64: one warp instruction = 8 x 8 x 4 FMAs and holds
// the FP64 pipe for 16 cycles -- measured 37.0 TFLOP/s on B200, the DFMA issue peak -- so one issue slot feeds the pipe
// for as long as 8 DFMA instructions would, and the loads ride in its shadow).  The complex product is a real one of
// twice the depth: with E and the phasors viewed as arrays of doubles (re, im interleaved along the ray index),
//   Re out[i][c] = sum_k' E'[i][k'] * Bre[c][k'],  Bre[c][2k] = P.re, Bre[c][2k+1] = -P.im
//   Im out[i][c] = sum_k' E'[i][k'] * Bim[c][k'],  Bim[c][2k] = P.im, Bim[c][2k+1] =  P.re
// and in the m8n8k4 fragment layout the 8 columns of a tile are (re, im) of 4 complex columns: lane l holds
// A'[row l/4][k' l%4], B[column (l/4)/2, part (l/4)%2][k' l%4] and the accumulators (re, im) of (row l/4, column l%4): a
// B operand is one double of the phasor element, offset part ^ (l & 1), negated for (part 0, odd k').
// CTA: 64 items x up to 112 complex columns x one ray split; 8 warps as 2 (rows) x 4 (columns), 4 x up to 7 MMA tiles per
// warp; 16 rays (32 doubles) per stage, two stages filled with cp.async (zero fill past the edges).
constexpr int DM_ROWS = 64, DM_COLS = 112, DM_KS = 16, DM_LD = 2 * DM_KS + 4;   // row pitch in doubles: conflict-free fragment loads
constexpr size_t DM_STAGE = (size_t)(DM_ROWS + DM_COLS) * DM_LD;                 // doubles per stage
static size_t nf_dmma_smem() { return 2 * DM_STAGE * sizeof(double); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool on)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = on ? 16 : 0;                         // src-size 0: the 16 bytes are zero filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes));
}
#ifndef SOLAT_DMMA_MINB
#define SOLAT_DMMA_MINB 2
#endif
__global__ void __launch_bounds__(256, SOLAT_DMMA_MINB) nf_contract_dmma_kernel(NfContractParams q)
{
    extern __shared__ __align__(16) double dm_smem[];
    const int count = min(__ldg(q.counters2) - q.row0, q.rows);       // rows of E that hold an item
    const int r0 = blockIdx.x * DM_ROWS;
    if (r0 >= count) return;
    const int N = 4 * q.nrx, c0 = blockIdx.y * DM_COLS, ncols = min(DM_COLS, N - c0), split = blockIdx.z;
    const int k_begin = split * q.k_len, k_end = min(k_begin + q.k_len, q.n_act);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3;
    const int tiles = (ncols + 3) >> 2;                                // 8-wide MMA tiles = 4 complex columns each
    const int t_lo = (wn * tiles) >> 2, t_hi = ((wn + 1) * tiles) >> 2; // this warp's tiles (<= 7)
    const int nsteps = (k_end - k_begin + DM_KS - 1) / DM_KS;

    auto fill = [&](int stage, int step) {
        double* As = dm_smem + stage * DM_STAGE;
        double* Bs = As + DM_ROWS * DM_LD;
        const int k0 = k_begin + step * DM_KS;
        for (int e = tid; e < DM_ROWS * DM_KS; e += 256) {             // E rows: 16 double2 per row
            const int r = e / DM_KS, k = e - r * DM_KS;
            const bool on = r0 + r < count && k0 + k < k_end;
            cp_async16(As + r * DM_LD + 2 * k, q.E + (size_t)(on ? r0 + r : 0) * q.ldE + (on ? k0 + k : 0), on);
        }
        for (int e = tid; e < DM_COLS * DM_KS; e += 256) {             // phasor rows m * nrx + j
            const int c = e / DM_KS, k = e - c * DM_KS;
            const bool on = c < ncols && k0 + k < k_end;
            cp_async16(Bs + c * DM_LD + 2 * k, q.phasor + ((size_t)(on ? c0 + c : 0) * q.nfreq + q.f) * q.n_act + (on ? k0 + k : 0), on);
        }
        asm volatile("cp.async.commit_group;");
    };

    double acc[4][7][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 7; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int part = (lane >> 2) & 1, kq = lane & 3;
    const int b_off = (kq & ~1) | ((kq & 1) ^ part);                   // which double of the phasor pair
    const double b_sign = (part == 0 && (kq & 1)) ? -1.0 : 1.0;
    const int a_base = (wm * 32 + (lane >> 2)) * DM_LD + kq;           // + i * 8 * DM_LD + 4 * kk
    const int b_base = (lane >> 3) * DM_LD + b_off;                    // + tile * 4 * DM_LD + 4 * kk

    if (nsteps > 0) fill(0, 0);
    for (int step = 0; step < nsteps; step++) {
        if (step + 1 < nsteps) {
            fill((step + 1) & 1, step + 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const double* As = dm_smem + (step & 1) * DM_STAGE;
        const double* Bs = As + DM_ROWS * DM_LD;
#pragma unroll
        for (int kk = 0; kk < DM_KS / 2; kk++) {                       // 4 doubles = 2 rays per MMA
            double a[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[a_base + i * 8 * DM_LD + 4 * kk];
#pragma unroll
            for (int j = 0; j < 7; j++) {
                if (t_lo + j >= t_hi) break;                           // warp-uniform
                const double b = b_sign * Bs[b_base + (t_lo + j) * 4 * DM_LD + 4 * kk];
#pragma unroll
                for (int i = 0; i < 4; i++)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b));
            }
        }
        __syncthreads();                                               // the stage is refilled two steps on
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = r0 + wm * 32 + i * 8 + (lane >> 2);
        if (r >= count) continue;
        const FfItem it = q.items2[q.row0 + r];
        int o0 = -1, o1 = -1, o2 = -1, o3 = -1;
        if (it.pos >= 0) { o0 = q.items[it.pos].pos; o1 = q.items[it.pos].neg; }
        if (it.neg >= 0) { o2 = q.items[it.neg].pos; o3 = q.items[it.neg].neg; }
#pragma unroll
        for (int j = 0; j < 7; j++) {
            if (t_lo + j >= t_hi) break;
            const int c = c0 + (t_lo + j) * 4 + (lane & 3);
            if (c >= N) continue;
            const int m = c / q.nrx, t = c - m * q.nrx;
            const int pixel = m == 0 ? o0 : (m == 1 ? o1 : (m == 2 ? o2 : o3));
            if (pixel >= 0)
                q.partial[((size_t)(q.pix0 + pixel) * q.splits + split) * q.nch_total + (size_t)t * q.nfreq + q.f] =
                    make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

// Fixed-order sum of the tile partials -> amp, phase (holo_sim.py:181-182).  (Folding this into nf_map_kernel -- the CTA
// that stores an item's last tile adds them up -- measured the same time: profiles/r01_variants.md.)
__global__ void nf_finalize_kernel(const double2* partial, int npix, int tiles, int nfreq, double* amp, double* phase)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // channel * npix + pix
    if (idx >= npix * nfreq) return;
    const int f = idx / npix, pix = idx - f * npix;
    double re = 0.0, im = 0.0;
    for (int t = 0; t < tiles; t++) {
        const double2 v = partial[((size_t)pix * tiles + t) * nfreq + f];
        re += v.x; im += v.y;
    }
    amp[idx] = hypot(re, im);
    phase[idx] = atan2(im, re);
}

// ------------------------------------------------------------------------------------------------
// far field: per-ray dumps
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ FfGeom to_geom(const solat_ff_geom& g)
{
    FfGeom r;
    r.th_1 = g.th_1; r.th2 = g.th2;
    for (int i = 0; i < 3; i++) { r.rotc[i] = g.rotc_mm[i]; r.src[i] = g.src_mm[i]; }
    r.foc_y = g.foc_y_mm;
    return r;
}

struct FfRowsParams {
    AngleTabs ang;
    solat_ff_geom g;
    double el, az;
    double centre[3];
    const int *pidx1, *pidx2;
    const double *pan1, *pan2;
    int np1, np2;
    double* out;
    int* status;
};

template <bool STAGE1>
__global__ void __launch_bounds__(128) ff_rows_kernel(FfRowsParams q)
{
    __shared__ FfFrames fr;
    __shared__ double t2e[9];
    const FfGeom g = to_geom(q.g);
    if (threadIdx.x == 0) {
        ff_build_frames(K_COEF, g, q.el, q.az, fr);
        for (int i = 0; i < 9; i++) t2e[i] = fr.t2.a[i];
        if (g.src[2] != 0) for (int i = 0; i < 3; i++) t2e[3 * i + 2] /= cos(atan(1.0 / 3.0));   // saf:828-829
    }
    __syncthreads();
    const int n = q.ang.rays();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    constexpr int ROWS = STAGE1 ? 9 : 17;
    double rows[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; k++) rows[k] = 0.0;
    int st = 0;
    bool go = true;
    if (!STAGE1) go = (q.pidx1[r] >= 0) && (q.pidx2[r] >= 0);   // saf:626-628, :775-779
    if (go) {
        const vec3 d = ray_dir(q.ang, r);
        vec3 Pa;
        double t_ap;
        if (ff_aperture(K_COEF, g, fr, d, t_ap, Pa)) {
            FfRay ray;
            if (STAGE1) {
                ff_ray_trace<true, false>(K_COEF, fr, d, t_ap, Pa, g.foc_y, nullptr, nullptr, nullptr, ray);
                rows[0] = ray.m2.x; rows[1] = ray.m2.y; rows[2] = ray.m2.z;
                rows[3] = ray.m1.x; rows[4] = ray.m1.y; rows[5] = ray.m1.z;
            } else {
                PanelErr e1, e2;
                const double* c = q.pan1 + q.pidx1[r];
                e1.a = c[2 * q.np1]; e1.b = c[3 * q.np1]; e1.c = c[4 * q.np1]; e1.d = c[5 * q.np1]; e1.e = c[6 * q.np1];
                e1.x0 = c[8 * q.np1]; e1.y0 = c[9 * q.np1];
                c = q.pan2 + q.pidx2[r];
                e2.a = c[2 * q.np2]; e2.b = c[3 * q.np2]; e2.c = c[4 * q.np2]; e2.d = c[5 * q.np2]; e2.e = c[6 * q.np2];
                e2.x0 = c[8 * q.np2]; e2.y0 = c[9 * q.np2];
                ff_ray_trace<false, true>(K_COEF, fr, d, t_ap, Pa, g.foc_y, &e1, &e2, t2e, ray);
                pack_ff_rows(ray, q.centre, rows);
            }
            st = ray.status;
        }
    }
    if (STAGE1 && r == 0) { rows[6] = fr.centre[0]; rows[7] = fr.centre[1]; rows[8] = fr.centre[2]; }   // saf:563-565
#pragma unroll
    for (int k = 0; k < ROWS; k++) q.out[(size_t)k * n + r] = rows[k];
    if (q.status) q.status[r] = st;
}

// ------------------------------------------------------------------------------------------------
// far field: project_to_data (far_field_latrt.py:9-32)
// ------------------------------------------------------------------------------------------------
__global__ void ff_project_scatter_kernel(const double* sim, int n_pts, FfGrid G, int* win)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pts) return;
    if (sim[(size_t)15 * n_pts + r] == 0.0) return;                               // ffl:10
    const int c = ff_cell(sim[(size_t)9 * n_pts + r], sim[(size_t)11 * n_pts + r], G);
    if (c >= 0) atomicMax(win + c, r);                                            // ffl:28: later rays overwrite
}
__global__ void ff_project_fill_kernel(const double* sim, int n_pts, const int* win, int cells, double k, double2* out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    double2 v = make_double2(0.0, 0.0);
    const int r = win[c];
    if (r >= 0) cis_large(k * sim[(size_t)15 * n_pts + r] / 1e3, v.x, v.y);       // ffl:28-30
    out[c] = v;
}

// ------------------------------------------------------------------------------------------------
// far field: fused map
// ------------------------------------------------------------------------------------------------
// CTA shapes of the fused far-field map.  The winner table (8 B per E_rec cell) lives in shared memory, so a large grid
// limits the CTAs per SM; the shape is chosen per plan so that 24 warps stay resident (80 registers per thread):
// 256 threads x 3 CTAs, 384 x 2 (grids up to ~118 x 118), 768 x 1 (up to 168 x 168).
constexpr int FF_T = SOLAT_FF_T;
constexpr int FF_WQ = 64;            // per-warp queue capacity
constexpr int FF_MAX_ROWS = 1024;    // phi rows the interval walk keeps on chip (FF_ROW_K intervals each); larger fans take the queue walk
// Rectangular panel layout of one mirror on the device: panel k covers [x0, x1) x [y0, y1) in the coordinates the
// reference hands to pan_mod.identify_panel (saf:596-612: M2 hit (x, z), M1 hit (x, y - y_cent), telescope frame, ideal
// mirrors); a later panel overrides an earlier one (an identify_panel that loops over the panels and writes their ids);
// col[k] = column of the panel model (first match of the panel's (row, col) ids, saf:626,805) or -1.  A G x G bucket
// grid over the bounding box lists, per bucket, the panels that overlap it (ascending): the exact FP64 test runs on
// those only.
constexpr int PANEL_G = 32, PANEL_BK = 8;
struct PanelLayoutDev {
    const double4* rect;             // [n] (x0, x1, y0, y1)
    const int* col;                  // [n]
    const unsigned short* bucket;    // [G * G][BK] panel indices, 0xffff = end of list
    int n, linear;                   // linear: a bucket overflowed, scan all panels
    float bx0, by0, inv_bw, inv_bh;
};
__device__ __forceinline__ bool panel_holds(const PanelLayoutDev& L, int k, double x, double y)
{
    const double2* r = reinterpret_cast<const double2*>(L.rect + k);
    const double2 rx = __ldg(r), ry = __ldg(r + 1);
    return x >= rx.x && x < rx.y && y >= ry.x && y < ry.y;
}
__device__ __forceinline__ int panel_lookup(const PanelLayoutDev& L, double x, double y)
{
    int found = -1;
    if (L.linear) {
        for (int k = 0; k < L.n; k++)
            if (panel_holds(L, k, x, y)) found = k;
    } else {
        const int gx = (int)floorf(((float)x - L.bx0) * L.inv_bw), gy = (int)floorf(((float)y - L.by0) * L.inv_bh);
        if (gx < 0 || gx >= PANEL_G || gy < 0 || gy >= PANEL_G) return -1;
        const unsigned short* b = L.bucket + (size_t)(gy * PANEL_G + gx) * PANEL_BK;
        for (int i = 0; i < PANEL_BK; i++) {
            const int k = __ldg(b + i);
            if (k == 0xffff) break;
            if (panel_holds(L, k, x, y)) found = k;
        }
    }
    return found >= 0 ? __ldg(L.col + found) : -1;
}

struct FfMapParams {
    AngleTabs ang;
    solat_ff_geom g;
    const double* azel;      // [npix][2] (az, el)
    FfGrid grid;
    const double2* e_rec;    // (M, M)
    double k;
    double key_base, key_unit, key_inv_unit;   // fixed-point frame of the winner keys
    int key_shift;           // bits of the path-length field below the ray index (44, or 40 for fans above 2^20 - 1 rays)
    double* amp;
    double* phase;
    int* status;
    int rows;                // host: theta is strictly monotone inside one branch of tan -> interval walk allowed
    const unsigned char* rec;   // per-pixel records of ff_prep_kernel
    size_t rec_stride;
    const FfItem* items;     // pair mode: work items (azel unused), counters[0] of them; two winner tables per CTA
    const int* counters;
    int item0;               // pair mode: CTA b handles item item0 + b (its record is record b)
    // panel models.  PANEL = 1: per (pixel, ray) column of pan1 / pan2 derived on the host, -1 = the ray sits on no
    // panel.  PANEL = 2: the panels are looked up on the device from the ideal-mirror hits (lay1, lay2).
    const short *pidx1, *pidx2;
    const double *pan1, *pan2;   // (10, np) row-major [row, col, a, b, c, d, e, f, x0, y0]  (saf:667-674)
    int np1, np2;
    PanelLayoutDev lay1, lay2;
    unsigned long long* win_g;   // winner tables in global memory, [CTAs of the launch][tables][M * M]; nullptr: shared memory
    // large fans (ff_map_interp_kernel): cells from interpolated focal positions, exact traces for the winners only
    int interp_s, interp_ls;     // lattice spacing in rays (both fan axes), a power of two: S = 1 << ls
    double th0, dth, ph0, dph;   // the fan's axes are uniform: theta_i = th0 + i dth, phi_j = ph0 + j dph (lattice nodes beyond the fan)
    double interp_eps, interp_chk;   // [cm] a ray closer than eps to a cell edge is traced exactly; winners must match their prediction to chk
    int* redo;                   // [CTAs of the launch] interp kernel: 1 = prediction check failed, the exact kernel redoes the item;
                                 // exact kernel: when set, only the flagged CTAs run
    float i_tol_x, i_tol_y, i_uhi, i_umx, i_chk_x, i_chk_y;   // the classifier's constants in cell units (ff_map_run): they reach the
                                 // single-precision instructions as constant-bank operands, no register, no load
    float* probe_err;            // [CTAs of the launch] or nullptr (SOLAT_FF_INTERP_DEBUG): largest |exact - predicted| of an item's winners [cells]
};

__device__ __forceinline__ PanelErr load_panel(const double* pan, int np, int col)
{
    const double* c = pan + col;
    PanelErr e;
    e.a = __ldg(c + 2 * np); e.b = __ldg(c + 3 * np); e.c = __ldg(c + 4 * np); e.d = __ldg(c + 5 * np); e.e = __ldg(c + 6 * np);
    e.x0 = __ldg(c + 8 * np); e.y0 = __ldg(c + 9 * np);
    return e;
}

// In place: row_pre[j + 1] = count of row j  ->  row_pre[j] = rays before row j (row_pre[nph] = total).  One warp.
__device__ __forceinline__ void ff_prefix_rows(int* row_pre, int nph, int lane)
{
    int carry = 0;
    for (int base = 0; base < nph; base += 32) {
        const int j = base + lane;
        int v = j < nph ? row_pre[j + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (j < nph) row_pre[j + 1] = carry + v;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) row_pre[0] = 0;
}

// Per-pixel set-up of the fused map, one small CTA per pixel, all pixels of a launch in flight at once (the set-up is
// latency bound: one thread builds the frames, one thread per phi row finds its intervals): frames, and for the
// interval walk the per-row ray intervals (aperture, then ff_pred_row_keep), as a record in global
// memory the map kernel copies into its shared memory.  Record: FfFrames | int queue_walk, n_list, pad[2] |
// row_lo[K nph] | row_pre[K nph + 1]  (K = FF_ROW_K intervals per phi row: virtual row K j + k).
constexpr int FF_PREP_T = 64;     // = FF_PRED_SAMPLES; 6 CTAs per SM (164 registers, no spills) keep every pixel of a 41 x 41 map in flight
constexpr int FF_PREP_T_BIG = 256; // fans of at least FF_PILE_MIN rays
constexpr size_t FF_REC_FRAMES = ((sizeof(FfFrames) + 15) / 16) * 16;
constexpr size_t FF_REC_HEAD = FF_REC_FRAMES + 16;
struct FfPrepParams {
    AngleTabs ang;
    solat_ff_geom g;
    const double* azel;      // direct mode: pixel list
    const FfItem* items;     // pair mode: work items (azel unused); counters[0] = number of items
    const int* counters;
    int item0;               // pair mode: CTA b handles item item0 + b (its record is record b)
    int rows, prefilter;
    int pile;                // scouts + pile shrinking for items that list at least FF_PILE_MIN rays
    FfGrid grid;             // edges of the E_rec grid (ff_pred_row_keep)
    unsigned char* rec;
    size_t stride;
    int* clear_status;       // host-buffer entries: the call's own status word (8-byte slot behind the results), zeroed here
};
static size_t ff_rec_stride(int rows_nph) { return FF_REC_HEAD + (rows_nph ? ((2 * (size_t)FF_ROW_K * rows_nph + 1) * sizeof(int) + 15) / 16 * 16 : 0); }

// T = 64 (FF_PREP_T): the small items of reference-sized maps, all in flight at once; T = 256 (FF_PREP_T_BIG): items of
// large fans, where the scout stage scans hundreds of thousands of listed rays per item.
template <int T>
__global__ void __launch_bounds__(T, T == 64 ? 6 : 3) ff_prep_kernel(FfPrepParams q)
{
    constexpr int FF_T = T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FfFrames* fr = reinterpret_cast<FfFrames*>(smem_raw);
    const int nvr = FF_ROW_K * q.ang.nph;                                  // virtual rows: FF_ROW_K intervals per phi row
    int* row_lo = reinterpret_cast<int*>(smem_raw + FF_REC_FRAMES);        // [nvr] first ray of an interval
    int* row_pre = row_lo + (q.rows ? nvr : 0);                            // [nvr + 1] rays before the interval
    __shared__ int s_queue_walk;
    __shared__ FfPred s_pred;
    __shared__ double s_samp[FF_PRED_SAMPLES * FF_PRED_LD];   // sample rays: basis values, X, Z
    __shared__ double s_A[FF_PRED_NB * FF_PRED_LD], s_diag[FF_PRED_NB];
    __shared__ float s_pres[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pix = blockIdx.x;
    if (q.items && q.item0 + pix >= __ldg(q.counters)) return;   // pair mode launches one CTA per pixel, items are fewer
    const FfItem* item = q.items ? q.items + q.item0 + pix : nullptr;
    const FfGeom g = to_geom(q.g);
    const Coef& sC = K_COEF;
    const double az = item ? item->az : __ldg(q.azel + 2 * pix), el = item ? item->el : __ldg(q.azel + 2 * pix + 1);
    const bool need_pos = item ? item->pos >= 0 : true, need_neg = item ? item->neg >= 0 : false;
    if (tid == 0) {
        ff_build_frames(K_COEF, g, el, az, *fr);
        s_queue_walk = !q.rows;
        s_pred.on = 0;
        if (pix == 0 && q.clear_status) { q.clear_status[0] = 0; q.clear_status[1] = 0; }
    }
    __syncthreads();
    const int nth = q.ang.nth;
    if (q.rows) {
        // accepted interval [lo, hi) of every phi row: interval 0 of the row, the others empty
        for (int j = tid; j < q.ang.nph; j += FF_T) {
            int lo, hi;
            if (!ff_row_interval(sC, g, *fr, q.ang.st, q.ang.ct, nth, __ldg(q.ang.sp + j), __ldg(q.ang.cp + j), lo, hi))
                s_queue_walk = 1;                                // the row does not stay on one side of the plane's horizon
            for (int k = 0; k < FF_ROW_K; k++) { row_lo[FF_ROW_K * j + k] = lo; row_pre[FF_ROW_K * j + k + 1] = k ? 0 : hi - lo; }
        }
        __syncthreads();
        if (warp == 0) ff_prefix_rows(row_pre, nvr, lane);
        __syncthreads();
        // focal-position predictor of this pixel (see FfPred): sample rays spread over the dense accepted list
        const int n_list = row_pre[nvr];
        if (q.prefilter && !s_queue_walk && n_list >= 512) {
            float su = 0.0f, sv = 0.0f;
            double sX = 0.0, sZ = 0.0;
            if (tid < FF_PRED_SAMPLES) {
                const int v = ff_pred_sample_pos(tid, n_list);
                int a = 0, b = nvr - 1;                    // last interval with row_pre[.] <= v
                while (a < b) {
                    const int m = (a + b + 1) >> 1;
                    if (row_pre[m] <= v) a = m; else b = m - 1;
                }
                const int i = row_lo[a] + (v - row_pre[a]);
                a /= FF_ROW_K;                             // its phi row
                const vec3 d = launch_dir(__ldg(q.ang.st + i), __ldg(q.ang.ct + i), __ldg(q.ang.sp + a), __ldg(q.ang.cp + a));
                vec3 Pa;
                double t_ap, opl, fx, fz;
                int rst;
                ff_aperture(sC, g, *fr, d, t_ap, Pa);
                ff_ray_lean<double>(sC, *fr, d, t_ap, Pa, g.foc_y, opl, fx, fz, rst);
                const float sf = __ldg(q.ang.stf + i) * FF_PRED_SCALE;
                su = sf * __ldg(q.ang.cpf + a); sv = sf * __ldg(q.ang.spf + a);
                sX = fx - fr->centre[0]; sZ = fz - fr->centre[2];
                double* row = s_samp + tid * FF_PRED_LD;
                ff_pred_basis<double>((double)su, (double)sv, row);
                row[FF_PRED_NB] = sX; row[FF_PRED_NB + 1] = sZ;
            }
            __syncthreads();
            // cubic least-squares fit (ff_pred_fit, spread over the CTA): normal equations, entry (a, c) per thread
            for (int t = tid; t < FF_PRED_NB * FF_PRED_LD; t += FF_T) {
                const int a = t / FF_PRED_LD, c = t - a * FF_PRED_LD;
                const bool low = c < FF_PRED_NB && c < a;                         // symmetric part: the same sum as (c, a)
                s_A[t] = ff_pred_normal_entry(s_samp, FF_PRED_SAMPLES, low ? c : a, low ? a : c);
            }
            __syncthreads();
            if (warp == 0) {
                if (lane < FF_PRED_NB) s_diag[lane] = s_A[lane * FF_PRED_LD + lane];
                for (int k = 0; k < FF_PRED_NB - 1; k++) {
                    __syncwarp();
                    const int i = k + 1 + lane;
                    if (i < FF_PRED_NB) ff_pred_eliminate(s_A, k, i);
                }
                __syncwarp();
                if (lane == 0) ff_pred_backsub(s_A, s_diag, s_pred);
            }
            __syncthreads();
            if (tid < FF_PRED_SAMPLES) {
                float x, z;
                ff_pred_eval(s_pred, su, sv, x, z);
                float e = fmaxf(fabsf(x - (float)sX), fabsf(z - (float)sZ));
                if (!(e == e)) e = 1e30f;                   // a sample that is not finite switches the predictor off
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) e = fmaxf(e, __shfl_down_sync(0xffffffffu, e, o));
                if (lane == 0) s_pres[warp] = e;
            }
            __syncthreads();
            if (tid == 0) ff_pred_set_margin(s_pred, fmaxf(s_pres[0], s_pres[1]));
            __syncthreads();
            if (s_pred.on) {
                // keep only the rays that can matter to the pixel or its mirror pixel (ff_pred_row_keep): up to FF_ROW_K
                // intervals per row.  (Batches from the last row down: a batch replaces entries (K j0, K (j0 + FF_T)] and
                // the next one reads [.., K j0].)
                for (int j0 = ((q.ang.nph - 1) / FF_T) * FF_T; j0 >= 0; j0 -= FF_T) {
                    const int j = j0 + tid;
                    int klo[FF_ROW_K], kcnt[FF_ROW_K];
                    if (j < q.ang.nph) {
                        const int lo = row_lo[FF_ROW_K * j], hi = lo + (row_pre[FF_ROW_K * j + 1] - row_pre[FF_ROW_K * j]);
                        ff_pred_row_keep(s_pred, q.ang.stf, __ldg(q.ang.cpf + j), __ldg(q.ang.spf + j), q.grid, need_pos, need_neg,
                                         0.1f * (float)(fr->centre[2] - fr->centre_mz), lo, hi, klo, kcnt);
                    }
                    __syncthreads();                         // every count of this batch was read before any is replaced
                    if (j < q.ang.nph)
                        for (int k = 0; k < FF_ROW_K; k++) { row_lo[FF_ROW_K * j + k] = klo[k]; row_pre[FF_ROW_K * j + k + 1] = kcnt[k]; }
                }
                __syncthreads();
                if (warp == 0) ff_prefix_rows(row_pre, nvr, lane);
                __syncthreads();
                // ---- pile shrinking (FfPile): scouts, their exact landings, second pass over the rows
                if (T != FF_PREP_T && q.pile && row_pre[nvr] >= FF_PILE_MIN) {   // (compiled into the large-fan variant only)
                    const int M = q.grid.M, nph = q.ang.nph;
                    int* scout = row_pre + nvr + 1;                      // [2 tables][2 lines][2 M] latest ray predicted into the half cell
                    int* L1 = scout + 8 * M;                             // [2][2][M] latest scout that landed in the cell
                    int* L2 = L1 + 4 * M;                                // the same from the surviving scouts only
                    int* sc = L2 + 4 * M;                                // [2 tables][8 M] landing cell of every scout in both tables
                    for (int i = tid; i < 16 * M; i += FF_T) scout[i] = 0;   // scout, L1, L2
                    __syncthreads();
                    const float mg = s_pred.margin_cm, dz_m = 0.1f * (float)(fr->centre[2] - fr->centre_mz);
                    const float x_lo = (float)q.grid.gx_min - mg, z_lo = (float)q.grid.gy_min - mg, x_hi = (float)q.grid.gx_max + mg,
                                z_hi = (float)q.grid.gy_max + mg;
                    // pile line of a predicted landing point in a table: 0 column 0, 1 row 0, -1 none (inside, corner, beyond)
                    auto pile_line = [&](float X, float Z, float& c) {
                        if (X > x_hi || Z > z_hi) return -1;
                        const bool left = X < x_lo, below = Z < z_lo;
                        if (left == below) return -1;
                        c = left ? Z : X;
                        return left ? 0 : 1;
                    };
                    for (int j = tid; j < nph; j += FF_T) {
                        float ax[4], az[4];
                        ff_pred_row_cubic(s_pred.cx, __ldg(q.ang.cpf + j), __ldg(q.ang.spf + j), ax);
                        ff_pred_row_cubic(s_pred.cz, __ldg(q.ang.cpf + j), __ldg(q.ang.spf + j), az);
                        for (int k = 0; k < FF_ROW_K; k++) {
                            const int lo = row_lo[FF_ROW_K * j + k], cnt = row_pre[FF_ROW_K * j + k + 1] - row_pre[FF_ROW_K * j + k];
#pragma unroll 4
                            for (int i = lo + cnt - 1; i >= lo; i -= FF_SCOUT_STEP) {
                                const float sf = __ldg(q.ang.stf + i) * FF_PRED_SCALE;
                                const float X = 0.1f * fmaf(fmaf(fmaf(ax[3], sf, ax[2]), sf, ax[1]), sf, ax[0]);
                                const float Z = 0.1f * fmaf(fmaf(fmaf(az[3], sf, az[2]), sf, az[1]), sf, az[0]);
                                float c;
                                if (need_pos) {
                                    const int line = pile_line(X, Z, c);
                                    if (line >= 0)
                                        atomicMax(scout + line * 2 * M + ff_pile_slot(q.grid, line, c), j * nth + i + 1);
                                }
                                if (need_neg) {
                                    const int line = pile_line(-X, Z + dz_m, c);
                                    if (line >= 0)
                                        atomicMax(scout + (2 + line) * 2 * M + ff_pile_slot(q.grid, line, c),
                                                  (nph - 1 - j) * nth + i + 1);
                                }
                            }
                        }
                    }
                    __syncthreads();
                    // every landing of a scout on a pile line of either table: fn(table, line, cell on the line, index + 1)
                    auto landings = [&](int e, int cp, int cn, int i, int j, auto&& fn) {
                        (void)e;
                        if (cp >= 0) {
                            if (cp % M == 0) fn(0, 0, cp / M, j * nth + i + 1);
                            if (cp / M == 0) fn(0, 1, cp % M, j * nth + i + 1);
                        }
                        if (cn >= 0) {
                            if (cn % M == 0) fn(1, 0, cn / M, (nph - 1 - j) * nth + i + 1);
                            if (cn / M == 0) fn(1, 1, cn % M, (nph - 1 - j) * nth + i + 1);
                        }
                    };
                    auto scout_ray = [&](int e, int& i, int& j) {
                        const int sidx = scout[e] - 1;
                        j = sidx / nth; i = sidx - j * nth;
                        if (e >= 4 * M) j = nph - 1 - j;                  // the mirror table counts the rows backwards
                    };
                    for (int e = tid; e < 8 * M; e += FF_T) {             // trace the scouts
                        sc[e] = sc[8 * M + e] = -1;
                        if (scout[e] <= 0) continue;
                        int i, j;
                        scout_ray(e, i, j);
                        const vec3 d = launch_dir(__ldg(q.ang.st + i), __ldg(q.ang.ct + i), __ldg(q.ang.sp + j), __ldg(q.ang.cp + j));
                        vec3 Pa;
                        double t_ap, opl, fx, fz;
                        int rst;
                        ff_aperture(sC, g, *fr, d, t_ap, Pa);
                        ff_ray_lean<double>(sC, *fr, d, t_ap, Pa, g.foc_y, opl, fx, fz, rst);
                        if (opl != 0.0 && !not_finite(opl)) {
                            const double X = fx - fr->centre[0];
                            if (need_pos) sc[e] = ff_cell(X, fz - fr->centre[2], q.grid);
                            if (need_neg) sc[8 * M + e] = ff_cell(-X, fz - fr->centre_mz, q.grid);
                        }
                        landings(e, sc[e], sc[8 * M + e], i, j, [&](int t, int line, int c, int v) { atomicMax(L1 + (2 * t + line) * M + c, v); });
                    }
                    __syncthreads();
                    // a scout survives when it is the latest in the cell of ITS OWN line (then its row segment keeps it,
                    // see ff_pred_row_keep); only survivors may stand for a cell
                    for (int e = tid; e < 8 * M; e += FF_T) {
                        if (scout[e] <= 0) continue;
                        const int t = e / (4 * M), line = (e / (2 * M)) & 1, cell = sc[t * 8 * M + e];
                        if (cell < 0 || (line == 0 ? cell % M != 0 : cell / M != 0)) continue;
                        if (L1[(2 * t + line) * M + (line == 0 ? cell / M : cell % M)] != scout[e]) continue;
                        int i, j;
                        scout_ray(e, i, j);
                        landings(e, sc[e], sc[8 * M + e], i, j, [&](int t2, int l2, int c, int v) { atomicMax(L2 + (2 * t2 + l2) * M + c, v); });
                    }
                    __syncthreads();
                    FfPile pile;
                    pile.L = L2; pile.nth = nth; pile.nph = nph;
                    for (int j = tid; j < nph; j += FF_T) {
                        int lo, hi, klo[FF_ROW_K], kcnt[FF_ROW_K];
                        ff_row_interval(sC, g, *fr, q.ang.st, q.ang.ct, nth, __ldg(q.ang.sp + j), __ldg(q.ang.cp + j), lo, hi);
                        pile.j = j;
                        ff_pred_row_keep(s_pred, q.ang.stf, __ldg(q.ang.cpf + j), __ldg(q.ang.spf + j), q.grid, need_pos, need_neg, dz_m, lo, hi,
                                         klo, kcnt, &pile);
                        for (int k = 0; k < FF_ROW_K; k++) { row_lo[FF_ROW_K * j + k] = klo[k]; row_pre[FF_ROW_K * j + k + 1] = kcnt[k]; }
                    }
                    __syncthreads();
                    if (warp == 0) ff_prefix_rows(row_pre, nvr, lane);
                    __syncthreads();
                }
            }
        }
    }
    // the record
    unsigned char* rec = q.rec + (size_t)pix * q.stride;
    {
        const int* src = reinterpret_cast<const int*>(fr);
        int* dst = reinterpret_cast<int*>(rec);
        for (int i = tid; i < (int)(sizeof(FfFrames) / sizeof(int)); i += FF_T) dst[i] = src[i];
        int* head = reinterpret_cast<int*>(rec + FF_REC_FRAMES);
        if (tid == 0) { head[0] = s_queue_walk; head[1] = q.rows ? row_pre[nvr] : 0; head[2] = head[3] = 0; }
        if (q.rows) {
            int* rows = reinterpret_cast<int*>(rec + FF_REC_HEAD);
            for (int i = tid; i < 2 * nvr + 1; i += FF_T) rows[i] = row_lo[i];   // row_lo and row_pre are contiguous
        }
    }
}

// One CTA per work item (a pixel, or a mirror pair of pixels), warps fully autonomous inside the ray loop (no CTA
// barrier).
//   * Which rays are traced:
//     - interval walk (theta monotone, the reference's linspace fans): ff_prep_kernel has listed the rays that can
//       matter as up to FF_ROW_K index intervals per phi row (accepted by the aperture, not ruled out by the
//       focal-position predictor); the record's prefix sums turn them into a dense list and every warp walks its
//       contiguous share of it, 32 rays per step, last ray first.
//     - queue walk (any other fan, a pixel whose plane-hit denominator changes sign along a row, or PANEL): each warp
//       tests chunks of 32 candidate rays (FP32 aperture prefilter, FP64 near the rim) and compacts the survivors
//       into its own shared-memory queue (ballot + popc), popping a full batch whenever there is one.
//     Either way the expensive two-mirror trace always runs with full warps.
//   * the reference's "later ray overwrites the cell" becomes a 64-bit shared-memory atomicMax on a key that carries
//     the ray index above an exact fixed-point copy of the path length, taken only if no later ray holds the cell; a
//     pair item bins every ray twice (the pixel's table and the mirror pixel's);
//   * after one barrier the winners -- and only they -- get their phasor, are weighted by E_rec and reduced.
//   * PANEL (adjuster-error studies, ff_sim.py:61-70 with non-zero offsets): every ray carries the panel-model
//     columns the host derived with pan_mod.identify_panel; rays on no panel are dropped (saf:626-628,805-809), the
//     others see their panels' error polynomials inside both Newton solves.
//   * LEAN (plain maps whose winner tables sit in shared memory and whose E_rec axes are sorted, near-uniform and
//     shared by all columns -- every grid ff_sim.py builds): the binning of a ray runs on shared-memory addresses
//     only (the generic version pays for pointers that may be global: LD.E / ATOM.E + QSPC), finds its cell with a
//     branch-free searchsorted (interpolated guess, exact +-1 correction against the staged axis, sentinels at both
//     ends), tests "a later ray holds the cell" on the high word of the key alone, and raises the key with a CAS loop
//     seeded by nothing but that read.  Same cells, same keys, same maps; a third fewer instructions per traced ray.
template <int PANEL, int T, int MINB, bool LEAN = false>
__global__ void __launch_bounds__(T, MINB) ff_map_kernel(FfMapParams q)
{
    constexpr int FF_T = T, FF_WARPS = T / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cells = q.grid.M * q.grid.M;
    FfFrames* fr = reinterpret_cast<FfFrames*>(smem_raw);
    // winner table(s): in shared memory, or -- E_rec grids too large for that (M > 168) -- this CTA's slice of a
    // global scratch buffer (L2 resident: 64-bit atomicMax on global memory, slower but unbounded)
    unsigned long long* win_s = reinterpret_cast<unsigned long long*>(smem_raw + ((sizeof(FfFrames) + 15) / 16) * 16);
    unsigned long long* win = (!LEAN && q.win_g) ? q.win_g + (size_t)blockIdx.x * (q.items ? 2 : 1) * cells : win_s;
    unsigned long long* win_m = win + cells;                    // pair mode: winner table of the mirror pixel (-az, el)
    int* queues = reinterpret_cast<int*>(win_s + ((!LEAN && q.win_g) ? 0 : (q.items ? 2 : 1) * cells));   // [FF_WARPS][FF_WQ]
    // grid axes on chip, one sentinel either side (LEAN: -inf below, +inf above)
    double* s_gx = reinterpret_cast<double*>(queues + FF_WARPS * FF_WQ) + 1;   // [-1 .. M] data_x[0, :]
    double* s_gy = s_gx + q.grid.M + 2;                                    // [-1 .. M] data_y[:, 0] when all columns agree
    const int nvr = FF_ROW_K * q.ang.nph;                                  // virtual rows: FF_ROW_K intervals per phi row
    int* row_lo = reinterpret_cast<int*>(s_gy + q.grid.M + 1);             // [nvr] first ray to trace of an interval
    int* row_pre = row_lo + (q.rows ? nvr : 0);                            // [nvr + 1] rays to trace before the interval
    __shared__ int s_queue_walk;
    __shared__ double s_t2e[9];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pix = blockIdx.x;
    if (q.items && q.item0 + pix >= __ldg(q.counters)) return;   // pair mode launches one CTA per pixel, items are fewer
    if (q.redo && q.redo[pix] == 0) return;                      // second pass behind ff_map_interp_kernel: flagged items only
    const int out_pos = q.items ? q.items[q.item0 + pix].pos : pix, out_neg = q.items ? q.items[q.item0 + pix].neg : -1;
    const FfGeom g = to_geom(q.g);
    const Coef& sC = K_COEF;   // constant bank; a shared-memory copy measured 4 % slower (profiles/r01_variants.md)
    {   // this pixel's record from ff_prep_kernel
        const unsigned char* rec = q.rec + (size_t)pix * q.rec_stride;
        const int* src = reinterpret_cast<const int*>(rec);
        int* dst = reinterpret_cast<int*>(fr);
        for (int i = tid; i < (int)(sizeof(FfFrames) / sizeof(int)); i += FF_T) dst[i] = __ldg(src + i);
        if (tid == 0) s_queue_walk = __ldg(reinterpret_cast<const int*>(rec + FF_REC_FRAMES));
        if (PANEL != 1 && q.rows) {
            const int* rows = reinterpret_cast<const int*>(rec + FF_REC_HEAD);
            for (int i = tid; i < 2 * nvr + 1; i += FF_T) row_lo[i] = __ldg(rows + i);
        }
    }
    for (int c = tid; c < (q.items ? 2 : 1) * cells; c += FF_T) win[c] = 0ull;
    FfGrid grid = q.grid;                                   // grid axes on chip: the binning does 2-4 dependent reads per ray
    for (int i = tid; i < grid.M; i += FF_T) {
        s_gx[i] = __ldg(q.grid.gx + i);
        if (grid.y_uniform) s_gy[i] = __ldg(q.grid.gy + (size_t)i * grid.M);
    }
    if (LEAN && tid < 4) {
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        (tid & 2 ? s_gy : s_gx)[tid & 1 ? grid.M : -1] = tid & 1 ? inf : -inf;
    }
    grid.gx = s_gx;
    if (grid.y_uniform) { grid.gy = s_gy; grid.gy_cs = 0; grid.gy_rs = 1; }
    __syncthreads();
    if (PANEL) {
        if (tid < 9) {
            double v = fr->t2.a[tid];
            if (g.src[2] != 0 && tid % 3 == 2) v /= cos(atan(1.0 / 3.0));   // saf:828-829
            s_t2e[tid] = v;
        }
        __syncthreads();
    }
    const int nth = q.ang.nth, n_rays = nth * q.ang.nph;
    const bool queue_walk = PANEL == 1 || s_queue_walk != 0;

    int* wq = queues + warp * FF_WQ;
    const double cx = fr->centre[0], cz = fr->centre[2];
    int st = 0, cnt = 0;                                   // cnt: queue fill, identical in all lanes of the warp
    // Both walks run from the LAST ray of the fan down to the first: the later ray wins a cell (ffl:28), so the rays
    // met first are the likely winners and the crowded cells (every ray beyond the low edge of the E_rec grid piles
    // into row / column 0, ffl:22-23) settle at once -- the later, smaller ray indices only read the cell and move on.
    int chunk = (n_rays + 31) / 32 - 1 - warp;
    // interval walk: every warp owns a contiguous share of the pixel's dense list of rays to trace, 32 per step
    const int n_acc = queue_walk ? 0 : row_pre[nvr];
    int row = max(nvr - 1, 0);
    // warp w owns rays [v_bot, v_top] of the list
    const int per_warp = (n_acc + FF_WARPS - 1) / FF_WARPS;
    const int v_bot = min(warp * per_warp, n_acc);
    int v_top = min(v_bot + per_warp, n_acc) - 1;
    if (!queue_walk && v_top >= v_bot) {                  // row of the warp's first ray: last row with row_pre[row] <= v_top
        int a = 0, b = nvr - 1;
        while (a < b) {
            const int m = (a + b + 1) >> 1;
            if (row_pre[m] <= v_top) a = m; else b = m - 1;
        }
        row = a;
    }
    // bin one traced ray (i, j) into the winner table(s): cells first (a later ray may hold them already), then ONE key
    // for both pixels of a pair
    auto bin = [&](int i, int j, double opl, double fx, double fz) __attribute__((always_inline)) {
        if (opl != 0.0 && !not_finite(opl)) {                                   // ffl:10
            const int r0 = j * nth + i;
            int cell = -1, cell_m = -1, rm = 0;
            if (out_pos >= 0) {
                cell = ff_cell(fx - cx, fz - cz, grid);
                if (cell >= 0 && (win[cell] >> q.key_shift) >= (unsigned long long)(r0 + 1)) cell = -1;
            }
            if (out_neg >= 0) {
                // the mirror pixel (-az, el): this ray is its ray (nph-1-j, i), lands at (-x, z) about its own centre
                rm = (q.ang.nph - 1 - j) * nth + i;
                cell_m = ff_cell(cx - fx, fz - fr->centre_mz, grid);
                if (cell_m >= 0 && (win_m[cell_m] >> q.key_shift) >= (unsigned long long)(rm + 1)) cell_m = -1;
            }
            if (cell >= 0 || cell_m >= 0) {
                bool exact;
                const unsigned long long low = ff_key(-1, opl, q.key_base, q.key_inv_unit, exact, q.key_unit, q.key_shift);   // ray bits 0
                if (!exact) st |= ST_RANGE;
                if (cell >= 0) atomicMax(win + cell, low | ((unsigned long long)(r0 + 1) << q.key_shift));     // ffl:28: last write wins
                if (cell_m >= 0) atomicMax(win_m + cell_m, low | ((unsigned long long)(rm + 1) << q.key_shift));
            }
        }
    };
    // the same on shared-memory addresses only (LEAN, see above)
    const float lean_nf = (float)grid.M;
    const int kh_shift = q.key_shift - 32;                  // the ray index sits in the high word of a key (shift 40 or 44)
    auto lean_cell = [&](double xx, double yy, const unsigned long long* tab, unsigned rkey) __attribute__((always_inline)) {
        if (!(xx <= grid.gx_max && yy <= grid.gy_max)) return -1;                  // ffl:19
        const int ix = searchsorted_lean(s_gx, lean_nf, xx, grid.gx_min, grid.inv_dx);   // ffl:22
        const int iy = searchsorted_lean(s_gy, lean_nf, yy, grid.gy_min, grid.inv_dy);   // ffl:23
        const int cell = iy * grid.M + ix;
        // a later ray holds the cell already: its key's high word is at least this ray's (equal: never, one ray per table)
        const unsigned held = reinterpret_cast<const volatile unsigned*>(tab)[2 * cell + 1];
        return held >= rkey ? -1 : cell;
    };
    auto lean_raise = [&](unsigned long long* cellp, unsigned long long key) __attribute__((always_inline)) {
        unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(cellp);
        while (old < key) {                                                        // ffl:28: the later ray wins
            const unsigned long long prev = atomicCAS(cellp, old, key);
            if (prev == old) break;
            old = prev;
        }
    };
    auto bin_lean = [&](int i, int j, double opl, double fx, double fz) __attribute__((always_inline)) {
        if (opl == 0.0 || not_finite(opl) || not_finite(fx) || not_finite(fz)) return;      // ffl:10
        const unsigned r0 = (unsigned)(j * nth + i + 1), rm = (unsigned)((q.ang.nph - 1 - j) * nth + i + 1);
        const double xx = div10(fx - cx);                                           // ffl:12-13
        int cell = -1, cell_m = -1;
        if (out_pos >= 0) cell = lean_cell(xx, div10(fz - cz), win_s, r0 << kh_shift);
        // the mirror pixel (-az, el): this ray is its ray (nph-1-j, i) and lands at (-x, z) about its own centre
        if (out_neg >= 0) cell_m = lean_cell(-xx, div10(fz - fr->centre_mz), win_s + cells, rm << kh_shift);
        if ((cell & cell_m) >= 0) {
            bool exact;
            const unsigned long long low = ff_key(-1, opl, q.key_base, q.key_inv_unit, exact, q.key_unit, q.key_shift);   // ray bits 0
            if (!exact) st |= ST_RANGE;
            if (cell >= 0) lean_raise(win_s + cell, low | ((unsigned long long)r0 << q.key_shift));
            if (cell_m >= 0) lean_raise(win_s + cells + cell_m, low | ((unsigned long long)rm << q.key_shift));
        }
    };
    // trace ray (i, j) and bin it
    auto trace = [&](int i, int j) __attribute__((always_inline)) {
        const double2 tsc = __ldg(q.ang.tsc + i), psc = __ldg(q.ang.psc + j);
        const vec3 d = launch_dir(tsc.x, tsc.y, psc.x, psc.y);
        vec3 Pa;
        double t_ap;
        ff_aperture(sC, g, *fr, d, t_ap, Pa);
        double opl, fx, fz;
        int rst;
        if (PANEL == 1) {
            const size_t at = (size_t)pix * n_rays + (j * nth + i);
            const PanelErr e1 = load_panel(q.pan1, q.np1, __ldg(q.pidx1 + at)), e2 = load_panel(q.pan2, q.np2, __ldg(q.pidx2 + at));
            FfRay ray;
            ff_ray_trace<false, true>(sC, *fr, d, t_ap, Pa, g.foc_y, &e1, &e2, s_t2e, ray);
            opl = ray.opl; fx = ray.foc.x; fz = ray.foc.z; rst = ray.status;
        } else if (PANEL == 2) {
            // the reference's two passes (ff_sim.py:208-209): ideal mirrors -> which panels (saf:596-628), then the
            // trace with those panels' error polynomials; a ray on no panel of either model is dropped (saf:626-628,805-809)
            FfRay ray;
            ff_ray_trace<true, false>(sC, *fr, d, t_ap, Pa, g.foc_y, nullptr, nullptr, nullptr, ray);
            rst = ray.status;
            const int c1 = panel_lookup(q.lay1, ray.m1.x, ray.m1.y - sC.ff[2]);
            const int c2 = c1 >= 0 ? panel_lookup(q.lay2, ray.m2.x, ray.m2.z) : -1;
            opl = 0.0; fx = fz = 0.0;
            if (c2 >= 0) {
                const PanelErr e1 = load_panel(q.pan1, q.np1, c1), e2 = load_panel(q.pan2, q.np2, c2);
                ff_ray_trace<false, true>(sC, *fr, d, t_ap, Pa, g.foc_y, &e1, &e2, s_t2e, ray);
                opl = ray.opl; fx = ray.foc.x; fz = ray.foc.z; rst |= ray.status;
            }
        } else {
            ff_ray_lean<double>(sC, *fr, d, t_ap, Pa, g.foc_y, opl, fx, fz, rst);
        }
        st |= rst;
        if (LEAN) bin_lean(i, j, opl, fx, fz); else bin(i, j, opl, fx, fz);
    };
    if (PANEL == 0 && !queue_walk) {
        // interval walk, the hot loop of the plain map on its own (no queue state alive): 32 rays of the warp's share of
        // the dense list per step
        for (; v_top >= v_bot; v_top -= 32) {
            const int v = v_top - lane;
            if (v >= v_bot) {
                while (row_pre[row] > v) --row;               // the walk only moves down the list, so does the row cursor
                trace(row_lo[row] + (v - row_pre[row]), row / FF_ROW_K);
            }
        }
    } else {
        for (;;) {
            int ti = -1, tj = 0;
            if (!queue_walk) {
                if (v_top < v_bot) break;
                const int v = v_top - lane;
                if (v >= v_bot) {
                    while (row_pre[row] > v) --row;
                    ti = row_lo[row] + (v - row_pre[row]); tj = row / FF_ROW_K;
                }
                v_top -= 32;
            } else {
                const bool more = chunk >= 0;
                if (cnt < 32 && more) {
                    // fill: aperture test (saf:359-379) on 32 candidate rays
                    const int r = chunk * 32 + lane;
                    bool ok = false;
                    if (r < n_rays) {
                        // FP32 prefilter on the FMA pipe; only rays within ~0.02 mm of the rim need the FP64 test
                        int i, j;
                        ray_ij(q.ang, r, i, j);
                        const float stf = __ldg(q.ang.stf + i);
                        const int c = ff_aperture_coarse(*fr, stf * __ldg(q.ang.cpf + j), stf * __ldg(q.ang.spf + j), __ldg(q.ang.ctf + i));
                        ok = c > 0;
                        if (c < 0) {
                            vec3 Pa;
                            double t_ap;
                            ok = ff_aperture(sC, g, *fr, ray_dir(q.ang, r), t_ap, Pa);
                        }
                        if (PANEL == 1) {
                            const size_t at = (size_t)pix * n_rays + r;
                            ok = ok && __ldg(q.pidx1 + at) >= 0 && __ldg(q.pidx2 + at) >= 0;
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (ok) wq[cnt + __popc(m & ((1u << lane) - 1u))] = r;
                    cnt += __popc(m);
                    chunk -= FF_WARPS;
                    __syncwarp();
                    continue;
                }
                if (cnt == 0) break;
                const int n_take = min(cnt, 32);
                const int r0 = lane < n_take ? wq[cnt - 1 - lane] : -1;
                cnt -= n_take;
                __syncwarp();
                if (r0 >= 0) ray_ij(q.ang, r0, ti, tj);
            }
            if (ti >= 0) trace(ti, tj);
        }
    }
    __syncthreads();
    // winners-only phasor sum: sum_cells exp(i k L/1e3) * E_rec / M^2   (ff_sim.py:215-221).  Only a fraction of the
    // cells holds a winner, so each warp compacts its winners through its queue and evaluates phasors with full lanes.
    if (st) atomicOr(q.status, st);
    __shared__ double2 total;
    for (int side = 0; side < 2; side++) {
        const int out = side ? out_neg : out_pos;
        if (out < 0) continue;                                  // CTA-uniform
        const unsigned long long* tab = LEAN ? win_s + (side ? cells : 0) : (side ? win_m : win);
        const bool tab_g = !LEAN && q.win_g;
        double re[1] = {0.0}, im[1] = {0.0};
        cnt = 0;
        for (int c0 = warp * 32;; c0 += FF_WARPS * 32) {
            const bool more = c0 < cells;
            if (more) {
                const int c = c0 + lane;
                // (a table in global memory was updated with atomics at L2: read it there, not through a stale L1 line)
                const bool has = c < cells && (tab_g ? __ldcg(tab + c) : tab[c]) != 0ull;
                const unsigned m = __ballot_sync(0xffffffffu, has);
                if (has) wq[cnt + __popc(m & ((1u << lane) - 1u))] = c;
                cnt += __popc(m);
                __syncwarp();
            }
            if (cnt >= 32 || (!more && cnt > 0)) {
                const int n_take = min(cnt, 32);
                const int c = lane < n_take ? wq[cnt - 1 - lane] : -1;
                cnt -= n_take;
                __syncwarp();
                if (c >= 0) {
                    double cs, sn;
                    cis_large(q.k * ff_key_opl(tab_g ? __ldcg(tab + c) : tab[c], q.key_base, q.key_unit, q.key_shift) / 1e3, cs, sn);
                    const double2 e = __ldg(q.e_rec + c);
                    re[0] += cs * e.x - sn * e.y;
                    im[0] += cs * e.y + sn * e.x;
                }
            }
            if (!more && cnt == 0) break;
        }
        block_reduce_store<1>(re, im, &total);
        if (tid == 0) {
            const double inv = 1.0 / ((double)q.grid.M * (double)q.grid.M);
            const double a = total.x * inv, b = total.y * inv;
            q.amp[out] = hypot(a, b);
            q.phase[out] = atan2(b, a);
        }
        __syncthreads();                                        // `total` and the reduction scratch are reused
    }
}


// ------------------------------------------------------------------------------------------------
// Large fans: the cells come from INTERPOLATED focal positions, only the winners are traced exactly.
//
// The reference keeps, per E_rec cell, the last ray that lands in it (ffl:22-28): of the ~177 000 rays a config-5 pair
// item lists, at most 2 x 101^2 = 20 402 count.  Which cell a ray lands in is a smooth function of its fan indices
// (i, j): cubic interpolation between exactly traced rays 16 apart is good to 2e-4 mm (measured against exact traces
// on a 1000^2 fan) against cells of 1 mm.  The lattice, the row nodes and the weights are kept in single precision and
// in CELL UNITS u = (x - gx[0]) / step of the pixel's own table (|u| < 2^9: 3e-5 per rounding), so that a ray's cell is
// ceil(u) and its distance to the nearest edge |u - rint(u)|, without a look at the axis.  So:
//   1a  the CTA traces a lattice of rays, every S-th in both fan directions (one ring beyond the fan on either side,
//       so that every ray has a full 4 x 4 stencil; the fan's axes are uniform, the host checks it): 66^2 rays;
//   1b  a warp takes a phi row: row nodes by cubic interpolation across the four lattice rows around it, then, for
//       every LISTED ray of the row (the intervals ff_prep_kernel left), the focal position by cubic interpolation
//       along the row -> cell of the pixel's table and of the mirror pixel's.  A ray more than eps = 3e-3 mm inside
//       a cell offers its index to the cell (32-bit atomicMax: ffl:28, the later ray wins); a ray surely beyond the
//       high edges is dropped (ffl:19); a ray within eps of an edge is AMBIGUOUS: it is traced exactly and offered to
//       the cells it really lands in;
//   2   every cell's winner is traced exactly: path length -> phasor x E_rec, summed exactly as in ff_map_kernel (same
//       scan order, same compaction, same lanes: identical bits).  Each winner's exact landing point is compared with
//       its prediction: it must sit in the predicted cell and within chk = eps / 3 of the predicted position -- 20 000
//       probes per item, one in every cell, of an error field that is smooth on the scale of S rays.  An item that
//       fails any of them (or has no row intervals) is flagged and ff_map_kernel, launched behind this kernel, redoes it.
// Exact by construction as long as the interpolation error stays below eps, which is what the probes watch; the maps
// are bit-identical to the exact walk (tests: test_ff_interp_*).  A ray that is never traced reports no status.
constexpr int FF_INTERP_T = 768;
constexpr int FF_INTERP_MIN_N = 400;           // smaller fans keep the exact walk (measured: no gain at 300 x 300 rays)
constexpr double FF_INTERP_EPS_CM = 3.0e-4;    // 3e-3 mm: ten times the interpolation error measured at S = 16 on a 1000^2 fan
struct FfInterpLayout {
    size_t axes, lat, rown, wts, cand, queues, rows, total;
    int nk;
};
__host__ __device__ inline FfInterpLayout ff_interp_layout(int M, int nth, int nph, int S, int tables)
{
    FfInterpLayout L;
    L.nk = (nth - 1) / S + 4;                      // lattice nodes k = -1 .. (nth-1)/S + 2 per axis (nth == nph)
    size_t o = ((sizeof(FfFrames) + 15) / 16) * 16;
    L.axes = o;   o += 2 * ((size_t)M + 2) * sizeof(double);
    o = (o + 15) / 16 * 16;
    L.lat = o;    o += ((size_t)L.nk * L.nk * sizeof(float2) + 15) / 16 * 16;
    L.rown = o;   o += ((size_t)(FF_INTERP_T / 32) * L.nk * sizeof(float2) + 15) / 16 * 16;
    L.wts = o;    o += (size_t)S * 4 * sizeof(float);
    L.cand = o;   o += ((size_t)tables * M * M * sizeof(unsigned) + 15) / 16 * 16;
    L.queues = o; o += (size_t)(FF_INTERP_T / 32) * FF_WQ * sizeof(int);
    L.rows = o;   o += (2 * (size_t)FF_ROW_K * nph + 1) * sizeof(int);
    L.total = o;
    return L;
}
__global__ void __launch_bounds__(FF_INTERP_T, 1) ff_map_interp_kernel(FfMapParams q)
{
    constexpr int T = FF_INTERP_T, WARPS = T / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = q.grid.M, cells = M * M, S = q.interp_s;
    const int nth = q.ang.nth, nph = q.ang.nph;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pix = blockIdx.x;
    if (q.items && q.item0 + pix >= __ldg(q.counters)) return;
    const int out_pos = q.items ? q.items[q.item0 + pix].pos : pix, out_neg = q.items ? q.items[q.item0 + pix].neg : -1;
    const int tables = q.items ? 2 : 1;
    const FfInterpLayout L = ff_interp_layout(M, nth, nph, S, tables);
    const int NK = L.nk;
    FfFrames* fr = reinterpret_cast<FfFrames*>(smem_raw);
    double* s_gx = reinterpret_cast<double*>(smem_raw + L.axes) + 1;
    double* s_gy = s_gx + M + 2;
    // [NK][NK] (phi node, theta node): focal-plane position in CELL UNITS of the pixel's table, u = (x/10 - gx[0]) / step
    float2* lat = reinterpret_cast<float2*>(smem_raw + L.lat);
    float2* rown = reinterpret_cast<float2*>(smem_raw + L.rown) + warp * NK;     // this warp's row nodes
    float4* wts = reinterpret_cast<float4*>(smem_raw + L.wts);              // [S] cubic Lagrange weights at t = 1 + f / S
    unsigned* cand = reinterpret_cast<unsigned*>(smem_raw + L.cand);        // [tables][cells] ray index + 1 of the latest ray
    int* wq = reinterpret_cast<int*>(smem_raw + L.queues) + warp * FF_WQ;
    const int nvr = FF_ROW_K * nph;
    int* row_lo = reinterpret_cast<int*>(smem_raw + L.rows);
    int* row_pre = row_lo + nvr;
    __shared__ int s_queue_walk, s_row, s_bad, s_err, s_hull[4];   // s_hull: min / max listed theta index, min / max listed phi row
    __shared__ double2 total;

    const FfGeom g = to_geom(q.g);
    const Coef& sC = K_COEF;
    {
        const unsigned char* rec = q.rec + (size_t)pix * q.rec_stride;
        const int* src = reinterpret_cast<const int*>(rec);
        int* dst = reinterpret_cast<int*>(fr);
        for (int i = tid; i < (int)(sizeof(FfFrames) / sizeof(int)); i += T) dst[i] = __ldg(src + i);
        if (tid == 0) {
            s_queue_walk = __ldg(reinterpret_cast<const int*>(rec + FF_REC_FRAMES)); s_row = nph - 1; s_bad = 0; s_err = 0;
            s_hull[0] = s_hull[2] = 0x7fffffff; s_hull[1] = s_hull[3] = -1;
        }
        const int* rows = reinterpret_cast<const int*>(rec + FF_REC_HEAD);
        for (int i = tid; i < 2 * nvr + 1; i += T) row_lo[i] = __ldg(rows + i);
    }
    for (int c = tid; c < tables * cells; c += T) cand[c] = 0u;
    for (int i = tid; i < M; i += T) {
        s_gx[i] = __ldg(q.grid.gx + i);
        s_gy[i] = __ldg(q.grid.gy + (size_t)i * M);
    }
    if (tid < 4) {
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        (tid & 2 ? s_gy : s_gx)[tid & 1 ? M : -1] = tid & 1 ? inf : -inf;
    }
    for (int f = tid; f < S; f += T) {
        const double t = 1.0 + (double)f / (double)S;
        wts[f] = make_float4((float)(-(t - 1.0) * (t - 2.0) * (t - 3.0) / 6.0), (float)(t * (t - 2.0) * (t - 3.0) / 2.0),
                             (float)(-t * (t - 1.0) * (t - 3.0) / 2.0), (float)(t * (t - 1.0) * (t - 2.0) / 6.0));
    }
    __syncthreads();
    if (s_queue_walk) {                                        // no row intervals for this pixel: the exact kernel does it
        if (tid == 0) q.redo[pix] = 1;
        return;
    }
    const double cx = fr->centre[0], cz = fr->centre[2];
    const double dzm = div10(cz - fr->centre_mz);              // the mirror pixel's z offset is this much larger [cm]
    const float nf = (float)M;
    // cell units: u = (v - a[0]) * inv_step; the mirror pixel's ray lands at (-x, z + dzm): u_x' = umx - u_x, u_y' = u_y + umy
    const double inv_x = (double)q.grid.inv_dx, inv_y = (double)q.grid.inv_dy;
    __shared__ float s_umy;          // the mirror pixel's z offset in cell units (the only item-dependent constant of the classifier)
    if (tid == 0) s_umy = (float)(dzm * inv_y);
    const int ls = q.interp_ls;                               // S = 1 << ls
    int st = 0;

    // exact trace of ray (i, j): path length and focal-plane offsets / 10 about the pixel's centre (ffl:12-13)
    auto trace = [&](int i, int j, double& opl, double& xx, double& yy, double& yym) __attribute__((always_inline)) {
        const double2 tsc = __ldg(q.ang.tsc + i), psc = __ldg(q.ang.psc + j);
        const vec3 d = launch_dir(tsc.x, tsc.y, psc.x, psc.y);
        vec3 Pa;
        double t_ap, fx, fz;
        int rst;
        ff_aperture(sC, g, *fr, d, t_ap, Pa);
        ff_ray_lean<double>(sC, *fr, d, t_ap, Pa, g.foc_y, opl, fx, fz, rst);
        st |= rst;
        xx = div10(fx - cx); yy = div10(fz - cz); yym = div10(fz - fr->centre_mz);
        return !(opl == 0.0 || not_finite(opl) || not_finite(fx) || not_finite(fz));   // ffl:10
    };
    // exact cell of a traced ray, -1 when ffl:19 skips it
    auto cell_of = [&](double xx, double yy) __attribute__((always_inline)) {
        if (!(xx <= q.grid.gx_max && yy <= q.grid.gy_max)) return -1;
        return searchsorted_lean(s_gy, nf, yy, q.grid.gy_min, q.grid.inv_dy) * M + searchsorted_lean(s_gx, nf, xx, q.grid.gx_min, q.grid.inv_dx);
    };
    auto offer = [&](unsigned* tab, int cell, unsigned r) __attribute__((always_inline)) {
        if (*reinterpret_cast<volatile unsigned*>(tab + cell) < r) atomicMax(tab + cell, r);   // ffl:28
    };

    // ---- 1a: the lattice, as far as the stencils of the listed rays reach
    {
        int i0 = 0x7fffffff, i1 = -1, j0 = 0x7fffffff, j1 = -1;
        for (int vr = tid; vr < nvr; vr += T) {
            const int c = row_pre[vr + 1] - row_pre[vr];
            if (c > 0) { const int lo = row_lo[vr], j = vr / FF_ROW_K; i0 = min(i0, lo); i1 = max(i1, lo + c - 1); j0 = min(j0, j); j1 = max(j1, j); }
        }
        if (i1 >= 0) { atomicMin(&s_hull[0], i0); atomicMax(&s_hull[1], i1); atomicMin(&s_hull[2], j0); atomicMax(&s_hull[3], j1); }
        __syncthreads();
    }
    const int hk0 = s_hull[0] >> ls, hk1 = (s_hull[1] >> ls) + 3, hm0 = s_hull[2] >> ls, hm1 = (s_hull[3] >> ls) + 3;   // node indices (array)
    const int hw = max(hk1 - hk0 + 1, 0), hn = hw * max(hm1 - hm0 + 1, 0);     // (an item without a listed ray: no node)
    for (int nn = tid; nn < hn; nn += T) {
        const int mm = hm0 + nn / hw, kk = hk0 + nn % hw, n = mm * NK + kk;
        const int i = (kk - 1) * S, j = (mm - 1) * S;
        double sth, cth, sph, cph;
        if (i >= 0 && i < nth) { const double2 v = __ldg(q.ang.tsc + i); sth = v.x; cth = v.y; }
        else sincos(q.th0 + (double)i * q.dth, &sth, &cth);
        if (j >= 0 && j < nph) { const double2 v = __ldg(q.ang.psc + j); sph = v.x; cph = v.y; }
        else sincos(q.ph0 + (double)j * q.dph, &sph, &cph);
        const vec3 d = launch_dir(sth, cth, sph, cph);
        vec3 Pa;
        double t_ap, opl, fx, fz;
        int rst;
        ff_aperture(sC, g, *fr, d, t_ap, Pa);
        ff_ray_lean<double>(sC, *fr, d, t_ap, Pa, g.foc_y, opl, fx, fz, rst);   // (a node's status is nobody's: it may lie outside the aperture)
        // a node whose trace did not end well (Newton not converged, a root outside its bracket, not finite) is NaN: every
        // ray whose stencil holds it becomes ambiguous and is traced exactly, instead of being placed by a garbage value
        const float qnan = __int_as_float(0x7fc00000);
        const bool good = rst == 0 && !not_finite(fx) && !not_finite(fz);
        lat[n] = good ? make_float2((float)((div10(fx - cx) - q.grid.gx_min) * inv_x), (float)((div10(fz - cz) - q.grid.gy_min) * inv_y))
                      : make_float2(qnan, qnan);
    }
    __syncthreads();

    // ---- 1b: rows, last row first
    int cnt = 0;                                              // queue of ambiguous rays (ray index j * nth + i), warp-uniform
    auto exact_offer = [&](int r) __attribute__((always_inline)) {
        const int j = r / nth, i = r - j * nth;
        double opl, xx, yy, yym;
        if (trace(i, j, opl, xx, yy, yym)) {
            if (out_pos >= 0) { const int c = cell_of(xx, yy); if (c >= 0) offer(cand, c, (unsigned)(r + 1)); }
            if (out_neg >= 0) { const int c = cell_of(-xx, yym); if (c >= 0) offer(cand + cells, c, (unsigned)((nph - 1 - j) * nth + i + 1)); }
        }
    };
    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicSub(&s_row, 1);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j < 0) break;
        const int vr0 = FF_ROW_K * j, base = row_pre[vr0], n_row = row_pre[vr0 + FF_ROW_K] - base;
        if (n_row <= 0) continue;
        int p[FF_ROW_K + 1], lo[FF_ROW_K];
        int imin = 0x7fffffff, imax = -1;
#pragma unroll
        for (int k = 0; k < FF_ROW_K; k++) {
            p[k] = row_pre[vr0 + k] - base; lo[k] = row_lo[vr0 + k];
            const int c = row_pre[vr0 + k + 1] - row_pre[vr0 + k];
            if (c > 0) { imin = min(imin, lo[k]); imax = max(imax, lo[k] + c - 1); }
        }
        {   // row nodes the listed rays reach: theta nodes imin / S - 1 .. imax / S + 2 (array index + 1)
            const int m0 = j >> ls;
            const float4 wj = wts[j - (m0 << ls)];
            for (int k = (imin >> ls) + lane; k <= (imax >> ls) + 3; k += 32) {
                const float2 a0 = lat[m0 * NK + k], a1 = lat[(m0 + 1) * NK + k], a2 = lat[(m0 + 2) * NK + k], a3 = lat[(m0 + 3) * NK + k];
                rown[k] = make_float2(fmaf(wj.x, a0.x, fmaf(wj.y, a1.x, fmaf(wj.z, a2.x, wj.w * a3.x))),
                                      fmaf(wj.x, a0.y, fmaf(wj.y, a1.y, fmaf(wj.z, a2.y, wj.w * a3.y))));
            }
        }
        __syncwarp();
        const unsigned r_row = (unsigned)(j * nth + 1), rm_row = (unsigned)((nph - 1 - j) * nth + 1);
        for (int v_top = n_row - 1; v_top >= 0; v_top -= 32) {
            const int v = v_top - lane;
            bool amb = false;
            int i = 0;
            if (v >= 0) {
                // the interval v falls in (prefix counts ascend; an empty interval never wins the comparison for long)
                int d = lo[0] - p[0];
#pragma unroll
                for (int t = 1; t < FF_ROW_K; t++) if (v >= p[t]) d = lo[t] - p[t];
                i = v + d;
                const int kb = i >> ls;
                const float4 wi = wts[i - (kb << ls)];
                const float2 n0 = rown[kb], n1 = rown[kb + 1], n2 = rown[kb + 2], n3 = rown[kb + 3];
                const float ux = fmaf(wi.x, n0.x, fmaf(wi.y, n1.x, fmaf(wi.z, n2.x, wi.w * n3.x)));
                const float uy = fmaf(wi.x, n0.y, fmaf(wi.y, n1.y, fmaf(wi.z, n2.y, wi.w * n3.y)));
                const bool nan = !(ux + uy == ux + uy);          // a stencil node that could not be traced
                if (out_pos >= 0) {
                    const int c = ff_cell_class(ux, uy, q.i_uhi, q.i_tol_x, q.i_tol_y, M);
                    if (c >= 0 && !nan) offer(cand, c, r_row + (unsigned)i);
                    amb = c == -2;
                }
                if (out_neg >= 0) {
                    const int c = ff_cell_class(q.i_umx - ux, uy + s_umy, q.i_uhi, q.i_tol_x, q.i_tol_y, M);
                    if (c >= 0 && !nan) offer(cand + cells, c, rm_row + (unsigned)i);
                    amb = amb || c == -2;
                }
                amb = amb || nan;
            }
            const unsigned m = __ballot_sync(0xffffffffu, amb);
            if (m) {
                if (amb) wq[cnt + __popc(m & ((1u << lane) - 1u))] = j * nth + i;
                cnt += __popc(m);
                __syncwarp();
                if (cnt >= 32) {
                    const int r = wq[cnt - 1 - lane];
                    cnt -= 32;
                    __syncwarp();
                    exact_offer(r);
                }
            }
        }
        __syncwarp();                                         // the row nodes are overwritten by the next row
    }
    if (cnt > 0) {
        const int r = lane < cnt ? wq[cnt - 1 - lane] : -1;
        __syncwarp();
        if (r >= 0) exact_offer(r);
        cnt = 0;
    }
    __syncthreads();

    // ---- 2: the winners, traced exactly; sum_cells exp(i k L/1e3) * E_rec / M^2   (ff_sim.py:215-221)
    int bad = 0;
    float worst = 0.0f;
    for (int side = 0; side < 2; side++) {
        const int out = side ? out_neg : out_pos;
        if (out < 0) continue;                                  // CTA-uniform
        const unsigned* tab = cand + (side ? cells : 0);
        double re[1] = {0.0}, im[1] = {0.0};
        cnt = 0;
        for (int c0 = warp * 32;; c0 += WARPS * 32) {
            const bool more = c0 < cells;
            if (more) {
                const int c = c0 + lane;
                const bool has = c < cells && tab[c] != 0u;
                const unsigned m = __ballot_sync(0xffffffffu, has);
                if (has) wq[cnt + __popc(m & ((1u << lane) - 1u))] = c;
                cnt += __popc(m);
                __syncwarp();
            }
            if (cnt >= 32 || (!more && cnt > 0)) {
                const int n_take = min(cnt, 32);
                const int c = lane < n_take ? wq[cnt - 1 - lane] : -1;
                cnt -= n_take;
                __syncwarp();
                if (c >= 0) {
                    const int r = (int)tab[c] - 1;
                    const int jr = r / nth, i = r - jr * nth, j = side ? nph - 1 - jr : jr;
                    double opl, xx, yy, yym;
                    const bool ok = trace(i, j, opl, xx, yy, yym);
                    // the prediction this winner was chosen by: bicubic on the lattice
                    const int m0 = j >> ls, kb = i >> ls;
                    const float4 wj = wts[j - (m0 << ls)], wi = wts[i - (kb << ls)];
                    const float wja[4] = {wj.x, wj.y, wj.z, wj.w};
                    float ui = 0.0f, vi = 0.0f;
#pragma unroll
                    for (int a = 0; a < 4; a++) {
                        const float2* row = lat + (m0 + a) * NK + kb;
                        const float2 n0 = row[0], n1 = row[1], n2 = row[2], n3 = row[3];
                        ui = fmaf(wja[a], fmaf(wi.x, n0.x, fmaf(wi.y, n1.x, fmaf(wi.z, n2.x, wi.w * n3.x))), ui);
                        vi = fmaf(wja[a], fmaf(wi.x, n0.y, fmaf(wi.y, n1.y, fmaf(wi.z, n2.y, wi.w * n3.y))), vi);
                    }
                    const double xs = side ? -xx : xx, ys = side ? yym : yy;
                    if (side) { ui = q.i_umx - ui; vi += s_umy; }
                    const float ue = (float)((xs - q.grid.gx_min) * inv_x), ve = (float)((ys - q.grid.gy_min) * inv_y);
                    // (no prediction where a stencil node could not be traced: that ray was placed by its exact trace)
                    const bool pred = ui + vi == ui + vi;
                    const bool near_x = !pred || fabsf(ue - ui) <= q.i_chk_x, near_y = !pred || fabsf(ve - vi) <= q.i_chk_y;
                    if (pred) worst = fmaxf(worst, fmaxf(fabsf(ue - ui), fabsf(ve - vi)));
                    bad |= (ok ? 0 : 1) | (near_x ? 0 : 2) | (near_y ? 0 : 4) | (cell_of(xs, ys) == c ? 0 : 8);
                    double cs, sn;
                    cis_large(q.k * opl / 1e3, cs, sn);
                    const double2 e = __ldg(q.e_rec + c);
                    re[0] += cs * e.x - sn * e.y;
                    im[0] += cs * e.y + sn * e.x;
                }
            }
            if (!more && cnt == 0) break;
        }
        block_reduce_store<1>(re, im, &total);
        if (tid == 0) {
            const double inv = 1.0 / ((double)M * (double)M);
            const double a = total.x * inv, b = total.y * inv;
            q.amp[out] = hypot(a, b);
            q.phase[out] = atan2(b, a);
        }
        __syncthreads();
    }
    if (bad) atomicOr(&s_bad, bad);
    if (q.probe_err) atomicMax(&s_err, __float_as_int(worst));     // (non-negative floats order like their bit patterns)
    __syncthreads();
    if (tid == 0 && q.probe_err) q.probe_err[pix] = __int_as_float(s_err);
    if (tid == 0 && s_bad) q.redo[pix] = s_bad;
    // the status of an item that is redone comes from the exact kernel (which traces a superset of these rays)
    if (st) atomicOr(q.status, st);
}

// What the reference hands to pan_mod.identify_panel for every pixel (saf:596-612 = rows 0, 2, 3, 4 of ray_mirror_pts'
// output): telescope-frame M2 hit (x, z) and M1 hit (x, y) on the error-free mirrors, 0 where the ray misses the
// aperture.  xy[(pix * 4 + k) * n_rays + r].
__global__ void __launch_bounds__(128) ff_mirror_xy_kernel(AngleTabs ang, solat_ff_geom gg, const double* azel, double* xy, int* status)
{
    __shared__ FfFrames fr;
    const FfGeom g = to_geom(gg);
    const int pix = blockIdx.y;
    if (threadIdx.x == 0) ff_build_frames(K_COEF, g, __ldg(azel + 2 * pix + 1), __ldg(azel + 2 * pix), fr);
    __syncthreads();
    const int n = ang.rays();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const vec3 d = ray_dir(ang, r);
    vec3 Pa;
    double t_ap;
    if (ff_aperture(K_COEF, g, fr, d, t_ap, Pa)) {
        FfRay ray;
        ff_ray_trace<true, false>(K_COEF, fr, d, t_ap, Pa, g.foc_y, nullptr, nullptr, nullptr, ray);
        v[0] = ray.m2.x; v[1] = ray.m2.z; v[2] = ray.m1.x; v[3] = ray.m1.y;
        if (ray.status) atomicOr(status, ray.status);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) xy[((size_t)pix * 4 + k) * n + r] = v[k];
}

// The interval walk needs the rays of a phi row to move monotonically along their line in the aperture plane:
// theta strictly monotone, inside one branch of tan (cos theta keeps its sign, never ~0), steps far above the
// rounding of the radius test.
static bool ff_rows_usable(const double* theta, int nth, int nph)
{
    if (nph <= 0 || nph > FF_MAX_ROWS || nth < 2) return false;
    const char* env = getenv("SOLAT_FF_QUEUE_WALK");
    if (env && env[0] == '1') return false;
    const double dir = theta[1] - theta[0];
    for (int i = 0; i < nth; i++) {
        const double c = cos(theta[i]);
        if (!(c * cos(theta[0]) > 0.0) || !(fabs(c) > 1e-6)) return false;
        if (i && !((theta[i] - theta[i - 1]) * dir > 0.0 && fabs(theta[i] - theta[i - 1]) >= 1e-10)) return false;
    }
    return true;
}

static size_t ff_map_smem(int M, int rows_nph, int threads, int tables)
{
    const size_t cells = (size_t)M * M * tables;
    return ((sizeof(FfFrames) + 15) / 16) * 16 + cells * sizeof(unsigned long long) + (size_t)(threads / 32) * FF_WQ * sizeof(int) +
           2 * ((size_t)M + 2) * sizeof(double) + (rows_nph ? (2 * (size_t)FF_ROW_K * rows_nph + 1) * sizeof(int) : 0);
}

// ------------------------------------------------------------------------------------------------
// regrid: interpolating bicubic spline resample in ray-index space (ray_trace_int.py:1385-1440)
// ------------------------------------------------------------------------------------------------
// The reference resamples a (n_old x n_old) field given on linspace(-70, 70, n_old) onto linspace(-70, 70, n_new)
// with interp2d(kind="cubic") = FITPACK's interpolating bicubic spline (s = 0), whose knot choice is the
// not-a-knot one.  A tensor-product interpolant is separable: a 1-D not-a-knot cubic spline along x for every row,
// then along y for every new column.  One thread per 1-D problem; second derivatives M by the Thomas algorithm.
__global__ void regrid_rows_kernel(const double* field, int n_old, int n_new, double* tmp, double* scratch)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;      // old row: spline along the fast axis
    if (r >= n_old) return;
    double* M = scratch + (size_t)r * 2 * n_old;
    spline_line(field + (size_t)r * n_old, 1, n_old, M, M + n_old, n_new, tmp + (size_t)r * n_new, 1);
}
__global__ void regrid_cols_kernel(const double* tmp, int n_old, int n_new, double* out, double* scratch)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;      // new column: spline along the slow axis
    if (c >= n_new) return;
    double* M = scratch + (size_t)c * 2 * n_old;
    spline_line(tmp + c, n_new, n_old, M, M + n_old, n_new, out + c, n_new);
}

// ------------------------------------------------------------------------------------------------
// FP64 peak microbenchmark
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// accuracy probe of the fast-math primitives: out[v][i], v = rcp with 0..3 Newton steps, rsqrt with 0..1 polish,
// cos/sin of the large-argument reducer, cos/sin of the mid-range routine
__global__ void selftest_math_kernel(const double* x, int n, double* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = x[i];
    out[0 * n + i] = rcp_n<0>(a);
    out[1 * n + i] = rcp_n<1>(a);
    out[2 * n + i] = rcp_n<2>(a);
    out[3 * n + i] = rcp_n<3>(a);
    out[4 * n + i] = rsqrt_n<0>(a);
    out[5 * n + i] = rsqrt_n<1>(a);
    double c, s;
    cis_large(a, c, s);
    out[6 * n + i] = c;
    out[7 * n + i] = s;
    sincos_mid(a, s, c);
    out[8 * n + i] = c;
    out[9 * n + i] = s;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int solat_regrid_bicubic(solat_ctx* c, const double* field, int n_old, int n_new, double* out)
{
    if (!c || !field || !out || n_old <= 0 || n_new <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DevBuf<double> d_in, d_tmp, d_out, d_scr;
    if (int rc = d_in.upload(field, (size_t)n_old * n_old, s)) return rc;
    if (int rc = d_tmp.reserve((size_t)n_old * n_new)) return rc;
    if (int rc = d_out.reserve((size_t)n_new * n_new)) return rc;
    if (int rc = d_scr.reserve((size_t)2 * n_old * (size_t)std::max(n_old, n_new))) return rc;
    regrid_rows_kernel<<<(n_old + 63) / 64, 64, 0, s>>>(d_in.p, n_old, n_new, d_tmp.p, d_scr.p);
    LAUNCHED();
    regrid_cols_kernel<<<(n_new + 63) / 64, 64, 0, s>>>(d_tmp.p, n_old, n_new, d_out.p, d_scr.p);
    LAUNCHED();
    CK(cudaMemcpyAsync(out, d_out.p, (size_t)n_new * n_new * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SOLAT_OK;
}

int solat_selftest_math(solat_ctx* c, const double* x, int n, double* out10)
{
    if (!c || !x || !out10 || n <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    DevBuf<double> dx, dout;
    if (int rc = dx.upload(x, n, c->stream)) return rc;
    if (int rc = dout.reserve(10 * (size_t)n)) return rc;
    selftest_math_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(dx.p, n, dout.p);
    LAUNCHED();
    CK(cudaMemcpyAsync(out10, dout.p, 10 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SOLAT_OK;
}

int solat_abi_version(void) { return SOLAT_ABI_VERSION; }
const char* solat_last_error(void) { return g_err.c_str(); }
long long solat_launch_count(void) { return g_launches.load(); }

int solat_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int solat_ctx_create(int device, solat_ctx** out)
{
    if (!out) return fail(SOLAT_ERR_ARG, "ctx out pointer is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(SOLAT_ERR_NODEV, "no CUDA device", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(SOLAT_ERR_ARG, "device index out of range");
    solat_ctx* c = new solat_ctx();
    c->device = device;
    CK(cudaSetDevice(device));
    CK(cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major != 10) {
        delete c;
        return fail(SOLAT_ERR_NODEV, "libsolat_b200 is built for sm_100a only");
    }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c;
    return SOLAT_OK;
}
void solat_ctx_destroy(solat_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaStreamDestroy(c->stream);
    delete c;
}
int solat_ctx_info(const solat_ctx* c, int* sm, int* maj, int* min_, long long* hbm)
{
    if (!c) return fail(SOLAT_ERR_ARG, "ctx is NULL");
    if (sm) *sm = c->prop.multiProcessorCount;
    if (maj) *maj = c->prop.major;
    if (min_) *min_ = c->prop.minor;
    if (hbm) *hbm = (long long)c->prop.totalGlobalMem;
    return SOLAT_OK;
}

int solat_fp64_peak(solat_ctx* c, int repeats, double* tflops)
{
    // Burst figure: best of `repeats` short (~2 ms) launches.  The map kernels are sub-millisecond bursts, so this is
    // the matching denominator; a back-to-back 75 ms run of the same kernel settles ~8 % lower on this part.
    if (!c || !tflops) return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const int blocks = c->prop.multiProcessorCount * 8, threads = 256, iters = 1024;
    DevBuf<double> out;
    if (int rc = out.reserve((size_t)blocks * threads)) return rc;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int r = 0; r < repeats + 2; r++) {
        CK(cudaEventRecord(e0, c->stream));
        fp64_peak_kernel<<<blocks, threads, 0, c->stream>>>(out.p, iters, 0.999999, 1e-9);
        LAUNCHED();
        CK(cudaEventRecord(e1, c->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 128.0 * iters * (double)blocks * threads;
        if (r >= 2) best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return SOLAT_OK;
}

// ---------------------------------------------------------------- near field rows
static int nf_rows(solat_ctx* c, int kind, const solat_nf_geom* g, const double* p, const double* theta, int nth,
                   const double* phi, int nph, double* out, int* status)
{
    if (!c || !g || !p || !theta || !phi || !out || nth <= 0 || nph < 0) return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const int rows = kind == NF_RX ? 20 : (kind == NF_SOURCE ? 26 : 32);
    const int nphi = nph == 0 ? nth : nph;                 // paired mode: phi has nth entries
    const size_t n = nph == 0 ? (size_t)nth : (size_t)nth * nph;
    AngleTabsOwner ang;
    if (int rc = ang.init(theta, nth, phi, nph, c->stream)) return rc;
    DevBuf<double> d_raw, d_out;
    DevBuf<int> d_st;
    std::vector<double> raw(nth + nphi);
    memcpy(raw.data(), theta, sizeof(double) * nth);
    memcpy(raw.data() + nth, phi, sizeof(double) * nphi);
    if (int rc = d_raw.upload(raw.data(), raw.size(), c->stream)) return rc;
    if (int rc = d_out.reserve(rows * n)) return rc;
    if (status) if (int rc = d_st.reserve(n)) return rc;
    NfRowsParams q;
    q.ang = ang.tabs; q.theta = d_raw.p; q.phi = d_raw.p + nth;
    q.p[0] = p[0]; q.p[1] = p[1]; q.p[2] = p[2];
    q.n_si = g->n_si; q.n_vac = g->n_vac; q.y_ap = g->y_ap;
    const double sig = g->th_fwhp / sqrt(8 * log(2.0));
    q.inv_2sig2 = 0.5 / (sig * sig);
    // np.mean over the raveled meshgrid == mean of the 1-D axis
    double mt = 0, mp = 0;
    for (int i = 0; i < nth; i++) mt += theta[i];
    for (int j = 0; j < nphi; j++) mp += phi[j];
    q.th_mean = mt / nth; q.ph_mean = mp / nphi;
    q.out = d_out.p; q.status = status ? d_st.p : nullptr;
    const int threads = 128, blocks = (int)((n + threads - 1) / threads);
    if (kind == NF_RX) nf_rows_kernel<NF_RX><<<blocks, threads, 0, c->stream>>>(q);
    else if (kind == NF_SOURCE) nf_rows_kernel<NF_SOURCE><<<blocks, threads, 0, c->stream>>>(q);
    else nf_rows_kernel<NF_TOY><<<blocks, threads, 0, c->stream>>>(q);
    LAUNCHED();
    CK(cudaMemcpyAsync(out, d_out.p, rows * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (status) CK(cudaMemcpyAsync(status, d_st.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SOLAT_OK;
}
int solat_nf_rx_trace(solat_ctx* c, const solat_nf_geom* g, const double p[3], const double* th, int nth,
                      const double* ph, int nph, double* out, int* st)
{
    return nf_rows(c, NF_RX, g, p, th, nth, ph, nph, out, st);
}
int solat_nf_rx_fields(solat_ctx* c, const solat_nf_geom* g, const double* p_rx, int nrx, const double* theta, int nth,
                       const double* phi, int nph, double* opl, double* amp, int* status)
{
    if (!c || !g || !p_rx || !theta || !phi || !opl || !amp || nrx <= 0 || nth <= 0 || nph <= 0 || nrx > 65535)
        return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)nth * nph;
    AngleTabsOwner ang;
    if (int rc = ang.init(theta, nth, phi, nph, c->stream)) return rc;
    DevBuf<double> d_raw, d_p, d_out;
    DevBuf<int> d_st;
    std::vector<double> raw(nth + nph);
    memcpy(raw.data(), theta, sizeof(double) * nth);
    memcpy(raw.data() + nth, phi, sizeof(double) * nph);
    if (int rc = d_raw.upload(raw.data(), raw.size(), c->stream)) return rc;
    if (int rc = d_p.upload(p_rx, 3 * (size_t)nrx, c->stream)) return rc;
    if (int rc = d_out.reserve(2 * n * nrx)) return rc;
    if (int rc = d_st.reserve(1)) return rc;
    CK(cudaMemsetAsync(d_st.p, 0, sizeof(int), c->stream));
    NfRowsParams q;
    q.ang = ang.tabs; q.theta = d_raw.p; q.phi = d_raw.p + nth;
    q.p[0] = q.p[1] = q.p[2] = 0.0;
    q.n_si = g->n_si; q.n_vac = g->n_vac; q.y_ap = g->y_ap;
    const double sig = g->th_fwhp / sqrt(8 * log(2.0));
    q.inv_2sig2 = 0.5 / (sig * sig);
    double mt = 0, mp = 0;
    for (int i = 0; i < nth; i++) mt += theta[i];
    for (int j = 0; j < nph; j++) mp += phi[j];
    q.th_mean = mt / nth; q.ph_mean = mp / nph;
    q.out = nullptr; q.status = nullptr;
    nf_rx_fields_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)nrx), 128, 0, c->stream>>>(q, d_p.p, d_out.p, d_out.p + n * nrx, d_st.p);
    LAUNCHED();
    CK(cudaMemcpyAsync(opl, d_out.p, n * nrx * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(amp, d_out.p + n * nrx, n * nrx * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    int st = 0;
    CK(cudaMemcpyAsync(&st, d_st.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (status) *status = st;
    return SOLAT_OK;
}
int solat_nf_source_trace(solat_ctx* c, const solat_nf_geom* g, const double p[3], const double* th, int nth,
                          const double* ph, int nph, double* out, int* st)
{
    return nf_rows(c, NF_SOURCE, g, p, th, nth, ph, nph, out, st);
}
int solat_nf_toy_trace(solat_ctx* c, const solat_nf_geom* g, const double p[3], const double* th, int nth,
                       const double* ph, int nph, double* out, int* st)
{
    return nf_rows(c, NF_TOY, g, p, th, nth, ph, nph, out, st);
}

}  // extern "C"

// ---------------------------------------------------------------- near field map plan
struct solat_nf_plan {
    solat_ctx* ctx;
    solat_nf_geom g;
    AngleTabsOwner ang;
    DevBuf<int> act_idx;
    DevBuf<double2> phasor;                 // [nrx][nfreq][n_act]
    DevBuf<double2> first_leg;              // [n_act][4]
    DevBuf<int> amap;                       // [3][n_act] active index of every ray's x, z and xz mirror ray
    DevBuf<FfItem> items;                   // work items of a paired map: [npix] pairs, then [npix] quads
    DevBuf<int> pair_ints;
    DevBuf<double2> phase_e;                // many-table sweeps: exp(i k L) of every (quad item, ray)
    int sym = 0;                            // bit 0 / bit 1: fan and active set are mirror symmetric in x / in z
    int pairs = 1, quads = 1;               // work items of two / four pixels allowed (SOLAT_NF_QUADS=0: pairs only)
    DevBuf<double2> partial;
    DevBuf<int> status;
    std::vector<double> kk;
    int nrx = 1;
    int n_act = 0, tiles = 1, tile_len = 0, threads = NF_T;
};

// Ray tiling of one pixel.  It depends on the ray count ONLY (never on the number of pixels), so the
// summation order of a pixel -- and therefore its bits -- is the same however the pixel list is sharded
// over calls or ranks.  Picks the tile count with the fewest idle threads in the last loop iteration; among equals the
// largest that still gives a thread 8 rays (more, smaller CTAs fill the last wave better: 51x51 px, 7328 rays,
// 2 / 3 / 5 / 6 tiles: 0.169 / 0.163 / 0.171 / 0.177 ms).
static void nf_choose_tiling(solat_nf_plan* p)
{
    const int min_tiles = std::max(1, (p->n_act + 4095) / 4096);
    double best = -1.0;
    int best_tiles = min_tiles;
    for (int tiles = min_tiles; tiles < min_tiles + 6; tiles++) {
        const int len = (p->n_act + tiles - 1) / tiles;
        const int granule = p->threads;                 // rays per CTA per loop trip
        const int iters = (len + granule - 1) / granule;
        const double thread_eff = (double)p->n_act / ((double)tiles * iters * granule);
        if (thread_eff > best + 1e-9 || (thread_eff > best - 1e-9 && iters >= 8)) { best = std::max(best, thread_eff); best_tiles = tiles; }
    }
    p->tiles = best_tiles;
    p->tile_len = (p->n_act + best_tiles - 1) / best_tiles;
}

template <int NF, int NT>
static void nf_launch_one(const NfMapParams& q, int grid, int threads, cudaStream_t s)
{
    if (q.items2) nf_map_kernel<NF, NT, 2><<<grid, threads, 0, s>>>(q);
    else if (q.items) nf_map_kernel<NF, NT, 1><<<grid, threads, 0, s>>>(q);
    else nf_map_kernel<NF, NT, 0><<<grid, threads, 0, s>>>(q);
}

// Many-table sweeps of a plan with both mirror symmetries: quad items, one trace per wavenumber, one contraction
// (nf_phase_kernel, nf_contract_kernel).  Lists above FF_PAIR_ONE_CTA pixels go through in pieces of that size and the
// ray range is cut by the ray count alone, so a pixel's bits do not depend on the list it comes in.
static int nf_map_contract(solat_nf_plan* p, const double* d_src, int npix, double* d_amp, double* d_phase, cudaStream_t s)
{
    const int nfreq = (int)p->kk.size(), nch = nfreq * p->nrx;
    int k_min = 256;
    if (const char* e = getenv("SOLAT_NF_KLEN")) k_min = std::max(16, atoi(e));   // experiments
    const int k_len = ((std::max(k_min, (p->n_act + 31) / 32) + 15) / 16) * 16;
    const int splits = (p->n_act + k_len - 1) / k_len;
    const int piece = std::min(npix, FF_PAIR_ONE_CTA);
    // rows of E per pass: 256 MB at most
    const int ldE = (p->n_act + 1) & ~1;
    const int rows_max = (int)std::max<size_t>(NF_GM, std::min<size_t>((size_t)piece, ((size_t)256 << 20) / ((size_t)ldE * sizeof(double2))));
    if (int rc = p->partial.reserve((size_t)npix * splits * nch)) return rc;
    if (int rc = p->items.reserve(2 * (size_t)piece)) return rc;
    if (int rc = p->pair_ints.reserve(8)) return rc;
    if (int rc = p->phase_e.reserve((size_t)rows_max * ldE)) return rc;
    NfMapParams q;
    q.ang = p->ang.tabs; q.src_xz = d_src; q.act_idx = p->act_idx.p; q.phasor = p->phasor.p; q.first_leg = p->first_leg.p;
    q.y_ap = p->g.y_ap; q.n_si = p->g.n_si; q.n_vac = p->g.n_vac;
    q.n_act = p->n_act; q.tiles = p->tiles; q.tile_len = p->tile_len; q.nfreq_total = nfreq; q.nch_total = nch;
    q.partial = p->partial.p; q.status = p->status.p; q.sym = p->sym; q.f0 = q.j0 = 0;
    q.items = p->items.p; q.items2 = p->items.p + piece; q.counters = p->pair_ints.p; q.counters2 = p->pair_ints.p + 2;
    // register-tiled SIMT contraction (default) or the FP64 tensor-core one (SOLAT_NF_DMMA=1).  Measured on B200, 25 tables
    // x 51 x 51 px (profiles/r02_variants.md): the two tie (0.378 vs 0.380 ms per map) -- a DMMA holds the same FP64 pipe for
    // as long as the 8 DFMAs it replaces (37.0 TFLOP/s either way), and both kernels are bound by the tail of their
    // (row block x ray split) grid, which is fixed by the ray count for the sake of list-independent bits.
    const bool dmma = getenv("SOLAT_NF_DMMA") && getenv("SOLAT_NF_DMMA")[0] == '1';
    if (dmma) {
        static std::mutex mu;
        static bool set[64] = {};
        std::lock_guard<std::mutex> lock(mu);
        if (!set[p->ctx->device & 63]) {
            CK(cudaFuncSetAttribute(nf_contract_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nf_dmma_smem()));
            set[p->ctx->device & 63] = true;
        }
    }
    NfContractParams g;
    g.E = p->phase_e.p; g.phasor = p->phasor.p; g.items = q.items; g.items2 = q.items2; g.counters2 = q.counters2;
    g.ldE = ldE; g.n_act = p->n_act; g.nrx = p->nrx; g.nfreq = nfreq; g.nch_total = nch; g.splits = splits; g.k_len = k_len;
    g.partial = p->partial.p;
    for (int pix0 = 0; pix0 < npix; pix0 += piece) {
        const int n = std::min(piece, npix - pix0);
        ff_pair_kernel<<<1, 1024, ff_pair_smem(n), s>>>(d_src + 2 * (size_t)pix0, n, ff_pair_slots(n), p->pair_ints.p, p->items.p, NF_PAIR_TOL,
                                                        p->items.p + piece);
        LAUNCHED();
        g.pix0 = pix0;
        // the number of quad items is known on the device only: passes over [row0, row0 + rows) up to the bound n, the
        // CTAs of a pass past the last item return at once
        for (int row0 = 0; row0 < n; row0 += rows_max) {
            const int rows = std::min(rows_max, n - row0);
            g.row0 = row0; g.rows = rows;
            for (int f = 0; f < nfreq; f++) {
                for (int k = 0; k < NF_MAXF; k++) q.kk[k] = k == 0 ? p->kk[f] / 1e3 : 0.0;
                nf_phase_kernel<<<rows * p->tiles, p->threads, 0, s>>>(q, row0, rows, p->phase_e.p, ldE);
                LAUNCHED();
                g.f = f;
                if (dmma) {
                    const dim3 grid((rows + DM_ROWS - 1) / DM_ROWS, (4 * p->nrx + DM_COLS - 1) / DM_COLS, splits);
                    nf_contract_dmma_kernel<<<grid, 256, nf_dmma_smem(), s>>>(g);
                } else {
                    const dim3 grid((rows + NF_GM - 1) / NF_GM, (4 * p->nrx + NF_GN - 1) / NF_GN, splits);
                    nf_contract_kernel<<<grid, 256, 0, s>>>(g);
                }
                LAUNCHED();
            }
        }
    }
    const int tot = npix * nch;
    nf_finalize_kernel<<<(tot + 127) / 128, 128, 0, s>>>(p->partial.p, npix, splits, nch, d_amp, d_phase);
    LAUNCHED();
    return SOLAT_OK;
}

static int nf_map_launch(solat_nf_plan* p, const double* d_src, int npix, double* d_amp, double* d_phase, cudaStream_t s)
{
    CK(cudaSetDevice(p->ctx->device));
    const int nfreq = (int)p->kk.size(), nch = nfreq * p->nrx;
    if (p->n_act == 0) {   // nothing passes the receiver mask: the reference returns amp 0, phase 0
        CK(cudaMemsetAsync(d_amp, 0, sizeof(double) * npix * nch, s));
        CK(cudaMemsetAsync(d_phase, 0, sizeof(double) * npix * nch, s));
        return SOLAT_OK;
    }
    nf_choose_tiling(p);
    // (the contraction traces once per wavenumber: it pays for many TABLES, 25 tables 51x51 px 2.44 -> 0.84 ms)
    if (p->sym == 3 && p->pairs && p->quads && p->nrx > NF_MAX_SUMS && !(getenv("SOLAT_NF_CONTRACT") && getenv("SOLAT_NF_CONTRACT")[0] == '0'))
        return nf_map_contract(p, d_src, npix, d_amp, d_phase, s);
    if (int rc = p->partial.reserve((size_t)npix * p->tiles * nch)) return rc;
    NfMapParams q;
    q.ang = p->ang.tabs; q.src_xz = d_src; q.act_idx = p->act_idx.p; q.phasor = p->phasor.p; q.first_leg = p->first_leg.p;
    q.y_ap = p->g.y_ap; q.n_si = p->g.n_si; q.n_vac = p->g.n_vac;
    q.n_act = p->n_act; q.tiles = p->tiles; q.tile_len = p->tile_len; q.nfreq_total = nfreq; q.nch_total = nch;
    q.partial = p->partial.p; q.status = p->status.p;
    q.items = q.items2 = nullptr; q.counters = q.counters2 = nullptr;
    q.sym = p->sym;
    // work items: pairs (+-x, z) when the plan is symmetric in x, quads (+-x, +-z) when it is symmetric in z as well
    // (with more channels than that the fused kernel is bound by the reads of the phasor tables, one per output pixel,
    // ray and table whatever the work item is: 25 tables, 51x51 px: 2.47 ms one pixel per CTA, 2.70 ms as quads;
    // sweeps over many tables take nf_map_contract above)
    const bool pairs = (p->sym & 1) && p->pairs && npix <= FF_PAIR_ONE_CTA && nch <= NF_MAX_SUMS;
    const bool quads = pairs && (p->sym & 2) && p->quads;
    if (pairs) {
        if (int rc = p->items.reserve((size_t)npix * (quads ? 2 : 1))) return rc;
        if (int rc = p->pair_ints.reserve(8)) return rc;
        ff_pair_kernel<<<1, 1024, ff_pair_smem(npix), s>>>(d_src, npix, ff_pair_slots(npix), p->pair_ints.p, p->items.p, NF_PAIR_TOL,
                                                           quads ? p->items.p + npix : nullptr);
        LAUNCHED();
        q.items = p->items.p; q.counters = p->pair_ints.p;
        if (quads) { q.items2 = p->items.p + npix; q.counters2 = p->pair_ints.p + 2; }
    }
    const int grid = npix * p->tiles;                      // pairs: an upper bound, the CTAs past the last item return
    // every launch re-traces the rays, so pack as many channels as the accumulators allow (NF * NT <= 8)
    const int n_out = quads ? 4 : (pairs ? 2 : 1);
    int max_sums = NF_MAX_SUMS;
    if (const char* e = getenv("SOLAT_NF_MAX_SUMS")) max_sums = std::max(1, atoi(e));
    const int f_step = std::max(1, std::min(NF_MAXF, max_sums / n_out));
    for (int f0 = 0; f0 < nfreq; f0 += f_step) {
        const int nf = std::min(f_step, nfreq - f0);
        q.f0 = f0;
        for (int f = 0; f < NF_MAXF; f++) q.kk[f] = f < nf ? p->kk[f0 + f] / 1e3 : 0.0;
        const int nt_cap = std::min(nf == 1 ? 8 : (nf == 2 ? 4 : 2), std::max(1, max_sums / (n_out * nf)));
        for (int j0 = 0; j0 < p->nrx;) {
            int nt = 1;
            while (nt * 2 <= nt_cap && j0 + nt * 2 <= p->nrx) nt *= 2;
            q.j0 = j0;
#define SOLAT_NF_CASE(F, T) if (nf == F && nt == T) nf_launch_one<F, T>(q, grid, p->threads, s);
            SOLAT_NF_CASE(1, 1) SOLAT_NF_CASE(1, 2) SOLAT_NF_CASE(1, 4) SOLAT_NF_CASE(1, 8)
            SOLAT_NF_CASE(2, 1) SOLAT_NF_CASE(2, 2) SOLAT_NF_CASE(2, 4)
            SOLAT_NF_CASE(3, 1) SOLAT_NF_CASE(3, 2) SOLAT_NF_CASE(4, 1) SOLAT_NF_CASE(4, 2)
#undef SOLAT_NF_CASE
            LAUNCHED();
            j0 += nt;
        }
    }
    const int tot = npix * nch;
    nf_finalize_kernel<<<(tot + 127) / 128, 128, 0, s>>>(p->partial.p, npix, p->tiles, nch, d_amp, d_phase);
    LAUNCHED();
    return SOLAT_OK;
}

extern "C" {

int solat_nf_plan_create_multi(solat_ctx* c, const solat_nf_geom* g, const double* theta, int nth, const double* phi,
                               int nph, const double* rx_opl, const double* rx_w, int nrx, const double* k_per_m,
                               int nfreq, solat_nf_plan** out)
{
    if (!c || !g || !theta || !phi || !rx_opl || !rx_w || !k_per_m || !out || nth <= 0 || nph <= 0 || nfreq <= 0 || nrx <= 0)
        return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    solat_nf_plan* p = new solat_nf_plan();
    p->ctx = c; p->g = *g; p->nrx = nrx;
    p->kk.assign(k_per_m, k_per_m + nfreq);
    int rc = p->ang.init(theta, nth, phi, nph, c->stream);
    const size_t n = (size_t)nth * nph;
    // cull: a ray is traced if any receiver table gives it weight (holo_sim.py:168-180 mask x feed taper)
    std::vector<int> idx;
    idx.reserve(n);
    for (size_t r = 0; r < n; r++) {
        bool any_w = false;
        for (int j = 0; j < nrx && !any_w; j++) any_w = rx_w[(size_t)j * n + r] != 0.0;
        if (any_w) idx.push_back((int)r);
    }
    p->n_act = (int)idx.size();
    // mirror images: x needs the fan symmetric in phi about pi/2 and the active set closed under j -> nph-1-j,
    // z the same in theta and i -> nth-1-i
    {
        const char* e = getenv("SOLAT_NF_PAIRS");
        int sym = (e && e[0] == '0') ? 0 : 3;
        e = getenv("SOLAT_NF_QUADS");
        p->quads = !(e && e[0] == '0');
        for (int j = 0; j < nph && (sym & 1); j++)
            if (!(fabs(phi[j] + phi[nph - 1 - j] - M_PI) <= 1e-12)) sym &= ~1;
        for (int i = 0; i < nth && (sym & 2); i++)
            if (!(fabs(theta[i] + theta[nth - 1 - i] - M_PI) <= 1e-12)) sym &= ~2;
        std::vector<int> am(3 * idx.size(), 0), where(sym ? n : 0, -1);      // where[ray] = its active index
        for (size_t a = 0; a < idx.size() && sym; a++) where[idx[a]] = (int)a;
        for (int m = 1; m <= 3 && sym; m++) {
            if ((m & sym) != m) continue;
            for (size_t a = 0; a < idx.size(); a++) {
                const int j = idx[a] / nth, i = idx[a] - j * nth;
                const int am_a = where[(size_t)((m & 1) ? nph - 1 - j : j) * nth + ((m & 2) ? nth - 1 - i : i)];
                if (am_a < 0) { sym &= ~(m == 3 ? 3 : m); break; }
                am[(size_t)(m - 1) * idx.size() + a] = am_a;
            }
        }
        p->sym = sym;
        if (sym && p->n_act) {
            if (!rc) rc = p->amap.upload(am.data(), am.size(), c->stream);
            if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "sync");   // `am` goes out of scope
            if (!rc && cudaFuncSetAttribute(ff_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff_pair_smem(FF_PAIR_ONE_CTA)) != cudaSuccess)
                rc = fail(SOLAT_ERR_CUDA, "cudaFuncSetAttribute");
        }
    }
    // receiver phasors w exp(i k L_rx / 1e3), one table per (receiver table, wavenumber)  (holo_sim.py:173-176):
    // the compacted (L_rx, w) pairs go up, the sincos runs on the device
    std::vector<double2> ow((size_t)nrx * p->n_act);
    for (int j = 0; j < nrx; j++)
        for (int a = 0; a < p->n_act; a++) {
            const size_t r = (size_t)j * n + idx[a];
            ow[(size_t)j * p->n_act + a] = make_double2(rx_opl[r], rx_w[r]);
        }
    DevBuf<double2> d_ow;
    DevBuf<double> d_k;
    if (!rc && p->n_act) rc = p->act_idx.upload(idx.data(), idx.size(), c->stream);
    if (!rc && p->n_act) rc = d_ow.upload(ow.data(), ow.size(), c->stream);
    if (!rc && p->n_act) rc = d_k.upload(k_per_m, nfreq, c->stream);
    DevBuf<double2> d_ph;                                   // symmetric plans: the plain table, before the mirror gather
    DevBuf<double2>& plain = p->sym ? d_ph : p->phasor;
    if (!rc && p->n_act) rc = plain.reserve((size_t)nrx * nfreq * p->n_act);
    if (!rc && p->n_act) {
        const long long total = (long long)nrx * nfreq * p->n_act;
        nf_phasor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(d_ow.p, d_k.p, nrx, nfreq, p->n_act, plain.p);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "nf_phasor_kernel");
    }
    if (!rc && p->n_act && p->sym) rc = p->phasor.reserve(4 * (size_t)nrx * nfreq * p->n_act);
    if (!rc && p->n_act && p->sym) {
        const long long per_image = (long long)nrx * nfreq * p->n_act;
        nf_phasor_mirror_kernel<<<(unsigned)((4 * per_image + 255) / 256), 256, 0, c->stream>>>(d_ph.p, p->amap.p, p->n_act, per_image, p->phasor.p);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "nf_phasor_mirror_kernel");
    }
    if (!rc && p->n_act) rc = p->first_leg.reserve(4 * (size_t)p->n_act);
    if (!rc && p->n_act) {
        nf_first_leg_kernel<<<(p->n_act + 127) / 128, 128, 0, c->stream>>>(p->ang.tabs, p->act_idx.p, p->n_act, g->y_ap, g->n_si, g->n_vac,
                                                                           p->first_leg.p);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "nf_first_leg_kernel");
    }
    if (!rc) rc = p->status.reserve(1);
    if (!rc && cudaMemsetAsync(p->status.p, 0, sizeof(int), c->stream) != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "memset");
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "sync");
    if (rc) { delete p; return rc; }
    *out = p;
    return SOLAT_OK;
}
int solat_nf_plan_create(solat_ctx* c, const solat_nf_geom* g, const double* theta, int nth, const double* phi, int nph,
                         const double* rx_opl, const double* rx_w, const double* k_per_m, int nfreq,
                         solat_nf_plan** out)
{
    return solat_nf_plan_create_multi(c, g, theta, nth, phi, nph, rx_opl, rx_w, 1, k_per_m, nfreq, out);
}
void solat_nf_plan_destroy(solat_nf_plan* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();   // device-pointer maps may still be in flight on the caller's streams
    delete p;
}
int solat_nf_plan_active_rays(const solat_nf_plan* p) { return p ? p->n_act : 0; }
int solat_nf_plan_channels(const solat_nf_plan* p) { return p ? p->nrx * (int)p->kk.size() : 0; }

int solat_nf_map_device(solat_nf_plan* p, const double* d_src, int npix, double* d_amp, double* d_phase, void* stream)
{
    if (!p || !d_src || !d_amp || !d_phase || npix <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    return nf_map_launch(p, d_src, npix, d_amp, d_phase, (cudaStream_t)stream);
}
// copy the plan's sticky status word into an 8-byte slot (int in the low word) and clear it
__global__ void move_status_kernel(int* status, int* slot) { slot[0] = *status; slot[1] = 0; *status = 0; }
int solat_nf_map_device_st(solat_nf_plan* p, const double* d_src, int npix, double* d_amp, double* d_phase, void* d_status,
                           void* stream)
{
    if (!p || !d_src || !d_amp || !d_phase || !d_status || npix <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    if (int rc = nf_map_launch(p, d_src, npix, d_amp, d_phase, (cudaStream_t)stream)) return rc;
    move_status_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p->status.p, static_cast<int*>(d_status));
    LAUNCHED();
    return SOLAT_OK;
}
int solat_nf_plan_status(solat_nf_plan* p, int* status)
{
    if (!p || !status) return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(p->ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(status, p->status.p, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemset(p->status.p, 0, sizeof(int)));
    return SOLAT_OK;
}
// copy the plan's persistent status word next to the results so that one D2H brings everything back
__global__ void pack_status_kernel(const int* status, double* slot) { *slot = (double)*status; }

int solat_nf_map_host(solat_nf_plan* p, const double* src_xz, int npix, double* amp, double* phase, int* status)
{
    if (!p || !src_xz || !amp || !phase || npix <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lock(p->ctx->stage_mu);
    CK(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    const size_t n = p->kk.size() * (size_t)p->nrx * npix;   // channels x pixels
    if (int rc = p->ctx->stage_in.upload(src_xz, 2 * (size_t)npix, s)) return rc;
    if (int rc = p->ctx->stage_out.reserve(2 * n + 1)) return rc;
    if (int rc = p->ctx->pin.reserve(2 * n + 1)) return rc;
    if (int rc = nf_map_launch(p, p->ctx->stage_in.p, npix, p->ctx->stage_out.p, p->ctx->stage_out.p + n, s)) return rc;
    pack_status_kernel<<<1, 1, 0, s>>>(p->status.p, p->ctx->stage_out.p + 2 * n);
    LAUNCHED();
    CK(cudaMemcpyAsync(p->ctx->pin.p, p->ctx->stage_out.p, (2 * n + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(amp, p->ctx->pin.p, n * sizeof(double));
    memcpy(phase, p->ctx->pin.p + n, n * sizeof(double));
    const int st = (int)p->ctx->pin.p[2 * n];
    if (st) CK(cudaMemset(p->status.p, 0, sizeof(int)));
    if (status) *status = st;
    return SOLAT_OK;
}

// ---------------------------------------------------------------- far field rows
static int ff_rows(solat_ctx* c, bool stage1, const solat_ff_geom* g, const double* theta, int nth, const double* phi,
                   int nph, const double* centre, const int* pidx1, const int* pidx2, const double* pan1, int np1,
                   const double* pan2, int np2, double el, double az, double* out, int* status)
{
    if (!c || !g || !theta || !phi || !out || nth <= 0 || nph < 0) return fail(SOLAT_ERR_ARG, "bad argument");
    if (!stage1 && (!centre || !pidx1 || !pidx2 || !pan1 || !pan2 || np1 <= 0 || np2 <= 0))
        return fail(SOLAT_ERR_ARG, "bad panel argument");
    CK(cudaSetDevice(c->device));
    const size_t n = nph == 0 ? (size_t)nth : (size_t)nth * nph;
    const int rows = stage1 ? 9 : 17;
    AngleTabsOwner ang;
    if (int rc = ang.init(theta, nth, phi, nph, c->stream)) return rc;
    DevBuf<double> d_out, d_p1, d_p2;
    DevBuf<int> d_i1, d_i2, d_st;
    if (int rc = d_out.reserve(rows * n)) return rc;
    if (status) if (int rc = d_st.reserve(n)) return rc;
    FfRowsParams q;
    memset(&q, 0, sizeof(q));
    q.ang = ang.tabs; q.g = *g; q.el = el; q.az = az;
    if (!stage1) {
        for (int i = 0; i < 3; i++) q.centre[i] = centre[i];
        if (int rc = d_i1.upload(pidx1, n, c->stream)) return rc;
        if (int rc = d_i2.upload(pidx2, n, c->stream)) return rc;
        if (int rc = d_p1.upload(pan1, 10 * (size_t)np1, c->stream)) return rc;
        if (int rc = d_p2.upload(pan2, 10 * (size_t)np2, c->stream)) return rc;
        q.pidx1 = d_i1.p; q.pidx2 = d_i2.p; q.pan1 = d_p1.p; q.pan2 = d_p2.p; q.np1 = np1; q.np2 = np2;
    }
    q.out = d_out.p; q.status = status ? d_st.p : nullptr;
    const int threads = 128, blocks = (int)((n + threads - 1) / threads);
    if (stage1) ff_rows_kernel<true><<<blocks, threads, 0, c->stream>>>(q);
    else ff_rows_kernel<false><<<blocks, threads, 0, c->stream>>>(q);
    LAUNCHED();
    CK(cudaMemcpyAsync(out, d_out.p, rows * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (status) CK(cudaMemcpyAsync(status, d_st.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SOLAT_OK;
}
int solat_ff_ray_mirror_pts(solat_ctx* c, const solat_ff_geom* g, const double* th, int nth, const double* ph, int nph,
                            double el, double az, double* out9, int* status)
{
    return ff_rows(c, true, g, th, nth, ph, nph, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, el, az, out9, status);
}
int solat_ff_apert_fields(solat_ctx* c, const solat_ff_geom* g, const double* th, int nth, const double* ph, int nph,
                          const double centre[3], const int* pidx1, const int* pidx2, const double* pan1, int np1,
                          const double* pan2, int np2, double el, double az, double* out17, int* status)
{
    return ff_rows(c, false, g, th, nth, ph, nph, centre, pidx1, pidx2, pan1, np1, pan2, np2, el, az, out17, status);
}

static FfGrid make_grid(const double* data_x, const double* data_y, int M, const double* d_gx, const double* d_gy)
{
    FfGrid G;
    G.gx = d_gx; G.gy = d_gy; G.M = M; G.gy_cs = 1; G.gy_rs = M;
    G.y_uniform = 1;
    for (int i = 0; i < M && G.y_uniform; i++)
        for (int j = 1; j < M; j++)
            if (data_y[(size_t)i * M + j] != data_y[(size_t)i * M]) { G.y_uniform = 0; break; }
    double a = data_x[0], b = data_y[0];
    for (int i = 1; i < M; i++) { a = fmax(a, data_x[i]); b = fmax(b, data_y[(size_t)i * M]); }   // ffl:19
    G.gx_max = a; G.gy_max = b;
    a = data_x[0]; b = data_y[0];
    for (int i = 1; i < M; i++) { a = fmin(a, data_x[i]); b = fmin(b, data_y[(size_t)i * M]); }
    G.gx_min = a; G.gy_min = b;
    const double sx = data_x[M - 1] - data_x[0], sy = data_y[(size_t)(M - 1) * M] - data_y[0];
    G.inv_dx = (M > 1 && sx > 0) ? (float)((M - 1) / sx) : 0.0f;   // first guess of searchsorted; the fix-up makes it exact
    G.inv_dy = (M > 1 && sy > 0) ? (float)((M - 1) / sy) : 0.0f;
    G.lean = G.y_uniform && axis_is_lean(data_x, M, 1) && axis_is_lean(data_y, M, M);
    G.dev_u = (float)fmax(axis_deviation(data_x, M, 1), axis_deviation(data_y, M, M));
    return G;
}

int solat_ff_project(solat_ctx* c, const double* sim, int n_pts, const double* data_x, const double* data_y, int M,
                     double k, double* out_reim)
{
    if (!c || !sim || !data_x || !data_y || !out_reim || n_pts <= 0 || M <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int cells = M * M;
    DevBuf<double> d_sim, d_gx, d_gy;
    DevBuf<int> d_win;
    DevBuf<double2> d_out;
    if (int rc = d_sim.upload(sim, 17 * (size_t)n_pts, s)) return rc;
    if (int rc = d_gx.upload(data_x, M, s)) return rc;
    if (int rc = d_gy.upload(data_y, cells, s)) return rc;
    if (int rc = d_win.reserve(cells)) return rc;
    if (int rc = d_out.reserve(cells)) return rc;
    CK(cudaMemsetAsync(d_win.p, 0xff, sizeof(int) * cells, s));
    const FfGrid G = make_grid(data_x, data_y, M, d_gx.p, d_gy.p);
    ff_project_scatter_kernel<<<(n_pts + 127) / 128, 128, 0, s>>>(d_sim.p, n_pts, G, d_win.p);
    LAUNCHED();
    ff_project_fill_kernel<<<(cells + 127) / 128, 128, 0, s>>>(d_sim.p, n_pts, d_win.p, cells, k, d_out.p);
    LAUNCHED();
    CK(cudaMemcpyAsync(out_reim, d_out.p, sizeof(double2) * cells, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SOLAT_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- far field map plan
struct solat_ff_plan {
    solat_ctx* ctx;
    solat_ff_geom g;
    AngleTabsOwner ang;
    DevBuf<double> gx, gy;
    DevBuf<double2> e_rec;
    DevBuf<int> status;
    DevBuf<short> pidx;                     // [2][npix][n_rays] panel columns of a panel-model map
    DevBuf<double> pan, xy;                 // panel models; mirror coordinates for identify_panel
    DevBuf<double4> lay_rect;               // device panel layouts of both mirrors (solat_ff_map_panel_layout_host)
    DevBuf<int> lay_col;
    DevBuf<unsigned short> lay_bucket;
    DevBuf<unsigned char> rec;              // per-pixel records of ff_prep_kernel (frames, ray intervals)
    DevBuf<FfPairSlot> pair_slots;          // mirror pairing of a pixel list (ff_pair_*_kernel)
    DevBuf<int> pair_ints;
    DevBuf<FfItem> items;
    int pairs;                              // geometry, fan and shared memory allow the mirror-pair walk
    FfGrid grid;
    double k, key_base, key_unit;
    int key_shift;
    int M, n_rays, rows, shape, prefilter;
    int win_global;                         // the winner tables do not fit in shared memory: global scratch
    int lean, lean_grid;                    // lean_grid: shared tables and lean E_rec axes (ff_map_kernel<0, .., LEAN> applies); lean: plain maps take it
    int interp_s;                           // > 0: large uniform fan, plain maps go through ff_map_interp_kernel with this lattice spacing
    double th0, dth, ph0, dph, interp_chk;
    DevBuf<float> probe_err;
    DevBuf<int> redo;
    DevBuf<unsigned long long> win_g;
    size_t smem;
};

constexpr int FF_NSHAPES = 3;
#ifndef SOLAT_FF_T2
#define SOLAT_FF_T2 384
#endif
#ifndef SOLAT_FF_T1
#define SOLAT_FF_T1 768
#endif
static const int FF_SHAPES[FF_NSHAPES][2] = {{FF_T, SOLAT_FF_MINB}, {SOLAT_FF_T2, 2}, {SOLAT_FF_T1, 1}};   // threads, CTAs per SM
template <int PANEL, bool LEAN = false>
static void ff_map_dispatch(int shape, const FfMapParams& q, int npix, size_t smem, cudaStream_t s)
{
    if (shape == 0) ff_map_kernel<PANEL, FF_T, SOLAT_FF_MINB, LEAN><<<npix, FF_T, smem, s>>>(q);
    else if (shape == 1) ff_map_kernel<PANEL, SOLAT_FF_T2, 2, LEAN><<<npix, SOLAT_FF_T2, smem, s>>>(q);
    else ff_map_kernel<PANEL, SOLAT_FF_T1, 1, LEAN><<<npix, SOLAT_FF_T1, smem, s>>>(q);
}
// The dynamic shared-memory limit is an attribute of the kernel, not of a plan: plans with different grids or fans
// coexist, so the limit of a shape only ever grows (per device).
static int ff_map_set_smem(int device, int shape, size_t smem)
{
    static std::mutex mu;
    static size_t granted[64][FF_NSHAPES] = {};
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = granted[device & 63][shape];
    if (smem <= have) return SOLAT_OK;
    const int b = (int)smem;
#define SOLAT_FF_SMEM_ATTR(T_, B_)                                                                                       \
    CK(cudaFuncSetAttribute(ff_map_kernel<0, T_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));                   \
    CK(cudaFuncSetAttribute(ff_map_kernel<0, T_, B_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));             \
    CK(cudaFuncSetAttribute(ff_map_kernel<1, T_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));                   \
    CK(cudaFuncSetAttribute(ff_map_kernel<2, T_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    if (shape == 0) { SOLAT_FF_SMEM_ATTR(FF_T, SOLAT_FF_MINB) }
    else if (shape == 1) { SOLAT_FF_SMEM_ATTR(SOLAT_FF_T2, 2) }
    else { SOLAT_FF_SMEM_ATTR(SOLAT_FF_T1, 1) }
#undef SOLAT_FF_SMEM_ATTR
    have = smem;
    return SOLAT_OK;
}

// prep + map over the pixels in chunks (the per-pixel records are scratch: ~1.7 kB per pixel for a 100 x 100 fan,
// ~9 kB for 1000 x 1000).  `panel`: ff_map_kernel<true> with q's panel fields set (queue walk, frames only).
// `panel`: 0 none; 1 ff_map_kernel<1> with q's per-ray panel columns (queue walk, frames only); 2 ff_map_kernel<2> with
// q's panel layouts (interval walk; the focal-position cull only when `cull`: every panel error is zero).
static int ff_map_run(solat_ff_plan* p, FfMapParams q, int npix, int panel, cudaStream_t s, bool clear_status = false, bool cull = true)
{
    const int rows = panel == 1 ? 0 : p->rows;
    const bool pairs = !panel && p->pairs;
    const size_t stride = ff_rec_stride(rows ? p->ang.tabs.nph : 0);
    // records of up to 1.5 GB per call (HBM holds 180 GB): every chunk boundary drains the GPU (tail of a wave of CTAs
    // that each run for milliseconds on a 10^6-ray fan), so a 401 x 401 map goes through in 2 chunks instead of 14
    int chunk = (int)std::min<size_t>((size_t)npix, std::max<size_t>(2048, ((size_t)1536 << 20) / stride));
    q.win_g = nullptr;
    if (p->win_global) {
        const size_t per_cta = (size_t)(pairs ? 2 : 1) * p->M * p->M;
        chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)chunk, ((size_t)1 << 27) / per_cta));   // <= 1 GiB of tables
        if (int rc = p->win_g.reserve((size_t)chunk * per_cta)) return rc;
        q.win_g = p->win_g.p;
    }
    if (int rc = p->rec.reserve((size_t)chunk * stride)) return rc;
    FfPrepParams pq;
    pq.ang = p->ang.tabs; pq.g = p->g; pq.rows = rows; pq.prefilter = (rows && cull) ? p->prefilter : 0;
    pq.grid = p->grid; pq.rec = p->rec.p; pq.stride = stride;
    pq.items = nullptr; pq.counters = nullptr; pq.item0 = 0;
    q.items = nullptr; q.counters = nullptr; q.item0 = 0;
    if (pairs) {
        // work items = canonical pixels (|az|, el) with the output slots of (+az, el) and (-az, el).  The slot table is
        // emptied by the emit phase itself.
        if (int rc = p->items.reserve((size_t)npix)) return rc;
        int* ints = nullptr;
        if (npix <= FF_PAIR_ONE_CTA) {
            if (int rc = p->pair_ints.reserve(8)) return rc;
            ints = p->pair_ints.p;
            ff_pair_kernel<<<1, 1024, ff_pair_smem(npix), s>>>(q.azel, npix, ff_pair_slots(npix), ints, p->items.p, FF_PAIR_TOL);
            LAUNCHED();
        } else {
            const unsigned nslots = 2u * (unsigned)npix + 1u;
            if (nslots > p->pair_slots.n || (size_t)npix + 4 > p->pair_ints.n) {
                if (int rc = p->pair_slots.reserve(nslots)) return rc;
                if (int rc = p->pair_ints.reserve((size_t)npix + 4)) return rc;   // [0..1] counters, [4..] the overflow list
                CK(cudaMemsetAsync(p->pair_slots.p, 0xFF, p->pair_slots.n * sizeof(FfPairSlot), s));
            }
            // the counters are zeroed per call: the one-CTA path above leaves its item count in ints[0]
            ints = p->pair_ints.p;
            CK(cudaMemsetAsync(ints, 0, 2 * sizeof(int), s));
            int* over = p->pair_ints.p + 4;
            ff_pair_insert_kernel<<<(npix + 255) / 256, 256, 0, s>>>(q.azel, npix, p->pair_slots.p, nslots, over, ints, FF_PAIR_TOL);
            LAUNCHED();
            ff_pair_attach_kernel<<<(npix + 255) / 256, 256, 0, s>>>(q.azel, npix, p->pair_slots.p, nslots, over, ints, FF_PAIR_TOL);
            LAUNCHED();
            ff_pair_emit_kernel<<<(nslots + 255) / 256, 256, 0, s>>>(q.azel, p->pair_slots.p, nslots, over, ints, nullptr, p->items.p);
            LAUNCHED();
        }
        pq.items = q.items = p->items.p; pq.counters = q.counters = ints;
    }
    const bool big_prep = p->n_rays >= FF_PILE_MIN;          // (an item lists at most n_rays rays)
    const char* pile_env = getenv("SOLAT_FF_PILE");
    pq.pile = big_prep && (pairs || !panel) && rows && pq.prefilter && !(pile_env && pile_env[0] == '0');
    const size_t prep_smem = FF_REC_FRAMES + (rows ? (2 * (size_t)FF_ROW_K * p->ang.tabs.nph + 1) * sizeof(int) : 0) +
                             (pq.pile ? 32 * (size_t)p->M * sizeof(int) : 0);
    if (prep_smem + 8192 > 48 * 1024) {        // (static shared memory counts towards the 48 kB a launch gets without opting in)
        static std::mutex mu;
        static size_t granted[64] = {};
        std::lock_guard<std::mutex> lock(mu);
        if (prep_smem > granted[p->ctx->device & 63]) {
            CK(cudaFuncSetAttribute(ff_prep_kernel<FF_PREP_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
            CK(cudaFuncSetAttribute(ff_prep_kernel<FF_PREP_T_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
            granted[p->ctx->device & 63] = prep_smem;
        }
    }
    q.rows = rows; q.rec = p->rec.p; q.rec_stride = stride;
    const bool interp = !panel && p->interp_s > 0 && rows && !q.win_g && p->lean_grid;
    q.redo = nullptr; q.probe_err = nullptr; q.interp_s = p->interp_s; q.interp_ls = 0;
    while ((1 << q.interp_ls) < p->interp_s) q.interp_ls++;
    q.th0 = p->th0; q.dth = p->dth; q.ph0 = p->ph0; q.dph = p->dph;
    q.interp_eps = FF_INTERP_EPS_CM; q.interp_chk = p->interp_chk;
    {   // cell units: u = (v - a[0]) * inv_step; 5e-4 of slack for the single-precision lattice, weights and roundings
        const double ix = (double)p->grid.inv_dx, iy = (double)p->grid.inv_dy;
        q.i_tol_x = (float)(q.interp_eps * ix) + p->grid.dev_u + 5.0e-4f; q.i_tol_y = (float)(q.interp_eps * iy) + p->grid.dev_u + 5.0e-4f;
        q.i_uhi = (float)(p->M - 1) + 0.75f; q.i_umx = (float)(-2.0 * p->grid.gx_min * ix);   // the mirror pixel's ray lands at (-x, z + dzm)
        q.i_chk_x = (float)(q.interp_chk * ix); q.i_chk_y = (float)(q.interp_chk * iy);
    }
    if (interp) if (int rc = p->redo.reserve((size_t)chunk)) return rc;
    const double* azel = q.azel;
    double *amp = q.amp, *phase = q.phase;
    const short *i1 = q.pidx1, *i2 = q.pidx2;
    for (int p0 = 0; p0 < npix; p0 += chunk) {
        const int n = std::min(chunk, npix - p0);
        pq.clear_status = (clear_status && p0 == 0) ? q.status : nullptr;
        if (pairs) {
            pq.item0 = q.item0 = p0;                        // CTAs past the last item return at once
        } else {
            pq.azel = azel + 2 * (size_t)p0;
            q.azel = azel + 2 * (size_t)p0; q.amp = amp + p0; q.phase = phase + p0;
        }
        if (big_prep) ff_prep_kernel<FF_PREP_T_BIG><<<n, FF_PREP_T_BIG, prep_smem, s>>>(pq);
        else ff_prep_kernel<FF_PREP_T><<<n, FF_PREP_T, prep_smem, s>>>(pq);
        LAUNCHED();
        if (panel == 1) { q.pidx1 = i1 + (size_t)p0 * p->n_rays; q.pidx2 = i2 + (size_t)p0 * p->n_rays; }
        if (panel == 1) ff_map_dispatch<1>(p->shape, q, n, p->smem, s);
        else if (panel == 2) ff_map_dispatch<2>(p->shape, q, n, p->smem, s);
        else if (interp) {
            // cells from interpolated positions, winners traced exactly; items whose predictions fail their probes
            // (none on the reference's geometry) are flagged and redone by the exact kernel right behind
            CK(cudaMemsetAsync(p->redo.p, 0, sizeof(int) * (size_t)n, s));
            q.redo = p->redo.p;
            const char* dbg = getenv("SOLAT_FF_INTERP_DEBUG");
            q.probe_err = nullptr;
            if (dbg && dbg[0] == '1') {
                if (int rc = p->probe_err.reserve((size_t)n)) return rc;
                CK(cudaMemsetAsync(p->probe_err.p, 0, sizeof(float) * (size_t)n, s));
                q.probe_err = p->probe_err.p;
            }
            const FfInterpLayout L = ff_interp_layout(p->M, p->ang.tabs.nth, p->ang.tabs.nph, p->interp_s, pairs ? 2 : 1);
            ff_map_interp_kernel<<<n, FF_INTERP_T, L.total, s>>>(q);
            LAUNCHED();
            ff_map_dispatch<0, true>(p->shape, q, n, p->smem, s);
            if (dbg && dbg[0] == '1') {      // how many items were redone, and how far the probes were from their predictions
                std::vector<int> h((size_t)n);
                std::vector<float> he((size_t)n);
                CK(cudaMemcpyAsync(he.data(), p->probe_err.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, s));
                CK(cudaMemcpyAsync(h.data(), p->redo.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                long bad = 0, bits[4] = {0, 0, 0, 0};
                for (int v : h) { bad += v != 0; for (int b = 0; b < 4; b++) bits[b] += (v >> b) & 1; }
                float worst = 0.0f;
                for (float v : he) worst = std::max(worst, v);
                fprintf(stderr, "[solat] ff_map_interp_kernel: %ld of %d items redone by the exact kernel (S = %d; winner invalid %ld, off in x %ld, off in y %ld, other cell %ld); "
                        "worst probe %.3g cells (allowed %.3g, ambiguity margin %.3g)\n",
                        bad, n, p->interp_s, bits[0], bits[1], bits[2], bits[3], worst, q.interp_chk * p->grid.inv_dx, q.interp_eps * p->grid.inv_dx);
            }
        }
        else if (p->lean && !q.win_g) ff_map_dispatch<0, true>(p->shape, q, n, p->smem, s);
        else ff_map_dispatch<0>(p->shape, q, n, p->smem, s);
        LAUNCHED();
    }
    return SOLAT_OK;
}

// own_status: device word the launch reports its ray status in (zeroed by the launch itself); nullptr = the plan's
// sticky status word (device-pointer entry, read with solat_ff_plan_status)
static int ff_map_launch(solat_ff_plan* p, const double* d_azel, int npix, double* d_amp, double* d_phase, cudaStream_t s,
                         int* own_status = nullptr)
{
    CK(cudaSetDevice(p->ctx->device));
    FfMapParams q;
    q.ang = p->ang.tabs; q.g = p->g; q.azel = d_azel; q.grid = p->grid; q.e_rec = p->e_rec.p;
    q.k = p->k; q.key_base = p->key_base; q.key_unit = p->key_unit; q.key_inv_unit = 1.0 / p->key_unit; q.key_shift = p->key_shift;
    q.amp = d_amp; q.phase = d_phase; q.status = own_status ? own_status : p->status.p;
    q.pidx1 = q.pidx2 = nullptr; q.pan1 = q.pan2 = nullptr; q.np1 = q.np2 = 0;
    memset(&q.lay1, 0, sizeof(q.lay1)); memset(&q.lay2, 0, sizeof(q.lay2));
    return ff_map_run(p, q, npix, 0, s, own_status != nullptr);
}

extern "C" {

int solat_ff_plan_create(solat_ctx* c, const solat_ff_geom* g, const double* theta, int nth, const double* phi, int nph,
                         const double* data_x, const double* data_y, const double* e_rec, int M, double k,
                         solat_ff_plan** out)
{
    if (!c || !g || !theta || !phi || !data_x || !data_y || !e_rec || !out || nth <= 0 || nph <= 0 || M <= 0)
        return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    int rows = ff_rows_usable(theta, nth, nph) ? 1 : 0;
    // CTA shape: the one that keeps the most warps resident next to the shared-memory winner table
    const size_t sm_smem = (size_t)c->prop.sharedMemPerMultiprocessor, cta_max = (size_t)c->prop.sharedMemPerBlockOptin;
    // mirror pairs (see FfItem): the telescope and the fan must be symmetric under x -> -x
    int pairs = rows && g->src_mm[0] == 0.0 && g->rotc_mm[0] == 0.0;
    for (int j = 0; j < nph && pairs; j++)
        if (!(fabs(phi[j] + phi[nph - 1 - j] - M_PI) <= 1e-12)) pairs = 0;
    if (const char* env = getenv("SOLAT_FF_PAIRS")) if (env[0] == '0') pairs = 0;
    int shape = -1;
    size_t smem = 0;
    // passes: pair walk (two winner tables), interval walk, queue walk (no row tables)
    for (int pass = pairs ? 0 : 1; pass < 3 && shape < 0; pass++) {
        if (pass >= 1) pairs = 0;
        if (pass == 2) rows = 0;
        int best_warps = 0;
        for (int k = 0; k < FF_NSHAPES; k++) {
            const size_t need = ff_map_smem(M, rows ? nph : 0, FF_SHAPES[k][0], pairs ? 2 : 1) + 4608;     // + static shared memory
            if (need > cta_max) continue;
            const int ctas = (int)std::min<size_t>((size_t)FF_SHAPES[k][1], sm_smem / (need + 1024));   // 1 KB reserved per CTA
            const int warps = ctas * FF_SHAPES[k][0] / 32;
            if (warps > best_warps) { best_warps = warps; shape = k; smem = need - 4608; }
        }
    }
    int win_global = 0;
    if (const char* env = getenv("SOLAT_FF_WIN_GLOBAL")) if (env[0] == '1') shape = -1;   // tests: force the global tables
    if (shape < 0) {
        // the winner table does not fit in shared memory (M > 168): keep it in global memory (L2), smallest CTA shape
        win_global = 1;
        rows = ff_rows_usable(theta, nth, nph) ? 1 : 0;
        pairs = rows && g->src_mm[0] == 0.0 && g->rotc_mm[0] == 0.0;
        for (int j = 0; j < nph && pairs; j++)
            if (!(fabs(phi[j] + phi[nph - 1 - j] - M_PI) <= 1e-12)) pairs = 0;
        if (const char* env = getenv("SOLAT_FF_PAIRS")) if (env[0] == '0') pairs = 0;
        shape = 0;
        smem = ff_map_smem(M, rows ? nph : 0, FF_SHAPES[0][0], 0);
        if (smem + 4608 > cta_max) { rows = 0; pairs = 0; smem = ff_map_smem(M, 0, FF_SHAPES[0][0], 0); }
        if (smem + 4608 > cta_max) return fail(SOLAT_ERR_ARG, "E_rec grid axes too large for shared memory");
    }
    if (const char* pad = getenv("SOLAT_FF_SMEM_PAD")) smem += (size_t)atol(pad);   // occupancy experiments only
    if ((long long)nth * nph > FF_MAX_RAYS_WIDE)
        return fail(SOLAT_ERR_ARG, "launch fan too large for the 24-bit ray index of the winner key (<= 16777215 rays)");
    if (int rc = ff_map_set_smem(c->device, shape, smem)) return rc;
    CK(cudaFuncSetAttribute(ff_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff_pair_smem(FF_PAIR_ONE_CTA)));
    solat_ff_plan* p = new solat_ff_plan();
    p->ctx = c; p->g = *g; p->M = M; p->k = k; p->smem = smem; p->n_rays = nth * nph; p->rows = rows; p->shape = shape; p->pairs = pairs; p->win_global = win_global;
    {
        const char* env = getenv("SOLAT_FF_PREFILTER");   // "0": trace every accepted ray (experiments, strict status)
        p->prefilter = !(env && env[0] == '0');
    }
    // fixed-point frame of the winner keys: every path length of this telescope lies within a few tens of metres
    // of |source - rotation centre| + 2.2e4 mm; unit = half an ulp of that, so the offsets are exact integers
    {
        const double dx = g->src_mm[0] - g->rotc_mm[0], dy = g->src_mm[1] - g->rotc_mm[1], dz = g->src_mm[2] - g->rotc_mm[2];
        p->key_base = rint(sqrt(dx * dx + dy * dy + dz * dz) + 2.2e4);
        p->key_unit = ldexp(1.0, ilogb(p->key_base) - 53);
        p->key_shift = (long long)nth * nph > FF_MAX_RAYS ? FF_KEY_SHIFT_WIDE : FF_KEY_SHIFT;
    }
    cudaStream_t s = c->stream;
    int rc = p->ang.init(theta, nth, phi, nph, s);
    if (!rc) rc = p->gx.upload(data_x, M, s);                      // data_x[0, :]
    if (!rc) rc = p->gy.upload(data_y, (size_t)M * M, s);
    if (!rc) p->grid = make_grid(data_x, data_y, M, p->gx.p, p->gy.p);
    {
        const char* env = getenv("SOLAT_FF_LEAN");        // "0": the generic binning (A/B runs, tests)
        p->lean_grid = !rc && p->grid.lean && !win_global && !(env && env[0] == '0');
        // measured (profiles/r02_variants.md): the lean binning pays on the 768 x 1 shape (config 5: -10 %), not on the
        // small items of the 256 x 3 shape (config 2: +2 %)
        p->lean = p->lean_grid && (shape == 2 || (env && env[0] == '1'));
    }
    // large fans: interpolated cells + exact winners (ff_map_interp_kernel).  Needs the lean grid, the interval walk,
    // a square fan with uniform axes, and room for the lattice next to the candidate tables.
    p->interp_s = 0;
    p->interp_chk = FF_INTERP_EPS_CM / 3;
    if (const char* env = getenv("SOLAT_FF_INTERP_CHK")) p->interp_chk = atof(env);   // tests: 0 fails every probe -> every item is redone
    if (!rc && p->lean_grid && rows && p->prefilter && nth == nph && nth >= FF_INTERP_MIN_N) {
        const char* env = getenv("SOLAT_FF_INTERP");       // "0": off; a number >= 2: that lattice spacing
        int S = 4;                                         // the largest power of two up to nth / 60: the probes were measured at 16 / 1000
        while (S < 32 && 120 * S <= nth) S *= 2;
        if (env && atoi(env) >= 2) { S = 2; while (S < atoi(env)) S *= 2; }
        const double dth = (theta[nth - 1] - theta[0]) / (nth - 1), dph = (phi[nph - 1] - phi[0]) / (nph - 1);
        bool uniform = true;
        for (int i = 0; i < nth && uniform; i++)
            uniform = fabs(theta[i] - (theta[0] + i * dth)) <= 1e-9 * fabs(dth) && fabs(phi[i] - (phi[0] + i * dph)) <= 1e-9 * fabs(dph);
        const FfInterpLayout L = ff_interp_layout(M, nth, nph, S, pairs ? 2 : 1);
        if (uniform && p->grid.dev_u < 1e-6f && !(env && env[0] == '0' && env[1] == 0) && L.total + 4608 <= cta_max && nth >= 4 * S) {
            p->interp_s = S; p->th0 = theta[0]; p->dth = dth; p->ph0 = phi[0]; p->dph = dph;
            if (cudaFuncSetAttribute(ff_map_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_max - 4608) != cudaSuccess)
                p->interp_s = 0;
        }
    }
    if (!rc) rc = p->e_rec.upload(reinterpret_cast<const double2*>(e_rec), (size_t)M * M, s);
    if (!rc) rc = p->status.reserve(1);
    if (!rc && cudaMemsetAsync(p->status.p, 0, sizeof(int), s) != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "memset");
    if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = fail(SOLAT_ERR_CUDA, "sync");
    if (rc) { delete p; return rc; }
    *out = p;
    return SOLAT_OK;
}
void solat_ff_plan_destroy(solat_ff_plan* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();   // device-pointer maps may still be in flight on the caller's streams
    delete p;
}
int solat_ff_map_device(solat_ff_plan* p, const double* d_azel, int npix, double* d_amp, double* d_phase, void* stream)
{
    if (!p || !d_azel || !d_amp || !d_phase || npix <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    return ff_map_launch(p, d_azel, npix, d_amp, d_phase, (cudaStream_t)stream);
}
int solat_ff_map_device_st(solat_ff_plan* p, const double* d_azel, int npix, double* d_amp, double* d_phase, void* d_status,
                           void* stream)
{
    if (!p || !d_azel || !d_amp || !d_phase || !d_status || npix <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    return ff_map_launch(p, d_azel, npix, d_amp, d_phase, (cudaStream_t)stream, static_cast<int*>(d_status));
}
int solat_ff_plan_status(solat_ff_plan* p, int* status)
{
    if (!p || !status) return fail(SOLAT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(p->ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(status, p->status.p, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemset(p->status.p, 0, sizeof(int)));
    return SOLAT_OK;
}
int solat_ff_map_host(solat_ff_plan* p, const double* azel, int npix, double* amp, double* phase, int* status)
{
    if (!p || !azel || !amp || !phase || npix <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lock(p->ctx->stage_mu);
    CK(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    const size_t n = (size_t)npix;
    if (int rc = p->ctx->stage_in.upload(azel, 2 * n, s)) return rc;
    if (int rc = p->ctx->stage_out.reserve(2 * n + 1)) return rc;
    if (int rc = p->ctx->pin.reserve(2 * n + 1)) return rc;
    // the status word travels in the 8-byte slot behind the results: one D2H, no extra launch
    int* d_status = reinterpret_cast<int*>(p->ctx->stage_out.p + 2 * n);
    if (int rc = ff_map_launch(p, p->ctx->stage_in.p, npix, p->ctx->stage_out.p, p->ctx->stage_out.p + n, s, d_status)) return rc;
    CK(cudaMemcpyAsync(p->ctx->pin.p, p->ctx->stage_out.p, (2 * n + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(amp, p->ctx->pin.p, n * sizeof(double));
    memcpy(phase, p->ctx->pin.p + n, n * sizeof(double));
    int st = 0;
    memcpy(&st, p->ctx->pin.p + 2 * n, sizeof(int));
    if (status) *status = st;
    return SOLAT_OK;
}

int solat_ff_map_mirror_xy_host(solat_ff_plan* p, const double* azel, int npix, double* xy, int* status)
{
    if (!p || !azel || !xy || npix <= 0) return fail(SOLAT_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lock(p->ctx->stage_mu);
    CK(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    const size_t n = (size_t)npix * 4 * p->n_rays;
    if (int rc = p->ctx->stage_in.upload(azel, 2 * (size_t)npix, s)) return rc;
    if (int rc = p->xy.reserve(n)) return rc;
    for (int p0 = 0; p0 < npix; p0 += 65535) {               // gridDim.y is limited to 65535
        const int np_ = std::min(65535, npix - p0);
        ff_mirror_xy_kernel<<<dim3((p->n_rays + 127) / 128, np_), 128, 0, s>>>(p->ang.tabs, p->g, p->ctx->stage_in.p + 2 * (size_t)p0,
                                                                              p->xy.p + (size_t)p0 * 4 * p->n_rays, p->status.p);
        LAUNCHED();
    }
    CK(cudaMemcpyAsync(xy, p->xy.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    int st = 0;
    CK(cudaMemcpyAsync(&st, p->status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (st) CK(cudaMemset(p->status.p, 0, sizeof(int)));
    if (status) *status = st;
    return SOLAT_OK;
}
int solat_ff_map_panels_host(solat_ff_plan* p, const double* azel, int npix, const short* pan_idx1, const short* pan_idx2,
                             const double* pan1, int np1, const double* pan2, int np2, double* amp, double* phase, int* status)
{
    if (!p || !azel || !pan_idx1 || !pan_idx2 || !pan1 || !pan2 || !amp || !phase || npix <= 0 || np1 <= 0 || np2 <= 0)
        return fail(SOLAT_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lock(p->ctx->stage_mu);
    CK(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    const size_t n = (size_t)npix, nr = n * p->n_rays;
    for (size_t i = 0; i < nr; i++)
        if (pan_idx1[i] >= np1 || pan_idx2[i] >= np2) return fail(SOLAT_ERR_ARG, "panel index outside its panel model");
    if (int rc = p->ctx->stage_in.upload(azel, 2 * n, s)) return rc;
    if (int rc = p->ctx->stage_out.reserve(2 * n + 1)) return rc;
    if (int rc = p->ctx->pin.reserve(2 * n + 1)) return rc;
    if (int rc = p->pidx.reserve(2 * nr)) return rc;
    if (int rc = p->pan.reserve(10 * ((size_t)np1 + np2))) return rc;
    CK(cudaMemcpyAsync(p->pidx.p, pan_idx1, nr * sizeof(short), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->pidx.p + nr, pan_idx2, nr * sizeof(short), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->pan.p, pan1, 10 * (size_t)np1 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->pan.p + 10 * (size_t)np1, pan2, 10 * (size_t)np2 * sizeof(double), cudaMemcpyHostToDevice, s));
    FfMapParams q;
    q.ang = p->ang.tabs; q.g = p->g; q.azel = p->ctx->stage_in.p; q.grid = p->grid; q.e_rec = p->e_rec.p;
    q.k = p->k; q.key_base = p->key_base; q.key_unit = p->key_unit; q.key_inv_unit = 1.0 / p->key_unit; q.key_shift = p->key_shift;
    q.amp = p->ctx->stage_out.p; q.phase = p->ctx->stage_out.p + n; q.status = p->status.p;
    q.pidx1 = p->pidx.p; q.pidx2 = p->pidx.p + nr; q.pan1 = p->pan.p; q.pan2 = p->pan.p + 10 * (size_t)np1; q.np1 = np1; q.np2 = np2;
    memset(&q.lay1, 0, sizeof(q.lay1)); memset(&q.lay2, 0, sizeof(q.lay2));
    if (int rc = ff_map_run(p, q, npix, 1, s)) return rc;
    pack_status_kernel<<<1, 1, 0, s>>>(p->status.p, p->ctx->stage_out.p + 2 * n);
    LAUNCHED();
    CK(cudaMemcpyAsync(p->ctx->pin.p, p->ctx->stage_out.p, (2 * n + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(amp, p->ctx->pin.p, n * sizeof(double));
    memcpy(phase, p->ctx->pin.p + n, n * sizeof(double));
    const int st = (int)p->ctx->pin.p[2 * n];
    if (st) CK(cudaMemset(p->status.p, 0, sizeof(int)));
    if (status) *status = st;
    return SOLAT_OK;
}


// host side of a device panel layout: bucket grid over the bounding box of the panels
struct PanelLayoutHost {
    std::vector<double4> rect;
    std::vector<int> col;
    std::vector<unsigned short> bucket;
    int linear = 0;
    float bx0 = 0, by0 = 0, inv_bw = 0, inv_bh = 0;
};
static int build_layout(const solat_panel_layout* in, int np_model, PanelLayoutHost& out)
{
    if (!in || in->n < 0 || in->n > 65000 || (in->n && (!in->rect || !in->model_col))) return fail(SOLAT_ERR_ARG, "bad panel layout");
    const int n = in->n;
    out.rect.resize(n); out.col.resize(n);
    double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    for (int k = 0; k < n; k++) {
        const double* r = in->rect + 4 * (size_t)k;
        if (!(r[0] < r[1]) || !(r[2] < r[3])) return fail(SOLAT_ERR_ARG, "panel rectangle with x0 >= x1 or y0 >= y1");
        if (in->model_col[k] >= np_model) return fail(SOLAT_ERR_ARG, "panel layout names a column outside its panel model");
        out.rect[k] = make_double4(r[0], r[1], r[2], r[3]);
        out.col[k] = in->model_col[k];
        if (k == 0) { x0 = r[0]; x1 = r[1]; y0 = r[2]; y1 = r[3]; }
        x0 = std::min(x0, r[0]); x1 = std::max(x1, r[1]); y0 = std::min(y0, r[2]); y1 = std::max(y1, r[3]);
    }
    out.bucket.assign((size_t)PANEL_G * PANEL_G * PANEL_BK, 0xffff);
    if (n == 0) { out.inv_bw = out.inv_bh = 0.0f; return SOLAT_OK; }
    // a little wider than the panels, and every panel is entered into the buckets its bounds padded by `pad` touch:
    // the bucket index is computed in single precision on the device
    const double pad = 1e-3 * std::max(x1 - x0, y1 - y0) + 1e-2;
    x0 -= 2 * pad; y0 -= 2 * pad; x1 += 2 * pad; y1 += 2 * pad;
    const double bw = (x1 - x0) / PANEL_G, bh = (y1 - y0) / PANEL_G;
    out.bx0 = (float)x0; out.by0 = (float)y0; out.inv_bw = (float)(1.0 / bw); out.inv_bh = (float)(1.0 / bh);
    std::vector<int> fill((size_t)PANEL_G * PANEL_G, 0);
    for (int k = 0; k < n && !out.linear; k++) {
        const double4 r = out.rect[k];
        const int gx0 = std::max(0, (int)floor((r.x - pad - x0) / bw)), gx1 = std::min(PANEL_G - 1, (int)floor((r.y + pad - x0) / bw));
        const int gy0 = std::max(0, (int)floor((r.z - pad - y0) / bh)), gy1 = std::min(PANEL_G - 1, (int)floor((r.w + pad - y0) / bh));
        for (int gy = gy0; gy <= gy1 && !out.linear; gy++)
            for (int gx = gx0; gx <= gx1; gx++) {
                int& f = fill[(size_t)gy * PANEL_G + gx];
                if (f >= PANEL_BK) { out.linear = 1; break; }
                out.bucket[((size_t)gy * PANEL_G + gx) * PANEL_BK + f++] = (unsigned short)k;
            }
    }
    return SOLAT_OK;
}

int solat_ff_map_panel_layout_host(solat_ff_plan* p, const double* azel, int npix, const solat_panel_layout* m1,
                                   const solat_panel_layout* m2, const double* pan1, int np1, const double* pan2, int np2,
                                   double* amp, double* phase, int* status)
{
    if (!p || !azel || !m1 || !m2 || !pan1 || !pan2 || !amp || !phase || npix <= 0 || np1 <= 0 || np2 <= 0)
        return fail(SOLAT_ERR_ARG, "bad argument");
    PanelLayoutHost h1, h2;
    if (int rc = build_layout(m1, np1, h1)) return rc;
    if (int rc = build_layout(m2, np2, h2)) return rc;
    std::lock_guard<std::mutex> lock(p->ctx->stage_mu);
    CK(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    const size_t n = (size_t)npix, n1 = h1.rect.size(), n2 = h2.rect.size(), nb = (size_t)PANEL_G * PANEL_G * PANEL_BK;
    if (int rc = p->ctx->stage_in.upload(azel, 2 * n, s)) return rc;
    if (int rc = p->ctx->stage_out.reserve(2 * n + 1)) return rc;
    if (int rc = p->ctx->pin.reserve(2 * n + 1)) return rc;
    if (int rc = p->pan.reserve(10 * ((size_t)np1 + np2))) return rc;
    if (int rc = p->lay_rect.reserve(n1 + n2 + 1)) return rc;
    if (int rc = p->lay_col.reserve(n1 + n2 + 1)) return rc;
    if (int rc = p->lay_bucket.reserve(2 * nb)) return rc;
    CK(cudaMemcpyAsync(p->pan.p, pan1, 10 * (size_t)np1 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->pan.p + 10 * (size_t)np1, pan2, 10 * (size_t)np2 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (n1) CK(cudaMemcpyAsync(p->lay_rect.p, h1.rect.data(), n1 * sizeof(double4), cudaMemcpyHostToDevice, s));
    if (n2) CK(cudaMemcpyAsync(p->lay_rect.p + n1, h2.rect.data(), n2 * sizeof(double4), cudaMemcpyHostToDevice, s));
    if (n1) CK(cudaMemcpyAsync(p->lay_col.p, h1.col.data(), n1 * sizeof(int), cudaMemcpyHostToDevice, s));
    if (n2) CK(cudaMemcpyAsync(p->lay_col.p + n1, h2.col.data(), n2 * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->lay_bucket.p, h1.bucket.data(), nb * sizeof(unsigned short), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->lay_bucket.p + nb, h2.bucket.data(), nb * sizeof(unsigned short), cudaMemcpyHostToDevice, s));
    // every panel error zero (the default of ff_sim.py:61-62): the focal-position cull stays valid
    bool zero_err = true;
    for (int k = 0; k < np1 && zero_err; k++) for (int r = 2; r <= 6; r++) zero_err = zero_err && pan1[(size_t)r * np1 + k] == 0.0;
    for (int k = 0; k < np2 && zero_err; k++) for (int r = 2; r <= 6; r++) zero_err = zero_err && pan2[(size_t)r * np2 + k] == 0.0;
    FfMapParams q;
    q.ang = p->ang.tabs; q.g = p->g; q.azel = p->ctx->stage_in.p; q.grid = p->grid; q.e_rec = p->e_rec.p;
    q.k = p->k; q.key_base = p->key_base; q.key_unit = p->key_unit; q.key_inv_unit = 1.0 / p->key_unit; q.key_shift = p->key_shift;
    q.amp = p->ctx->stage_out.p; q.phase = p->ctx->stage_out.p + n; q.status = p->status.p;
    q.pidx1 = q.pidx2 = nullptr; q.pan1 = p->pan.p; q.pan2 = p->pan.p + 10 * (size_t)np1; q.np1 = np1; q.np2 = np2;
    q.lay1 = PanelLayoutDev{p->lay_rect.p, p->lay_col.p, p->lay_bucket.p, (int)n1, h1.linear, h1.bx0, h1.by0, h1.inv_bw, h1.inv_bh};
    q.lay2 = PanelLayoutDev{p->lay_rect.p + n1, p->lay_col.p + n1, p->lay_bucket.p + nb, (int)n2, h2.linear, h2.bx0, h2.by0, h2.inv_bw, h2.inv_bh};
    if (int rc = ff_map_run(p, q, npix, 2, s, false, zero_err)) return rc;
    pack_status_kernel<<<1, 1, 0, s>>>(p->status.p, p->ctx->stage_out.p + 2 * n);
    LAUNCHED();
    CK(cudaMemcpyAsync(p->ctx->pin.p, p->ctx->stage_out.p, (2 * n + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));      // h1, h2 go out of scope
    memcpy(amp, p->ctx->pin.p, n * sizeof(double));
    memcpy(phase, p->ctx->pin.p + n, n * sizeof(double));
    const int st = (int)p->ctx->pin.p[2 * n];
    if (st) CK(cudaMemset(p->status.p, 0, sizeof(int)));
    if (status) *status = st;
    return SOLAT_OK;
}

}  // extern "C"
