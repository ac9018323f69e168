// Scene preparation (to_tensor + data_padding, function/function.py:99-124) and the K1 patch
// gather (dataset_dual / dataset_tri + collate, train/dataset.py:158-188, 248-282).
//
// HBM layout: the scene lives on the device as normalised fp32, already reflect-padded, so every
// patch is a contiguous window: MS [Hp][Wp][4] (one float4 per pixel), PAN [H4p][pitch] with the
// pitch rounded to 4 floats so that a PAN window row (4p floats starting at column 4y) is
// 16-byte aligned.  K1 is pure data movement bounded by the HBM write of the patch tensors
// (4 p^2 + 16 p^2 floats per pixel); the source windows overlap and stay L2-resident, so outputs
// are written with streaming stores.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"
#include "net_geom.cuh"

// K1 extras of a scene: the TMA tensor maps of the PAN-grid rasters.
struct dmf_scene_k1 {
    alignas(64) CUtensorMap tm_pan, tm_mspan;
    int p_maps = 0, rc = 0;                          // patch size / PAN rows per chunk the maps were encoded for (0 = none: gather_scalar_kernel)
    bool has_mspan_map = false;
};

namespace dmf {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}

// ---------------------------------------------------------------- min / max over a raster
template <typename T>
__global__ void minmax_partial_kernel(const T* __restrict__ a, int64_t n, double* __restrict__ part) {
    double lo = 1e300, hi = -1e300;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = (double)a[i];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ double slo[32], shi[32];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { slo[w] = lo; shi[w] = hi; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        lo = l < nw ? slo[l] : 1e300;
        hi = l < nw ? shi[l] : -1e300;
        for (int o = 16; o; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (l == 0) { part[2 * blockIdx.x] = lo; part[2 * blockIdx.x + 1] = hi; }
    }
}

__global__ void minmax_final_kernel(const double* __restrict__ part, int nblk, double* __restrict__ lohi) {
    double lo = 1e300, hi = -1e300;
    for (int i = threadIdx.x; i < nblk; i += 32) {
        lo = fmin(lo, part[2 * i]);
        hi = fmax(hi, part[2 * i + 1]);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (threadIdx.x == 0) { lohi[0] = lo; lohi[1] = hi; }
}

// (v - lo) / (hi - lo) with numpy's promotion rules: integer rasters subtract in the integer type
// and true-divide in float64; float32 rasters stay float32; float64 rasters stay float64.
template <typename T> struct Norm;
template <> struct Norm<uint8_t> {
    __device__ static double q(uint8_t v, double lo, double hi) {
        return __ddiv_rn((double)(uint8_t)(v - (uint8_t)lo), (double)(uint8_t)((uint8_t)hi - (uint8_t)lo));
    }
    static constexpr bool f32_native = false;
};
template <> struct Norm<uint16_t> {
    __device__ static double q(uint16_t v, double lo, double hi) {
        return __ddiv_rn((double)(uint16_t)(v - (uint16_t)lo), (double)(uint16_t)((uint16_t)hi - (uint16_t)lo));
    }
    static constexpr bool f32_native = false;
};
template <> struct Norm<float> {
    __device__ static double q(float v, double lo, double hi) {
        return (double)__fdiv_rn(__fsub_rn(v, (float)lo), __fsub_rn((float)hi, (float)lo));
    }
    static constexpr bool f32_native = true;
};
template <> struct Norm<double> {
    __device__ static double q(double v, double lo, double hi) { return __ddiv_rn(__dsub_rn(v, lo), __dsub_rn(hi, lo)); }
    static constexpr bool f32_native = false;
};

// out[r][c][b] = norm(raw[reflect(r)][reflect(c)][b]); out rows have `out_pitch` elements.
// One output row per blockIdx.y (no 64-bit div/mod per element); threads sweep the row's Wp*bands elements.
template <typename T, typename O>
__global__ void __launch_bounds__(256) normalize_pad_kernel(const T* __restrict__ raw, int H, int W, int bands, int Hp, int Wp,
                                                            int64_t out_pitch, const double* __restrict__ lohi, O* __restrict__ out) {
    const double lo = lohi[0], hi = lohi[1];
    const int row_elems = Wp * bands;
    for (int r = blockIdx.y; r < Hp; r += gridDim.y) {
        const T* src = raw + (int64_t)reflect101(r, H) * W * bands;
        O* dst = out + (int64_t)r * out_pitch;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < row_elems; e += gridDim.x * blockDim.x) {
            const int c = e / bands, b = e - c * bands;
            const T v = src[(int64_t)reflect101(c, W) * bands + b];
            dst[e] = (O)Norm<T>::q(v, lo, hi);      // double -> float is round-to-nearest-even
        }
    }
}

template <typename T>
static int minmax_t(const T* raw, int64_t n, double* lohi_dev, cudaStream_t st) {
    const int nblk = (int)std::min<int64_t>(num_sms() * 8, (n + 255) / 256);
    double* scratch = nullptr;
    DMF_CUDA(cudaMallocAsync(&scratch, sizeof(double) * 2 * nblk, st));
    minmax_partial_kernel<T><<<nblk, 256, 0, st>>>(raw, n, scratch);
    DMF_LAUNCHED();
    minmax_final_kernel<<<1, 32, 0, st>>>(scratch, nblk, lohi_dev);
    DMF_LAUNCHED();
    DMF_CUDA(cudaFreeAsync(scratch, st));
    return DMF_OK;
}

// lohi_given: device {min, max} to normalise with (a band of a larger raster), or null = this raster's own range
template <typename T>
static int normalize_pad_t(const T* raw, int H, int W, int bands, int P, void* out, int out_dtype,
                           int64_t out_pitch, const double* lohi_given, cudaStream_t st) {
    const int64_t n = (int64_t)H * W * bands;
    double* scratch = nullptr;
    DMF_CUDA(cudaMallocAsync(&scratch, sizeof(double) * 2, st));
    const double* lohi = lohi_given;
    if (!lohi) {
        DMF_TRY(minmax_t(raw, n, scratch, st));
        lohi = scratch;
    }
    const int Hp = H + P - 1, Wp = W + P - 1;
    const dim3 grid((unsigned)std::min<int64_t>(((int64_t)Wp * bands + 255) / 256, 64), (unsigned)std::min(Hp, 65535));
    if (out_dtype == DMF_F32)
        normalize_pad_kernel<T, float><<<grid, 256, 0, st>>>(raw, H, W, bands, Hp, Wp, out_pitch, lohi, (float*)out);
    else
        normalize_pad_kernel<T, double><<<grid, 256, 0, st>>>(raw, H, W, bands, Hp, Wp, out_pitch, lohi, (double*)out);
    DMF_LAUNCHED();
    DMF_CUDA(cudaFreeAsync(scratch, st));
    return DMF_OK;
}

static int normalize_pad_any(const void* raw, int dt, int H, int W, int bands, int P, void* out, int out_dtype,
                             int64_t out_pitch, cudaStream_t st, const double* lohi_given = nullptr) {
    DMF_REQUIRE(raw && out && H > 0 && W > 0 && bands > 0 && P > 0, "normalize_pad: bad shape/pointer");
    DMF_REQUIRE(out_dtype == DMF_F32 || out_dtype == DMF_F64, "normalize_pad: out_dtype must be f32/f64");
    switch (dt) {
        case DMF_U8: return normalize_pad_t((const uint8_t*)raw, H, W, bands, P, out, out_dtype, out_pitch, lohi_given, st);
        case DMF_U16: return normalize_pad_t((const uint16_t*)raw, H, W, bands, P, out, out_dtype, out_pitch, lohi_given, st);
        case DMF_F32: return normalize_pad_t((const float*)raw, H, W, bands, P, out, out_dtype, out_pitch, lohi_given, st);
        case DMF_F64: return normalize_pad_t((const double*)raw, H, W, bands, P, out, out_dtype, out_pitch, lohi_given, st);
    }
    set_error("normalize_pad: unknown dtype %d", dt);
    return DMF_ERR_ARG;
}

// ---------------------------------------------------------------- IHS product straight into the scene
// dataset_tri's third raster (train/dataset.py:249-268) is the product of IHS_tran (image_convert/IHS.py:40-54) on the normalised
// rasters, reflect-padded like PAN.  One pass: per padded output element, source element (r, c) = reflect-101 of the padded
// coordinate; the four bands of MS pixel (r / 4, c / 4) and the PAN element are normalised in float64 (to_tensor,
// function/function.py:120-124, numpy promotion rules via Norm<T>), the zero-stuffed bands `up` follow from the (m, n) offsets
// (unpooling, :22-29), then I = running band mean, delta = PAN - I, result = up + delta, MSPAN = running band mean, all float64
// in the reference's operation order, and the float32 cast of the dataset (train/dataset.py:265-268) on the way out.
template <typename TM, typename TP>
__global__ void __launch_bounds__(256) ihs_scene_kernel(const TM* __restrict__ ms, const TP* __restrict__ pan, const int8_t* __restrict__ offs,
                                                        const double* __restrict__ ms_lohi, const double* __restrict__ pan_lohi, int H, int W,
                                                        int H4p, int W4p, int pitch, float* __restrict__ out) {
    const double mlo = ms_lohi[0], mhi = ms_lohi[1], plo = pan_lohi[0], phi = pan_lohi[1];
    const int64_t HW = (int64_t)H * W;
    for (int R = blockIdx.y; R < H4p; R += gridDim.y) {
        const int r = reflect101(R, 4 * H), j = r >> 2, rr = r & 3;
        for (int Cc = blockIdx.x * blockDim.x + threadIdx.x; Cc < W4p; Cc += gridDim.x * blockDim.x) {
            const int c = Cc < 4 * W ? Cc : reflect101(Cc, 4 * W), k = c >> 2, cc = c & 3;      // the modulo only on the padded fringe
            const int64_t px = (int64_t)j * W + k;
            double up[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const char2 o = __ldg(reinterpret_cast<const char2*>(offs) + i * HW + px);
                up[i] = 0.0;
                if (o.x == rr && o.y == cc) up[i] = Norm<TM>::q(ms[px * 4 + i], mlo, mhi);      // 1 element in 16: a branch, not a select (FP64 divide)
            }
            const double pv = Norm<TP>::q(pan[(int64_t)r * (4 * W) + c], plo, phi);
            const double I = run_mean4(up[0], up[1], up[2], up[3]);
            const double delta = __dsub_rn(pv, I);
            out[(int64_t)R * pitch + Cc] =
                (float)run_mean4(__dadd_rn(up[0], delta), __dadd_rn(up[1], delta), __dadd_rn(up[2], delta), __dadd_rn(up[3], delta));
        }
    }
}

template <typename TM, typename TP>
static void launch_ihs_scene(const void* ms, const void* pan, const int8_t* offs, const double* mlohi, const double* plohi, dmf_scene* s, cudaStream_t st) {
    const dim3 grid((unsigned)std::min((s->W4p + 255) / 256, 64), (unsigned)std::min(s->H4p, 65535));
    ihs_scene_kernel<TM, TP><<<grid, 256, 0, st>>>((const TM*)ms, (const TP*)pan, offs, mlohi, plohi, s->H, s->W, s->H4p, s->W4p, s->pan_pitch, s->mspan);
}

template <typename TM>
static int ihs_scene_pan(const void* ms, const void* pan, int pan_dtype, const int8_t* offs, const double* mlohi, const double* plohi, dmf_scene* s,
                         cudaStream_t st) {
    switch (pan_dtype) {
        case DMF_U8: launch_ihs_scene<TM, uint8_t>(ms, pan, offs, mlohi, plohi, s, st); return DMF_OK;
        case DMF_U16: launch_ihs_scene<TM, uint16_t>(ms, pan, offs, mlohi, plohi, s, st); return DMF_OK;
        case DMF_F32: launch_ihs_scene<TM, float>(ms, pan, offs, mlohi, plohi, s, st); return DMF_OK;
        case DMF_F64: launch_ihs_scene<TM, double>(ms, pan, offs, mlohi, plohi, s, st); return DMF_OK;
    }
    set_error("scene_set_mspan_ihs: unknown PAN dtype %d", pan_dtype);
    return DMF_ERR_ARG;
}

// padded f32/f64 [rows][cols*bands] -> f32 [rows][pitch]
template <typename T>
__global__ void repitch_cast_kernel(const T* __restrict__ in, int rows, int64_t row_elems, int64_t out_pitch,
                                    float* __restrict__ out) {
    const int64_t total = (int64_t)rows * row_elems;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / row_elems, c = i - r * row_elems;
        out[r * out_pitch + c] = (float)in[i];
    }
}

// keep freed stream-ordered allocations cached in the device's default pool (the default threshold of 0
// hands them back to the driver at every synchronisation, which costs 100+ ms for scene-sized buffers)
static void keep_pool_memory() {
    static bool done = false;
    if (done) return;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done = true;
}

static int upload(const void* src, size_t bytes, int on_device, cudaStream_t st, void** tmp, const void** dev) {
    *tmp = nullptr;
    keep_pool_memory();
    if (on_device) { *dev = src; return DMF_OK; }
    DMF_CUDA(cudaMallocAsync(tmp, bytes, st));
    DMF_CUDA(cudaMemcpyAsync(*tmp, src, bytes, cudaMemcpyHostToDevice, st));
    *dev = *tmp;
    return DMF_OK;
}

static int scene_alloc(dmf_scene* s, int H, int W, int p) {
    s->H = H; s->W = W; s->p = p;
    s->Hp = H + p - 1; s->Wp = W + p - 1;
    s->H4p = 4 * H + 4 * p - 1; s->W4p = 4 * W + 4 * p - 1;
    s->pan_pitch = (s->W4p + 3) & ~3;
    DMF_CUDA(cudaMalloc(&s->ms, sizeof(float) * 4 * (size_t)s->Hp * s->Wp));
    DMF_CUDA(cudaMalloc(&s->pan, sizeof(float) * (size_t)s->H4p * s->pan_pitch));
    return DMF_OK;
}

// ---------------------------------------------------------------- K1 gather
// dataset_dual / dataset_tri + default_collate + .to(device) (train/dataset.py:168-185, 259-279; solver/mainsolver.py:50) is pure
// data movement: per pixel a p x p x 4 MS window (HWC -> CHW) and one or two 4p x 4p windows of the PAN grid, 20 KB (28 KB tri) of
// fp32 at p = 16, bounded by the HBM WRITE of the batch (the windows overlap and are served by L2).
//   * PAN / MSPAN windows (80 - 89 % of the bytes): no register ever touches the data.  A window (or a <= 16 KB chunk of its rows)
//     is one TMA tensor load (cp.async.bulk.tensor, 2-D box of the pitched raster; the window starts at column 4y, so the box is
//     16-byte aligned as TMA requires) into shared memory; the box lands densely packed = the output tensor's layout, and leaves
//     again as one bulk copy shared -> global (cp.async.bulk, L2 evict-first so that the batch does not push the scene out of L2).
//     One warp = one shared-memory stage driven by its elected lane: wait until the stage's previous store has read it, arm the
//     mbarrier, load, wait, store.
//   * MS windows start at an arbitrary column y (4-byte granularity: a TMA box cannot start there — measured: illegal instruction)
//     and need the HWC -> CHW transposition: kMsWarps ordinary warps per CTA, a lane takes 4 neighbouring pixels (4 x 16-byte
//     loads) and writes one float4 per band plane (the 32 lanes of a store cover 512 contiguous bytes), all loads of a patch in
//     flight before its first store.  They also write the targets.
// CTAs are persistent (2 per SM, ~96 KB of stages each).
constexpr int kK1StageMax = 16384, kMsWarps = 4;

struct GatherParams {
    const int64_t* idx;
    int64_t N, HW;
    int W, p, rc, n_chunks, upp;          // PAN rows per chunk, chunks per window, TMA units per patch
    int n_tma_warps;
    uint32_t stage_bytes;
    int dbg;                              // diagnostics (DMF_K1_DBG): 1 = no loads, 2 = no stores, 4 = stores without the L2 hint, 8 = MS only, 16 = PAN only
    const float4* ms;                     // scene MS [Hp][Wp] float4 (HWC)
    int Wp;
    float *ms_out, *pan_out, *mspan_out, *target_out;
    const uint8_t* label;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_evict_first(void* dst, uint32_t src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"((uint64_t)dst), "r"(src), "r"(bytes),
                 "l"(policy)
                 : "memory");
}

__global__ void __launch_bounds__(896) gather_tma_kernel(const __grid_constant__ CUtensorMap tm_pan, const __grid_constant__ CUtensorMap tm_mspan,
                                                         const GatherParams P) {
    extern __shared__ __align__(1024) uint8_t k1_smem[];
    __shared__ uint64_t bars[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = P.p, P4 = 4 * p;
    if (warp >= P.n_tma_warps) {
        // ------------------------------------------------ MS windows + targets: ordinary loads / stores
        if (P.dbg & 16) return;
        const int items = p * p / 4, pq = p / 4;                 // work items (4 pixels of one row) per patch
        const int ppw = items >= 32 ? 1 : 32 / items;            // patches a warp handles at once (p < 12)
        const int sub = items >= 32 ? 0 : lane / items;
        const int mw = (int)blockIdx.x * kMsWarps + (warp - P.n_tma_warps);
        const int64_t stride = (int64_t)gridDim.x * kMsWarps * ppw;
        for (int64_t n0 = (int64_t)mw * ppw; n0 < P.N; n0 += stride) {
            const int64_t n = n0 + sub;
            if (n >= P.N || sub >= ppw) continue;
            int64_t k = __ldg(P.idx + n);
            k = k < 0 ? 0 : (k >= P.HW ? P.HW - 1 : k);         // never address outside the scene
            const int x = (int)((uint32_t)k / (uint32_t)P.W), y = (int)((uint32_t)k - (uint32_t)x * (uint32_t)P.W);
            float* mo = P.ms_out + n * (int64_t)(4 * p * p);
            const int first = items >= 32 ? lane : lane - sub * items, step = items >= 32 ? 32 : items;
            for (int i0 = first; i0 < items; i0 += 2 * step) {
                float4 a[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = i0 + h * step;
                    if (i < items) {
                        const int r = i / pq, c = (i - r * pq) * 4;
                        const float4* src = P.ms + (int64_t)(x + r) * P.Wp + y + c;
                        a[h][0] = __ldg(src); a[h][1] = __ldg(src + 1); a[h][2] = __ldg(src + 2); a[h][3] = __ldg(src + 3);
                    }
                }
                if (P.dbg & 2) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = i0 + h * step;
                    if (i < items) {
                        const int r = i / pq, c = (i - r * pq) * 4;
                        float* o = mo + r * p + c;
                        __stcs(reinterpret_cast<float4*>(o), make_float4(a[h][0].x, a[h][1].x, a[h][2].x, a[h][3].x));
                        __stcs(reinterpret_cast<float4*>(o + p * p), make_float4(a[h][0].y, a[h][1].y, a[h][2].y, a[h][3].y));
                        __stcs(reinterpret_cast<float4*>(o + 2 * p * p), make_float4(a[h][0].z, a[h][1].z, a[h][2].z, a[h][3].z));
                        __stcs(reinterpret_cast<float4*>(o + 3 * p * p), make_float4(a[h][0].w, a[h][1].w, a[h][2].w, a[h][3].w));
                    }
                }
            }
            if (P.target_out && first == 0) P.target_out[n] = (float)__ldg(P.label + k);
        }
        return;
    }
    // ---------------------------------------------------- PAN / MSPAN windows: TMA load -> bulk store, one lane per warp
    if ((P.dbg & 8) || !tc::elect_one()) return;
    const uint32_t bar = tc::smem_u32(&bars[warp]);
    const uint32_t stage = tc::smem_u32(k1_smem) + (uint32_t)warp * P.stage_bytes;
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint64_t policy, keep;                  // the batch streams through L2 (evict-first); the scene windows should stay (evict-last)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    if (warp == 0) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_pan) : "memory");
    const int64_t total = P.N * P.upp, stride = (int64_t)gridDim.x * P.n_tma_warps;
    const uint32_t bytes = (uint32_t)P.rc * P4 * 4u;
    uint32_t phase = 0;
    int64_t g = (int64_t)blockIdx.x * P.n_tma_warps + warp;
    int64_t k_next = g < total ? __ldg(P.idx + g / P.upp) : 0;
    for (; g < total; g += stride) {
        const int64_t n = g / P.upp;
        const int c = (int)(g - n * P.upp);
        int64_t k = k_next;
        if (g + stride < total) k_next = __ldg(P.idx + (g + stride) / P.upp);      // in flight while this unit moves
        k = k < 0 ? 0 : (k >= P.HW ? P.HW - 1 : k);
        const int x = (int)((uint32_t)k / (uint32_t)P.W), y = (int)((uint32_t)k - (uint32_t)x * (uint32_t)P.W);
        const bool third = c >= P.n_chunks;
        const int ck = third ? c - P.n_chunks : c;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");             // the stage's previous store has read it
        if (P.dbg & 1) tc::mbar_arrive(bar);
        else {
            tc::mbar_expect_tx(bar, bytes);
            tma_load_2d(stage, third ? &tm_mspan : &tm_pan, bar, 4 * y, 4 * x + ck * P.rc, keep);
        }
        float* dst = (third ? P.mspan_out : P.pan_out) + n * (int64_t)(P4 * P4) + (int64_t)ck * P.rc * P4;
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        if (!(P.dbg & 2)) {
            if (P.dbg & 4) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((uint64_t)dst), "r"(stage), "r"(bytes) : "memory");
            else bulk_store_evict_first(dst, stage, bytes, policy);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// generic path for patch sizes that are not a multiple of 4 (TMA boxes need 16-byte rows): one CTA per patch, scalar copies
__global__ void __launch_bounds__(256) gather_scalar_kernel(dmf_scene s, const int64_t* __restrict__ idx, int64_t N,
                                                            float* __restrict__ ms_out, float* __restrict__ pan_out,
                                                            float* __restrict__ mspan_out, float* __restrict__ target_out) {
    const int64_t n = blockIdx.x;
    const int64_t HW = (int64_t)s.H * s.W;
    int64_t k = idx[n];
    k = k < 0 ? 0 : (k >= HW ? HW - 1 : k);
    const int x = (int)(k / s.W), y = (int)(k % s.W);
    const int p = s.p, P = 4 * p;
    const float* base = s.pan + (int64_t)(4 * x) * s.pan_pitch + 4 * y;
    for (int i = threadIdx.x; i < P * P; i += blockDim.x) {
        int r = i / P, c = i - r * P;
        pan_out[n * (int64_t)P * P + i] = base[(int64_t)r * s.pan_pitch + c];
        if (mspan_out)
            mspan_out[n * (int64_t)P * P + i] = s.mspan[(int64_t)(4 * x + r) * s.pan_pitch + 4 * y + c];
    }
    for (int i = threadIdx.x; i < 4 * p * p; i += blockDim.x) {
        int b = i / (p * p), rem = i - b * p * p, r = rem / p, c = rem - r * p;
        ms_out[n * (int64_t)4 * p * p + i] = s.ms[((int64_t)(x + r) * s.Wp + y + c) * 4 + b];
    }
    if (threadIdx.x == 0 && target_out) target_out[n] = (float)s.label[k];
}

static int encode_f32_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return DMF_ERR_CUDA; }
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (scene raster, rank %d) failed: CUresult %d", rank, (int)r); return DMF_ERR_CUDA; }
    return DMF_OK;
}

// (re)build the tensor maps of the PAN-grid rasters (their device buffers never move once allocated)
static int scene_k1_refresh(dmf_scene* s) {
    if (!s->k1) s->k1 = new dmf_scene_k1();
    dmf_scene_k1* k = s->k1;
    const int p = s->p;
    if (p % 4 != 0 || p > 64) { k->p_maps = 0; return DMF_OK; }   // TMA boxes: rows of <= 256 elements, MS float4 groups: gather_scalar_kernel
    const int P4 = 4 * p;
    const cuuint64_t dims[2] = {(cuuint64_t)s->W4p, (cuuint64_t)s->H4p};
    const cuuint64_t strides[1] = {(cuuint64_t)s->pan_pitch * 4};
    if (k->p_maps != p) {
        k->rc = std::min(P4, kK1StageMax / (P4 * 4));
        while (P4 % k->rc) --k->rc;                                  // chunks tile the window
        const cuuint32_t box[2] = {(cuuint32_t)P4, (cuuint32_t)k->rc};
        DMF_TRY(encode_f32_map(&k->tm_pan, s->pan, 2, dims, strides, box));
        k->tm_mspan = k->tm_pan;
        k->has_mspan_map = false;
        k->p_maps = p;
    }
    if (s->mspan && !k->has_mspan_map) {
        const cuuint32_t box[2] = {(cuuint32_t)P4, (cuuint32_t)k->rc};
        DMF_TRY(encode_f32_map(&k->tm_mspan, s->mspan, 2, dims, strides, box));
        k->has_mspan_map = true;
    }
    return DMF_OK;
}

__global__ void unpitch_kernel(const float* __restrict__ in, int rows, int cols, int pitch, float* __restrict__ out) {
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / cols, c = i - r * cols;
        out[i] = in[r * pitch + c];
    }
}

}  // namespace dmf

using namespace dmf;

extern "C" {

int dmf_abi_version(void) { return DMF_ABI_VERSION; }
const char* dmf_last_error(void) { return dmf::g_err; }
int64_t dmf_launch_count(void) { return dmf::g_launches.load(); }

int dmf_normalize_pad(const void* raw_dev, int raw_dtype, int H, int W, int bands, int P, void* out_dev,
                      int out_dtype, void* stream) {
    return normalize_pad_any(raw_dev, raw_dtype, H, W, bands, P, out_dev, out_dtype,
                             (int64_t)(W + P - 1) * bands, (cudaStream_t)stream);
}

static int scene_fill_raw(dmf_scene* s, const void* ms, int ms_dtype, const void* pan, int pan_dtype, int on_device,
                          cudaStream_t st, const double* ms_lohi = nullptr, const double* pan_lohi = nullptr) {
    const int H = s->H, W = s->W, p = s->p;
    void *t1 = nullptr, *t2 = nullptr;
    const void *dms, *dpan;
    int rc = upload(ms, dtype_size(ms_dtype) * 4 * (size_t)H * W, on_device, st, &t1, &dms);
    if (rc == DMF_OK) rc = upload(pan, dtype_size(pan_dtype) * 16 * (size_t)H * W, on_device, st, &t2, &dpan);
    if (rc == DMF_OK) rc = normalize_pad_any(dms, ms_dtype, H, W, 4, p, s->ms, DMF_F32, (int64_t)s->Wp * 4, st, ms_lohi);
    if (rc == DMF_OK) rc = normalize_pad_any(dpan, pan_dtype, 4 * H, 4 * W, 1, 4 * p, s->pan, DMF_F32, s->pan_pitch, st, pan_lohi);
    if (rc == DMF_OK) rc = scene_k1_refresh(s);
    if (t1) cudaFreeAsync(t1, st);
    if (t2) cudaFreeAsync(t2, st);
    return rc;
}

int dmf_scene_create_raw(dmf_scene** out, const void* ms, int ms_dtype, const void* pan, int pan_dtype, int H, int W,
                         int p, int on_device, void* stream) {
    DMF_REQUIRE(out && ms && pan && H > 0 && W > 0 && p > 0, "scene_create_raw: bad argument");
    dmf_scene* s = new dmf_scene();
    int rc = scene_alloc(s, H, W, p);
    if (rc == DMF_OK) rc = scene_fill_raw(s, ms, ms_dtype, pan, pan_dtype, on_device, (cudaStream_t)stream);
    if (rc != DMF_OK) { dmf_scene_destroy(s); return rc; }
    *out = s;
    return DMF_OK;
}

int dmf_scene_update_raw(dmf_scene* s, const void* ms, int ms_dtype, const void* pan, int pan_dtype, int on_device,
                         void* stream) {
    DMF_REQUIRE(s && ms && pan, "scene_update_raw: bad argument");
    return scene_fill_raw(s, ms, ms_dtype, pan, pan_dtype, on_device, (cudaStream_t)stream);
}

int dmf_raster_minmax(const void* raw_dev, int dtype, int64_t n, double* lohi_out_dev, void* stream) {
    DMF_REQUIRE(raw_dev && lohi_out_dev && n > 0, "raster_minmax: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case DMF_U8: return minmax_t((const uint8_t*)raw_dev, n, lohi_out_dev, st);
        case DMF_U16: return minmax_t((const uint16_t*)raw_dev, n, lohi_out_dev, st);
        case DMF_F32: return minmax_t((const float*)raw_dev, n, lohi_out_dev, st);
        case DMF_F64: return minmax_t((const double*)raw_dev, n, lohi_out_dev, st);
    }
    set_error("raster_minmax: unknown dtype %d", dtype);
    return DMF_ERR_ARG;
}

int dmf_scene_update_raw_range(dmf_scene* s, const void* ms, int ms_dtype, const void* pan, int pan_dtype, int on_device,
                               const double* ms_lohi_dev, const double* pan_lohi_dev, void* stream) {
    DMF_REQUIRE(s && ms && pan && ms_lohi_dev && pan_lohi_dev, "scene_update_raw_range: bad argument");
    return scene_fill_raw(s, ms, ms_dtype, pan, pan_dtype, on_device, (cudaStream_t)stream, ms_lohi_dev, pan_lohi_dev);
}

static int copy_padded(const void* src, int dtype, int rows, int64_t row_elems, int64_t pitch, float* dst,
                       int on_device, cudaStream_t st) {
    DMF_REQUIRE(dtype == DMF_F32 || dtype == DMF_F64, "padded rasters must be f32 or f64");
    void* tmp;
    const void* dev;
    DMF_TRY(upload(src, dtype_size(dtype) * (size_t)rows * row_elems, on_device, st, &tmp, &dev));
    const int64_t total = (int64_t)rows * row_elems;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 32);
    if (dtype == DMF_F32)
        repitch_cast_kernel<float><<<grid, 256, 0, st>>>((const float*)dev, rows, row_elems, pitch, dst);
    else
        repitch_cast_kernel<double><<<grid, 256, 0, st>>>((const double*)dev, rows, row_elems, pitch, dst);
    DMF_LAUNCHED();
    if (tmp) DMF_CUDA(cudaFreeAsync(tmp, st));
    return DMF_OK;
}

int dmf_scene_create_padded(dmf_scene** out, const void* ms_pad, const void* pan_pad, int dtype, int H, int W, int p,
                            int on_device, void* stream) {
    DMF_REQUIRE(out && ms_pad && pan_pad && H > 0 && W > 0 && p > 0, "scene_create_padded: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    dmf_scene* s = new dmf_scene();
    int rc = scene_alloc(s, H, W, p);
    if (rc == DMF_OK) rc = copy_padded(ms_pad, dtype, s->Hp, (int64_t)s->Wp * 4, (int64_t)s->Wp * 4, s->ms, on_device, st);
    if (rc == DMF_OK) rc = copy_padded(pan_pad, dtype, s->H4p, s->W4p, s->pan_pitch, s->pan, on_device, st);
    if (rc == DMF_OK) rc = scene_k1_refresh(s);
    if (rc != DMF_OK) { dmf_scene_destroy(s); return rc; }
    *out = s;
    return DMF_OK;
}

int dmf_scene_set_mspan(dmf_scene* s, const void* mspan_pad, int dtype, int on_device, void* stream) {
    DMF_REQUIRE(s && mspan_pad, "scene_set_mspan: null");
    if (!s->mspan) DMF_CUDA(cudaMalloc(&s->mspan, sizeof(float) * (size_t)s->H4p * s->pan_pitch));
    DMF_TRY(copy_padded(mspan_pad, dtype, s->H4p, s->W4p, s->pan_pitch, s->mspan, on_device, (cudaStream_t)stream));
    return scene_k1_refresh(s);
}

int dmf_scene_set_mspan_ihs(dmf_scene* s, const void* ms, int ms_dtype, const void* pan, int pan_dtype, int on_device,
                            const int8_t* offsets_dev, const double* ms_lohi_dev, const double* pan_lohi_dev, void* stream) {
    DMF_REQUIRE(s && ms && pan && offsets_dev, "scene_set_mspan_ihs: null");
    DMF_REQUIRE((ms_lohi_dev == nullptr) == (pan_lohi_dev == nullptr), "scene_set_mspan_ihs: give both ranges or neither");
    DMF_REQUIRE(((uintptr_t)offsets_dev & 1) == 0, "scene_set_mspan_ihs: offsets must be 2-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int H = s->H, W = s->W;
    if (!s->mspan) DMF_CUDA(cudaMalloc(&s->mspan, sizeof(float) * (size_t)s->H4p * s->pan_pitch));
    void *t1 = nullptr, *t2 = nullptr;
    double* lohi = nullptr;
    const void *dms = nullptr, *dpan = nullptr;
    int rc = upload(ms, dtype_size(ms_dtype) * 4 * (size_t)H * W, on_device, st, &t1, &dms);
    if (rc == DMF_OK) rc = upload(pan, dtype_size(pan_dtype) * 16 * (size_t)H * W, on_device, st, &t2, &dpan);
    if (rc == DMF_OK && !ms_lohi_dev) {
        if (cudaMallocAsync(&lohi, sizeof(double) * 4, st) != cudaSuccess) { set_error("scene_set_mspan_ihs: out of memory"); rc = DMF_ERR_CUDA; }
        if (rc == DMF_OK) rc = dmf_raster_minmax(dms, ms_dtype, (int64_t)H * W * 4, lohi, st);
        if (rc == DMF_OK) rc = dmf_raster_minmax(dpan, pan_dtype, (int64_t)H * W * 16, lohi + 2, st);
        ms_lohi_dev = lohi; pan_lohi_dev = lohi + 2;
    }
    if (rc == DMF_OK) {
        switch (ms_dtype) {
            case DMF_U8: rc = ihs_scene_pan<uint8_t>(dms, dpan, pan_dtype, offsets_dev, ms_lohi_dev, pan_lohi_dev, s, st); break;
            case DMF_U16: rc = ihs_scene_pan<uint16_t>(dms, dpan, pan_dtype, offsets_dev, ms_lohi_dev, pan_lohi_dev, s, st); break;
            case DMF_F32: rc = ihs_scene_pan<float>(dms, dpan, pan_dtype, offsets_dev, ms_lohi_dev, pan_lohi_dev, s, st); break;
            case DMF_F64: rc = ihs_scene_pan<double>(dms, dpan, pan_dtype, offsets_dev, ms_lohi_dev, pan_lohi_dev, s, st); break;
            default: set_error("scene_set_mspan_ihs: unknown MS dtype %d", ms_dtype); rc = DMF_ERR_ARG;
        }
        if (rc == DMF_OK) {
            g_launches.fetch_add(1, std::memory_order_relaxed);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { set_error("scene_set_mspan_ihs: kernel launch -> %s", cudaGetErrorString(e)); rc = DMF_ERR_CUDA; }
        }
    }
    if (lohi) cudaFreeAsync(lohi, st);
    if (t1) cudaFreeAsync(t1, st);
    if (t2) cudaFreeAsync(t2, st);
    if (rc == DMF_OK) rc = scene_k1_refresh(s);
    return rc;
}

int dmf_scene_set_labels(dmf_scene* s, const uint8_t* label, int on_device, void* stream) {
    DMF_REQUIRE(s && label, "scene_set_labels: null");
    if (!s->label) DMF_CUDA(cudaMalloc(&s->label, (size_t)s->H * s->W));
    DMF_CUDA(cudaMemcpyAsync(s->label, label, (size_t)s->H * s->W,
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return DMF_OK;
}

int dmf_scene_destroy(dmf_scene* s) {
    if (!s) return DMF_OK;
    cudaFree(s->ms);
    cudaFree(s->pan);
    cudaFree(s->mspan);
    cudaFree(s->label);
    delete s->k1;
    delete s;
    return DMF_OK;
}

int dmf_scene_dims(const dmf_scene* s, int32_t dims[6]) {
    DMF_REQUIRE(s && dims, "scene_dims: null");
    dims[0] = s->H; dims[1] = s->W; dims[2] = s->p; dims[3] = 4; dims[4] = s->Hp; dims[5] = s->Wp;
    return DMF_OK;
}

int dmf_scene_export(const dmf_scene* s, int which, float* out_dev, void* stream) {
    DMF_REQUIRE(s && out_dev && which >= 0 && which <= 2, "scene_export: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (which == 0) {
        DMF_CUDA(cudaMemcpyAsync(out_dev, s->ms, sizeof(float) * 4 * (size_t)s->Hp * s->Wp, cudaMemcpyDeviceToDevice, st));
        return DMF_OK;
    }
    const float* src = which == 1 ? s->pan : s->mspan;
    DMF_REQUIRE(src, "scene_export: raster %d not set", which);
    const int64_t total = (int64_t)s->H4p * s->W4p;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 32);
    unpitch_kernel<<<grid, 256, 0, st>>>(src, s->H4p, s->W4p, s->pan_pitch, out_dev);
    DMF_LAUNCHED();
    return DMF_OK;
}

int dmf_gather(const dmf_scene* s, const int64_t* flat_idx_dev, int64_t N, float* ms_out_dev, float* pan_out_dev,
               float* mspan_out_dev, float* target_out_dev, void* stream) {
    DMF_REQUIRE(s && N >= 0, "gather: bad argument");
    if (N == 0) return DMF_OK;
    DMF_REQUIRE(flat_idx_dev && ms_out_dev && pan_out_dev, "gather: null pointer");
    DMF_REQUIRE(!mspan_out_dev || s->mspan, "gather: tri mode needs dmf_scene_set_mspan");
    DMF_REQUIRE(!target_out_dev || s->label, "gather: targets need dmf_scene_set_labels");
    if (N == 0) return DMF_OK;
    DMF_REQUIRE(N < (int64_t)1 << 31, "gather: N too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const dmf_scene_k1* k = s->k1;
    static const bool force_scalar = getenv("DMF_K1_SCALAR") != nullptr;        // debugging aid: the generic kernel for every patch size
    if (k && k->p_maps == s->p && !force_scalar) {
        const int p = s->p, P4 = 4 * p;
        DMF_REQUIRE((int64_t)s->H * s->W < ((int64_t)1 << 31), "gather: scene too large for 32-bit pixel indices");
        DMF_REQUIRE(!mspan_out_dev || k->has_mspan_map, "gather: tri mode needs dmf_scene_set_mspan");
        GatherParams P{};
        P.idx = flat_idx_dev; P.N = N; P.HW = (int64_t)s->H * s->W; P.W = s->W; P.p = p; P.rc = k->rc; P.n_chunks = P4 / k->rc;
        P.upp = P.n_chunks * (mspan_out_dev ? 2 : 1);
        P.stage_bytes = ((uint32_t)(k->rc * P4 * 4) + 1023u) & ~1023u;
        static const int dbg = getenv("DMF_K1_DBG") ? atoi(getenv("DMF_K1_DBG")) : 0;
        P.dbg = dbg;
        P.ms = reinterpret_cast<const float4*>(s->ms); P.Wp = s->Wp;
        P.ms_out = ms_out_dev; P.pan_out = pan_out_dev; P.mspan_out = mspan_out_dev; P.target_out = target_out_dev; P.label = s->label;
        const int warps = (int)std::max<uint32_t>(1, std::min<uint32_t>(24, (96u * 1024u) / P.stage_bytes));
        P.n_tma_warps = warps;
        const size_t smem = (size_t)warps * P.stage_bytes;
        static bool attr_set = false;
        if (!attr_set) {
            DMF_CUDA(cudaFuncSetAttribute(gather_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr_set = true;
        }
        const int64_t units = N * P.upp;
        const int grid = (int)std::min<int64_t>((units + warps - 1) / warps, (int64_t)2 * num_sms());
        gather_tma_kernel<<<grid, (warps + kMsWarps) * 32, smem, st>>>(k->tm_pan, k->tm_mspan, P);
    } else {
        gather_scalar_kernel<<<(unsigned)N, 256, 0, st>>>(*s, flat_idx_dev, N, ms_out_dev, pan_out_dev, mspan_out_dev, target_out_dev);
    }
    DMF_LAUNCHED();
    return DMF_OK;
}

}  // extern "C"
