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

#include "common.cuh"

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
// One CTA per patch.  PAN window: 4p rows of p float4.  MS window: p x p pixels, each a float4 of
// 4 bands (HWC) that has to land in 4 CHW planes -> a thread takes 4 neighbouring pixels and
// writes one float4 per plane.
template <bool VEC>
__global__ void __launch_bounds__(256) gather_kernel(dmf_scene s, const int64_t* __restrict__ idx, int64_t N,
                                                     float* __restrict__ ms_out, float* __restrict__ pan_out,
                                                     float* __restrict__ mspan_out, float* __restrict__ target_out) {
    const int64_t n = blockIdx.x;
    const int64_t k = idx[n];
    const int x = (int)(k / s.W), y = (int)(k % s.W);
    const int p = s.p, P = 4 * p;
    if (VEC) {
        const int nq = P * p;   // float4 per PAN window
        const float* base = s.pan + (int64_t)(4 * x) * s.pan_pitch + 4 * y;
        float4* dst = reinterpret_cast<float4*>(pan_out + n * (int64_t)P * P);
        // 4 independent 16-byte loads in flight per thread before the first store (latency-bound otherwise)
        for (int i0 = threadIdx.x; i0 < nq; i0 += 4 * blockDim.x) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < nq) {
                    const int r = i / p, c4 = i - r * p;
                    v[u] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)r * s.pan_pitch) + c4);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < nq) __stcs(dst + i, v[u]);
            }
        }
        if (mspan_out) {
            const float* b2 = s.mspan + (int64_t)(4 * x) * s.pan_pitch + 4 * y;
            float4* d2 = reinterpret_cast<float4*>(mspan_out + n * (int64_t)P * P);
            for (int i = threadIdx.x; i < nq; i += blockDim.x) {
                int r = i / p, c4 = i - r * p;
                float4 v = __ldg(reinterpret_cast<const float4*>(b2 + (int64_t)r * s.pan_pitch) + c4);
                __stcs(d2 + i, v);
            }
        }
        const int pq = p / 4;
        const float4* ms4 = reinterpret_cast<const float4*>(s.ms);
        float* mo = ms_out + n * (int64_t)4 * p * p;
        for (int i = threadIdx.x; i < p * pq; i += blockDim.x) {
            int r = i / pq, c = (i - r * pq) * 4;
            const float4* src = ms4 + (int64_t)(x + r) * s.Wp + y + c;
            float4 a0 = __ldg(src), a1 = __ldg(src + 1), a2 = __ldg(src + 2), a3 = __ldg(src + 3);
            int o = r * p + c;
            __stcs(reinterpret_cast<float4*>(mo + o), make_float4(a0.x, a1.x, a2.x, a3.x));
            __stcs(reinterpret_cast<float4*>(mo + p * p + o), make_float4(a0.y, a1.y, a2.y, a3.y));
            __stcs(reinterpret_cast<float4*>(mo + 2 * p * p + o), make_float4(a0.z, a1.z, a2.z, a3.z));
            __stcs(reinterpret_cast<float4*>(mo + 3 * p * p + o), make_float4(a0.w, a1.w, a2.w, a3.w));
        }
    } else {
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
    }
    if (threadIdx.x == 0 && target_out) target_out[n] = (float)s.label[k];
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
    if (rc != DMF_OK) { dmf_scene_destroy(s); return rc; }
    *out = s;
    return DMF_OK;
}

int dmf_scene_set_mspan(dmf_scene* s, const void* mspan_pad, int dtype, int on_device, void* stream) {
    DMF_REQUIRE(s && mspan_pad, "scene_set_mspan: null");
    if (!s->mspan) DMF_CUDA(cudaMalloc(&s->mspan, sizeof(float) * (size_t)s->H4p * s->pan_pitch));
    return copy_padded(mspan_pad, dtype, s->H4p, s->W4p, s->pan_pitch, s->mspan, on_device, (cudaStream_t)stream);
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
    if (s->p % 4 == 0)
        gather_kernel<true><<<(unsigned)N, 256, 0, st>>>(*s, flat_idx_dev, N, ms_out_dev, pan_out_dev, mspan_out_dev, target_out_dev);
    else
        gather_kernel<false><<<(unsigned)N, 256, 0, st>>>(*s, flat_idx_dev, N, ms_out_dev, pan_out_dev, mspan_out_dev, target_out_dev);
    DMF_LAUNCHED();
    return DMF_OK;
}

}  // extern "C"
