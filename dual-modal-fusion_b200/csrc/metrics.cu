// K4 argmax + confusion matrix (solver/mainsolver.py:139-141; clean copy train/test.py:58-60) and
// K5 label scatter / RGB paint (solver/mainsolver.py:171-173, 186-189).
//
// K4 reads C floats + one label per sample and is HBM-read bound (C*4+4 B per sample).  Counts go
// to a per-CTA shared-memory histogram with warp-aggregated atomics (one atomic per warp when all lanes hit the same bin
// key, one atomicAdd per distinct bin per warp), then one 64-bit global atomic per non-empty bin per
// CTA.  Integer counts make the result independent of the order of the additions, so the matrix is
// bit-exact with the reference's float64 loop (exact below 2^53).
#include "common.cuh"

namespace dmf {

constexpr int kMaxClasses = 64;

__device__ __forceinline__ void hist_add_warp(unsigned int* hist, int key, bool valid) {
    // Warp-aggregated histogram update.  Label maps of real scenes are spatially coherent: most warps see ONE bin -> one atomic
    // for the whole warp.  Otherwise each lane adds its own count: shared-memory atomics on distinct bins do not serialise, and a
    // general __match_any_sync aggregation costs more than the few same-bin conflicts it removes (measured: confusion_at 37 -> 21 us on
    // uniformly random 4.2 M-pixel maps).
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const int leader = __ffs(active) - 1;
    const bool uniform = __all_sync(active, key == __shfl_sync(active, key, leader));
    if (uniform) {
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(&hist[key], __popc(active));
    } else {
        atomicAdd(&hist[key], 1u);
    }
}

// One thread per row: its C loads hit the same one or two 128-byte lines, the 32 rows of a warp are 32 * C * 4 contiguous bytes, and
// L1 serves the re-touched sectors (ncu: 85 % L1 hit rate, DRAM traffic = the algorithmic bytes).  A warp-cooperative variant that
// staged coalesced loads through shared memory was measured SLOWER (36.9 vs 28.7 us for 10^6 x 13): the kernel is bound by the
// per-row scan and the histogram, not by the loads.
template <typename TT>
__global__ void __launch_bounds__(256) argmax_confusion_kernel(const float* __restrict__ logits,
                                                               const TT* __restrict__ target, int64_t N, int C,
                                                               int64_t* __restrict__ pred_out,
                                                               unsigned long long* __restrict__ cm) {
    extern __shared__ unsigned int hist[];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // all lanes of a warp iterate together so that the warp-collectives stay converged
    const int64_t n_iter = (N + stride - 1) / stride;
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t i = it * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool valid = i < N;
        int best = 0;
        if (valid) {
            const float* row = logits + i * C;
            float bv = __ldg(row);
            for (int c = 1; c < C; ++c) {
                float v = __ldg(row + c);
                if (v > bv) { bv = v; best = c; }   // strict > keeps the first maximum (torch.max semantics)
            }
            if (pred_out) pred_out[i] = best;
        }
        int key = 0;
        bool ok = valid;
        if (valid && cm) {
            int t = (int)target[i];
            ok = t >= 0 && t < C;
            key = best * C + t;
        }
        if (cm) hist_add_warp(hist, key, ok);
    }
    __syncthreads();
    if (cm)
        for (int i = threadIdx.x; i < C * C; i += blockDim.x)
            if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

__global__ void scatter_labels_kernel(const int64_t* __restrict__ x, const int64_t* __restrict__ y,
                                      const int64_t* __restrict__ pred, int64_t N, uint8_t* __restrict__ map, int W) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        map[x[i] * W + y[i]] = (uint8_t)pred[i];
}

struct Palette { uint8_t rgb[kMaxClasses * 3]; };

// 4 pixels per thread: one 32-bit load of labels, three 32-bit stores of RGB bytes.
__global__ void __launch_bounds__(256) paint_kernel(const uint8_t* __restrict__ map, int64_t npix, Palette pal, int C,
                                                    uint8_t* __restrict__ rgb) {
    __shared__ uint8_t lut[kMaxClasses * 3];
    for (int i = threadIdx.x; i < C * 3; i += blockDim.x) lut[i] = pal.rgb[i];
    __syncthreads();
    const int64_t nq = npix / 4;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t l4 = __ldg(reinterpret_cast<const uint32_t*>(map) + q);
        uint8_t b[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int l = (l4 >> (8 * j)) & 0xff;
            l = l < C ? l : 0;
            b[3 * j] = lut[3 * l]; b[3 * j + 1] = lut[3 * l + 1]; b[3 * j + 2] = lut[3 * l + 2];
        }
        uint32_t* o = reinterpret_cast<uint32_t*>(rgb) + 3 * q;
        o[0] = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
        o[1] = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
        o[2] = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
    }
    if (blockIdx.x == 0 && threadIdx.x < (npix & 3)) {
        int64_t i = nq * 4 + threadIdx.x;
        int l = map[i] < C ? map[i] : 0;
        rgb[3 * i] = lut[3 * l]; rgb[3 * i + 1] = lut[3 * l + 1]; rgb[3 * i + 2] = lut[3 * l + 2];
    }
}

}  // namespace dmf

using namespace dmf;

// cm[pred_map[k]][label_map[k]] += 1 for the listed pixels: the confusion matrix of a loader's sample set taken from
// whole-scene maps (Solver.test() after a scene-dense pass; order of the samples is irrelevant to the counts).
__global__ void __launch_bounds__(256) confusion_at_kernel(const uint8_t* __restrict__ pred_map, const uint8_t* __restrict__ label_map,
                                                           const int64_t* __restrict__ idx, int64_t N, int C,
                                                           unsigned long long* __restrict__ cm) {
    extern __shared__ unsigned int hist[];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_iter = (N + stride - 1) / stride;
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t i = it * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        bool ok = i < N;
        int key = 0;
        if (ok) {
            const int64_t k = idx ? idx[i] : i;
            const int pr = pred_map[k], t = label_map[k];
            ok = pr < C && t < C;
            key = pr * C + t;
        }
        hist_add_warp(hist, key, ok);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
        if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

extern "C" {

int dmf_confusion_at(const uint8_t* pred_map_dev, const uint8_t* label_map_dev, const int64_t* flat_idx_dev, int64_t N, int C,
                     int64_t* cm_dev, void* stream) {
    DMF_REQUIRE(pred_map_dev && label_map_dev && cm_dev && N >= 0, "confusion_at: bad argument");
    DMF_REQUIRE(C >= 1 && C <= kMaxClasses, "confusion_at: 1 <= C <= %d", kMaxClasses);
    if (N == 0) return DMF_OK;
    const int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)num_sms() * 8);
    confusion_at_kernel<<<grid, 256, sizeof(unsigned int) * C * C, (cudaStream_t)stream>>>(pred_map_dev, label_map_dev, flat_idx_dev, N, C,
                                                                                         (unsigned long long*)cm_dev);
    DMF_LAUNCHED();
    return DMF_OK;
}

int dmf_argmax_confusion(const float* logits_dev, const void* target_dev, int target_dtype, int64_t N, int C,
                         int64_t* pred_out_dev, int64_t* cm_dev, void* stream) {
    DMF_REQUIRE(logits_dev && N >= 0 && C > 0 && C <= kMaxClasses, "argmax_confusion: bad argument (C<=%d)", kMaxClasses);
    DMF_REQUIRE(!cm_dev || target_dev, "argmax_confusion: confusion matrix needs targets");
    DMF_REQUIRE(target_dtype == DMF_F32 || target_dtype == DMF_U8, "argmax_confusion: target dtype f32 or u8");
    if (N == 0) return DMF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)num_sms() * 8);
    const size_t sm = sizeof(unsigned int) * C * C;
    if (target_dtype == DMF_F32)
        argmax_confusion_kernel<float><<<grid, 256, sm, st>>>(logits_dev, (const float*)target_dev, N, C, pred_out_dev,
                                                              (unsigned long long*)cm_dev);
    else
        argmax_confusion_kernel<uint8_t><<<grid, 256, sm, st>>>(logits_dev, (const uint8_t*)target_dev, N, C,
                                                                pred_out_dev, (unsigned long long*)cm_dev);
    DMF_LAUNCHED();
    return DMF_OK;
}

int dmf_scatter_labels(const int64_t* x_dev, const int64_t* y_dev, const int64_t* pred_dev, int64_t N,
                       uint8_t* label_map_dev, int W, void* stream) {
    DMF_REQUIRE(x_dev && y_dev && pred_dev && label_map_dev && W > 0 && N >= 0, "scatter_labels: bad argument");
    if (N == 0) return DMF_OK;
    const int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)num_sms() * 8);
    scatter_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, pred_dev, N, label_map_dev, W);
    DMF_LAUNCHED();
    return DMF_OK;
}

int dmf_paint_labels(const uint8_t* label_map_dev, int64_t npix, const uint8_t* palette_host, int C,
                     uint8_t* rgb_out_dev, void* stream) {
    DMF_REQUIRE(label_map_dev && palette_host && rgb_out_dev && C > 0 && C <= kMaxClasses && npix >= 0,
                "paint_labels: bad argument");
    DMF_REQUIRE(((uintptr_t)label_map_dev & 3) == 0 && ((uintptr_t)rgb_out_dev & 3) == 0, "paint_labels: 4-byte alignment");
    if (npix == 0) return DMF_OK;
    Palette pal;
    for (int i = 0; i < C * 3; ++i) pal.rgb[i] = palette_host[i];
    const int grid = (int)std::min<int64_t>((npix / 4 + 255) / 256 + 1, (int64_t)num_sms() * 8);
    paint_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(label_map_dev, npix, pal, C, rgb_out_dev);
    DMF_LAUNCHED();
    return DMF_OK;
}

}  // extern "C"
