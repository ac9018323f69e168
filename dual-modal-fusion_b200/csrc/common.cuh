// Shared helpers for libdmf_b200: error plumbing, handle layouts, activation-layout constants.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "../../include/dmf_b200.h"

namespace dmf {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define DMF_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            dmf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return DMF_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define DMF_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            dmf::set_error(__VA_ARGS__);  \
            return DMF_ERR_ARG;           \
        }                                 \
    } while (0)

#define DMF_TRY(call)                \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != DMF_OK) return rc__; \
    } while (0)

// count + check a kernel launch
#define DMF_LAUNCHED()                                                                           \
    do {                                                                                         \
        dmf::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess) {                                                                \
            dmf::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return DMF_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline size_t dtype_size(int dt) { return dt == DMF_U8 ? 1 : dt == DMF_U16 ? 2 : dt == DMF_F32 ? 4 : 8; }

// cv2.BORDER_REFLECT_101 on an axis padded at the far end only (function/function.py:104-110)
__host__ __device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    int period = 2 * (n - 1);
    int t = i % period;
    return t < n ? t : period - t;
}

int num_sms();

#ifdef __CUDACC__
__device__ __forceinline__ double run_mean4(double a0, double a1, double a2, double a3) {
    // I = a0; I = (I*i + a_i) / (i+1) for i = 1..3   (image_convert/IHS.py:42-46, 49-53)
    // x / 2 and x / 4 are computed as x * 0.5 and x * 0.25: the same real number rounded once, hence the same double for every
    // input (IEEE 754); only the division by 3 needs a true (and, on the FP64 pipe, ~25-instruction) divide
    double I = a0;
    I = __dmul_rn(__dadd_rn(__dmul_rn(I, 1.0), a1), 0.5);
    I = __ddiv_rn(__dadd_rn(__dmul_rn(I, 2.0), a2), 3.0);
    I = __dmul_rn(__dadd_rn(__dmul_rn(I, 3.0), a3), 0.25);
    return I;
}
#endif

}  // namespace dmf

// Device scene: normalised, reflect-padded fp32 rasters.
struct dmf_scene {
    int H = 0, W = 0, p = 0;
    int Hp = 0, Wp = 0;        // padded MS grid   (H+p-1, W+p-1), 4 floats per pixel (HWC)
    int H4p = 0, W4p = 0;      // padded PAN grid  (4H+4p-1, 4W+4p-1)
    int pan_pitch = 0;         // PAN row pitch in floats (multiple of 4 -> 16-byte aligned windows)
    float* ms = nullptr;       // [Hp][Wp][4]
    float* pan = nullptr;      // [H4p][pan_pitch]
    float* mspan = nullptr;    // [H4p][pan_pitch] or null
    uint8_t* label = nullptr;  // [H][W] or null
    struct dmf_scene_k1* k1 = nullptr;   // host-side extras of the K1 gather (planar MS copy + TMA tensor maps), scene.cu
};
