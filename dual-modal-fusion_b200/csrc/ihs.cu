// K2: IHS_tran and pan2ms (image_convert/IHS.py:6-54) in float64 with the reference's operation
// order (explicit round-to-nearest intrinsics: no FMA contraction), so results are bit-exact.
//
// IHS_tran per MS pixel: 32 B of MS + 8 B of offsets + 128 B of PAN read, 128 B written
// (296 B / MS pixel) -> one pass over HBM instead of the reference's ~5 full-size
// float64 temporaries.  One thread produces 4 horizontally adjacent outputs (one sub-row of the
// pixel's 4x4 block): 2 x 16-byte loads of PAN, 2 x 16-byte stores.
#include <type_traits>

#include "common.cuh"

namespace dmf {

__global__ void __launch_bounds__(256) ihs_tran_kernel(const double* __restrict__ ms, const double* __restrict__ pan,
                                                       const int8_t* __restrict__ offs, double* __restrict__ out,
                                                       int H, int W) {
    // work item = (j, r, k): MS row j, sub-row r in 0..3 (blockIdx.y = 4j + r), MS col k fastest -> coalesced 32-byte pieces;
    // no 64-bit division per element
    const int64_t HW = (int64_t)H * W;
    for (int jr = blockIdx.y; jr < 4 * H; jr += gridDim.y)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < W; k += gridDim.x * blockDim.x) {
        const int r = jr & 3;
        const int j = jr >> 2;
        const int64_t px = (int64_t)j * W + k;
        const double2* m2 = reinterpret_cast<const double2*>(ms + px * 4);
        const double2 m01 = __ldg(m2), m23 = __ldg(m2 + 1);
        const double v[4] = {m01.x, m01.y, m23.x, m23.y};
        int mrow[4], ncol[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const char2 o = __ldg(reinterpret_cast<const char2*>(offs) + i * HW + px);
            mrow[i] = o.x; ncol[i] = o.y;
        }
        const int64_t orow = (int64_t)(4 * j + r) * (4 * (int64_t)W) + 4 * k;
        const double2* p2 = reinterpret_cast<const double2*>(pan + orow);
        const double2 p01 = __ldg(p2), p23 = __ldg(p2 + 1);
        const double pv[4] = {p01.x, p01.y, p23.x, p23.y};
        double res[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double up[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) up[i] = (mrow[i] == r && ncol[i] == c) ? v[i] : 0.0;
            const double I = run_mean4(up[0], up[1], up[2], up[3]);
            const double delta = __dsub_rn(pv[c], I);
            res[c] = run_mean4(__dadd_rn(up[0], delta), __dadd_rn(up[1], delta), __dadd_rn(up[2], delta),
                               __dadd_rn(up[3], delta));
        }
        double2* o2 = reinterpret_cast<double2*>(out + orow);
        o2[0] = make_double2(res[0], res[1]);
        o2[1] = make_double2(res[2], res[3]);
    }
}

// block mean of a 2x2 tile, np.mean's order ((a00 + a01) + a10) + a11, accumulated in float32 for
// float32 rasters and float64 otherwise (numpy's mean), then / 4.
template <typename T>
__device__ __forceinline__ double mean2x2(const T* __restrict__ p, int64_t pitch) {
    if constexpr (sizeof(T) == 4 && !std::is_integral<T>::value) {
        float s = __fadd_rn(__fadd_rn(__fadd_rn((float)p[0], (float)p[1]), (float)p[pitch]), (float)p[pitch + 1]);
        return (double)__fdiv_rn(s, 4.0f);
    } else {
        double s = __dadd_rn(__dadd_rn(__dadd_rn((double)p[0], (double)p[1]), (double)p[pitch]), (double)p[pitch + 1]);
        return __ddiv_rn(s, 4.0);
    }
}

// out[h][w][i] = P[2h + i%2][2w + i/2] with P = 2x2 block mean of pan  (image_convert/IHS.py:14-19)
template <typename T>
__global__ void __launch_bounds__(256) pan2ms_kernel(const T* __restrict__ pan, int H, int W, int64_t W4,
                                                     double* __restrict__ out) {
    const int64_t total = (int64_t)H * W;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(t % W);
        const int h = (int)(t / W);
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pr = 2 * h + (i & 1), pc = 2 * w + (i >> 1);
            r[i] = mean2x2(pan + (int64_t)(2 * pr) * W4 + 2 * pc, W4);
        }
        double2* o = reinterpret_cast<double2*>(out + t * 4);
        o[0] = make_double2(r[0], r[1]);
        o[1] = make_double2(r[2], r[3]);
    }
}

}  // namespace dmf

using namespace dmf;

extern "C" {

int dmf_ihs_tran(const double* ms_dev, const double* pan_dev, const int8_t* offsets_dev, double* mspan_out_dev, int H,
                 int W, void* stream) {
    DMF_REQUIRE(ms_dev && pan_dev && offsets_dev && mspan_out_dev && H > 0 && W > 0, "ihs_tran: bad argument");
    DMF_REQUIRE(((uintptr_t)ms_dev & 15) == 0 && ((uintptr_t)pan_dev & 15) == 0 && ((uintptr_t)mspan_out_dev & 15) == 0 &&
                    ((uintptr_t)offsets_dev & 1) == 0,
                "ihs_tran: pointers must be 16-byte aligned");
    const dim3 grid((unsigned)std::min((W + 255) / 256, 64), (unsigned)std::min(4 * H, 65535));
    ihs_tran_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ms_dev, pan_dev, offsets_dev, mspan_out_dev, H, W);
    DMF_LAUNCHED();
    return DMF_OK;
}

int dmf_pan2ms(const void* pan_dev, int pan_dtype, int H4, int W4, double* out_dev, void* stream) {
    DMF_REQUIRE(pan_dev && out_dev && H4 > 0 && W4 > 0 && H4 % 4 == 0 && W4 % 4 == 0,
                "pan2ms: PAN dims must be positive multiples of 4");
    const int H = H4 / 4, W = W4 / 4;
    const int64_t total = (int64_t)H * W;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (pan_dtype) {
        case DMF_U8: pan2ms_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)pan_dev, H, W, W4, out_dev); break;
        case DMF_U16: pan2ms_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)pan_dev, H, W, W4, out_dev); break;
        case DMF_F32: pan2ms_kernel<float><<<grid, 256, 0, st>>>((const float*)pan_dev, H, W, W4, out_dev); break;
        case DMF_F64: pan2ms_kernel<double><<<grid, 256, 0, st>>>((const double*)pan_dev, H, W, W4, out_dev); break;
        default: set_error("pan2ms: unknown dtype %d", pan_dtype); return DMF_ERR_ARG;
    }
    DMF_LAUNCHED();
    return DMF_OK;
}

}  // extern "C"
