// PAN stem on the tensor cores: conv3x3 1->32 (zero padding at the patch border) + BN + ReLU +
// maxpool2, straight from the fp32 PAN window of each patch (the K1 gather is fused in).
//
// A 1-channel 3x3 convolution has K = 9, far too thin for tcgen05 as a plain implicit GEMM, and the
// first CUDA-core version of this stem was the most expensive stage of the network (23 % of the
// time for 2 % of the FLOPs).  Here the operand is BUILT in shared memory:
//   * a loader warp streams the fp32 window rows of the NEXT patch into a zero-bordered,
//     double-buffered staging tile with cp.async.bulk (one 16-byte-aligned row per copy);
//   * builder warps split every tap into bf16 hi + lo (x = hi + lo to ~2^-16) and write, for every
//     output pixel, one K = 32 row
//         [ (hi_t, lo_t) t=0..8 | (hi_0,hi_1) (hi_2,hi_3) (hi_4,hi_5) (hi_6,hi_7) (hi_8,0) | 0 0 ]
//     matching weights [ (w_hi_t, w_hi_t) | w_lo pairs | 0 ], i.e. x_hi*w_hi + x_lo*w_hi + x_hi*w_lo:
//     fp32-grade products from two bf16 K=16 MMAs;
//   * rows are grouped by POOLED pixel: the four pixels of a 2x2 pooling window go to four separate
//     A matrices whose products land in four 32-column blocks of the same TMEM accumulator, so the
//     epilogue thread that owns a pooled pixel finds its 4 candidates in its own TMEM lane: BN affine,
//     max, ReLU, bf16, one 16-byte store per 8 channels — no shuffles.
// Roles (64 + 256 + 128*G threads): warp 0 = MMA issuer / TMEM owner, warp 1 = window loader,
// warps 2..9 = builders, then G epilogue groups of 4 warps.  Persistent over patches; 2-3 stage A ring
// of 32 KB.
#pragma once
#include "conv_tc.cuh"

namespace dmf {
namespace tc {

constexpr int kStemCout = 32;
constexpr int kStemMaxStages = 3;
constexpr int kStemAq = 4 * 128 * 16;            // one A matrix: 4 k-chunks x 128 rows x 16 B
constexpr int kStemStage = 4 * kStemAq;          // four window positions

struct StemPanParams {
    // source: scene windows (idx == null -> consecutive pixels from `first`) or materialised patches
    const float* scene_pan;
    int pan_pitch, scene_W;
    const int64_t* idx;
    int64_t first;
    const float* patches;
    int from_scene;
    int p;                         // MS patch size; PAN window is 4p x 4p, pooled map 2p x 2p
    int S_l2;                      // log2(2p)
    int tpp_l2;                    // log2(tiles per patch) = log2(4 p^2 / 128)
    int n_stage;                   // A ring depth (2 or 3)
    int raw_pitch;                 // floats per staging row: 4 (left pad, keeps rows 16-byte aligned) + 4p + 4
    int64_t N;
    const __nv_bfloat16* w;        // packed [4 k-chunks][32 co][8]
    const float* scale;
    const float* shift;
    __nv_bfloat16* out;            // [N][4][2p][2p][8]
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

template <int G>
__global__ void __launch_bounds__(320 + 128 * G, 1) stem_pan_tc_kernel(const StemPanParams P) {
    constexpr int kBuilders = 256;
    constexpr uint32_t TMEM_USED = G * 128;
    constexpr uint32_t TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int PW = 4 * P.p, S = 2 * P.p, RP = P.raw_pitch;
    const int raw_floats = (PW + 2) * RP;                                 // one staging buffer
    uint8_t* a_s = smem;                                                  // n_stage x kStemStage
    uint8_t* w_s = a_s + P.n_stage * kStemStage;                          // 2 KB
    float* scale_s = reinterpret_cast<float*>(w_s + 4 * kStemCout * 16);
    float* shift_s = scale_s + kStemCout;
    uint64_t* bars = reinterpret_cast<uint64_t*>(shift_s + kStemCout);    // full[3] empty[3] tfull[4] tempty[4] rawfull[2] rawempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    float* raw = reinterpret_cast<float*>(tmem_slot + 4);                 // 2 x (PW+2) x RP fp32, zero border

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (3 + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (6 + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (10 + a); };
    auto rawfull_bar = [&](int b) { return bar0 + 8u * (14 + b); };
    auto rawempty_bar = [&](int b) { return bar0 + 8u * (16 + b); };

    for (int i = threadIdx.x; i < 4 * kStemCout * 4; i += blockDim.x)    // 2 KB of weights
        reinterpret_cast<uint32_t*>(w_s)[i] = reinterpret_cast<const uint32_t*>(P.w)[i];
    for (int i = threadIdx.x; i < 2 * raw_floats; i += blockDim.x) raw[i] = 0.f;   // borders stay zero for good
    if (threadIdx.x < kStemCout) {
        scale_s[threadIdx.x] = P.scale[threadIdx.x];
        shift_s[threadIdx.x] = P.shift[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stage; ++s) { mbar_init(full_bar(s), kBuilders); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < G; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(rawfull_bar(b), 1); mbar_init(rawempty_bar(b), kBuilders); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes above vs async-proxy users below
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tpp = 1 << P.tpp_l2;
    const int64_t n_patches = (P.N - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;    // patches of this CTA
    const int64_t n_local = n_patches << P.tpp_l2;                                        // tiles of this CTA

    if (warp == 0) {
        // ------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_bf16(128, kStemCout);
        const bool leader = elect_one();
        const uint64_t w_desc0 = umma_desc(smem_u32(w_s), kStemCout * 16, 128);
        int st = 0;
        uint32_t ph = 0;
        for (int64_t i = 0; i < n_local; ++i) {
            const int acc = (int)(i % G);
            mbar_wait(tempty_bar(acc), (uint32_t)((i / G) & 1) ^ 1);
            mbar_wait(full_bar(st), ph);
            tc_fence_after();
            const uint64_t a_desc0 = umma_desc(smem_u32(a_s) + (uint32_t)st * kStemStage, 128 * 16, 128);
            const uint32_t d0 = tmem_base + (uint32_t)(acc * 128);
            if (leader) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        umma_bf16(d0 + q * kStemCout, a_desc0 + (uint64_t)((q * kStemAq + 2 * j * 128 * 16) >> 4),
                                  w_desc0 + (uint64_t)((2 * j * kStemCout * 16) >> 4), idesc, j ? 1u : 0u);
                umma_commit(empty_bar(st));
                umma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++st == P.n_stage) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ window loader: one bulk copy per window row
        for (int64_t pl = 0; pl < n_patches; ++pl) {
            const int64_t n = blockIdx.x + pl * gridDim.x;
            const int b = (int)(pl & 1);
            const float* src;
            int pitch;
            if (P.from_scene) {
                const int64_t k = P.idx ? P.idx[n] : P.first + n;
                const int x = (int)(k / P.scene_W), y = (int)(k % P.scene_W);
                src = P.scene_pan + (int64_t)(4 * x) * P.pan_pitch + 4 * y;
                pitch = P.pan_pitch;
            } else {
                src = P.patches + n * PW * PW;
                pitch = PW;
            }
            mbar_wait(rawempty_bar(b), (uint32_t)((pl >> 1) & 1) ^ 1);
            if (lane == 0) mbar_expect_tx(rawfull_bar(b), (uint32_t)(PW * PW * 4));
            __syncwarp();
            float* dst = raw + b * raw_floats + RP + 4;            // row 1, column 4
            for (int r = lane; r < PW; r += 32)
                bulk_load(smem_u32(dst + r * RP), src + (int64_t)r * pitch, (uint32_t)(PW * 4), rawfull_bar(b));
        }
    } else if (warp <= 9) {
        // ------------------------------------------------ builders: hi/lo split + im2col rows
        const int bt = threadIdx.x - 64;                       // 0..255
        const int m = bt & 127, qh = bt >> 7;                  // this thread builds q = qh and q = qh + 2
        int st = 0;
        uint32_t ph = 1;
        for (int64_t pl = 0; pl < n_patches; ++pl) {
            const int b = (int)(pl & 1);
            const float* win = raw + b * raw_floats;
            mbar_wait(rawfull_bar(b), (uint32_t)((pl >> 1) & 1));
            for (int t = 0; t < tpp; ++t) {
                mbar_wait(empty_bar(st), ph);
                const int pp = t * 128 + m;                     // pooled pixel of this row
                const int prow = pp >> P.S_l2, pcol = pp & (S - 1);
                uint8_t* stage = a_s + (size_t)st * kStemStage;
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int q = qh + 2 * it;                  // window position: qy = q >> 1, qx = q & 1
                    // pixel (y, x) = (2 prow + qy, 2 pcol + qx); tap (dy, dx) sits at staging row y + dy, column x + dx + 3
                    const float* wp = win + (2 * prow + (q >> 1)) * RP + 2 * pcol + (q & 1) + 3;
                    uint32_t W9[9];
#pragma unroll
                    for (int tp = 0; tp < 9; ++tp) {
                        const float v = wp[(tp / 3) * RP + tp % 3];
                        const float hi = __bfloat162float(__float2bfloat16_rn(v));
                        W9[tp] = pack_bf16x2(hi, v - hi);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(stage + q * kStemAq) + m;
                    dst[0] = make_uint4(W9[0], W9[1], W9[2], W9[3]);
                    dst[128] = make_uint4(W9[4], W9[5], W9[6], W9[7]);
                    dst[256] = make_uint4(W9[8], __byte_perm(W9[0], W9[1], 0x5410), __byte_perm(W9[2], W9[3], 0x5410),
                                          __byte_perm(W9[4], W9[5], 0x5410));
                    dst[384] = make_uint4(__byte_perm(W9[6], W9[7], 0x5410), W9[8] & 0xffffu, 0u, 0u);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(full_bar(st));
                if (++st == P.n_stage) { st = 0; ph ^= 1; }
            }
            mbar_arrive(rawempty_bar(b));                       // this thread no longer reads staging buffer b
        }
    } else {
        // ------------------------------------------------ epilogue (G groups of 4 warps)
        const int eg = (warp - 10) >> 2;
        const int q4 = warp & 3;
        const int m = q4 * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(eg * 128);
        for (int64_t i = eg; i < n_local; i += G) {
            const int64_t n = blockIdx.x + (i >> P.tpp_l2) * gridDim.x;
            const int pp = (int)(i & (tpp - 1)) * 128 + m;
            __nv_bfloat16* const obase = P.out + ((n * 4) * (int64_t)(S * S) + pp) * 8;
            mbar_wait(tfull_bar(eg), (uint32_t)((i / G) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v0[8], v1[8], v2[8], v3[8];
                tmem_ld8(t_row + ch * 8, v0);
                tmem_ld8(t_row + 32 + ch * 8, v1);
                tmem_ld8(t_row + 64 + ch * 8, v2);
                tmem_ld8(t_row + 96 + ch * 8, v3);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float4 sc0 = *reinterpret_cast<const float4*>(scale_s + ch * 8), sc1 = *reinterpret_cast<const float4*>(scale_s + ch * 8 + 4);
                const float4 sh0 = *reinterpret_cast<const float4*>(shift_s + ch * 8), sh1 = *reinterpret_cast<const float4*>(shift_s + ch * 8 + 4);
                const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
                float r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float a = fmaxf(fmaf(__uint_as_float(v0[k]), scv[k], shv[k]), fmaf(__uint_as_float(v1[k]), scv[k], shv[k]));
                    const float c = fmaxf(fmaf(__uint_as_float(v2[k]), scv[k], shv[k]), fmaf(__uint_as_float(v3[k]), scv[k], shv[k]));
                    r[k] = fmaxf(fmaxf(a, c), 0.f);
                }
                *reinterpret_cast<uint4*>(obase + (int64_t)ch * (S * S) * 8) =
                    make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(eg));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace dmf
