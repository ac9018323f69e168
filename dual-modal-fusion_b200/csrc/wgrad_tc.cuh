// Weight gradients of the dense conv layers on the tensor cores (training step, SURVEY.md 8 row a10:
// loss.backward() of solver/mainsolver.py:54 for the convolutions).
//
//   dW[co][ci][tap] = sum over (patch, h, w) of dZ[co][h][w] * A[ci][h + dy][w + dx]
//
// is a GEMM whose contraction runs over PIXELS.  Both tensors live in the C8-planar layout
// [N][C/8][S][S][8] bf16 — 16 bytes of channels per pixel, pixels consecutive — which is exactly the UMMA
// MN-major no-swizzle operand layout: a core matrix is 8 K-rows (pixels) x 16 bytes (8 channels), the
// next 8 pixels are one image row further (LBO), the next 8 channels one plane further (SBO).  So the
// SAME two TMA boxes the forward/dgrad kernels use (the halo tile of the layer input and the dense tile
// of dZ) are consumed directly with the transpose bits of the instruction descriptor set; a 3x3 tap is
// again only a different start address into the halo tile.  No transposed copies of anything are made.
//
// One tcgen05.mma = D_tap[128 co, CI] += dZ_tile[16 px, 128 co]^T * A_tap[16 px, CI] (M = 128, N = CI, K = 16 pixels =
// two rows of 8).  The accumulators of all taps stay in TMEM for the whole kernel (persistent over this
// CTA's pixel tiles, TAPS/ROLES x CI columns); one epilogue at the end adds them with atomics to an fp32 scratch
// tensor [tap][ci][co] (see below).  When 9 x CI columns do not fit in TMEM the taps are split
// by kernel row over ROLES = 3 CTAs that walk the same tiles.  Layers with 64 output channels still issue
// M = 128: rows 64..127 read whatever follows the dZ tile in shared memory and are never looked at.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = final epilogue.
#pragma once
#include "conv_tc.cuh"

namespace dmf {
namespace tc {

struct WgradParams {
    int n_tiles;        // pixel tiles (same tiling as the forward layer)
    int tpg_l2, tiles_x_l2, NP_l2, PX_l2;
    int n_stage;
    int swap_lbo_sbo;   // diagnostics: exchange the two descriptor strides
    float* dw;          // fp32 scratch [TAPS][CI][CO], accumulated (co contiguous: one 128-byte line per warp atomic)
};

// instruction descriptor: bf16 x bf16 -> f32, A and B both MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
    return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

// (Measured on B200: with the atomics going straight to PyTorch's [co][ci][kh][kw] layout every lane of a warp hit
// a different cache line and the epilogue cost 2-10x the MMAs; hence the co-contiguous scratch + one small
// re-layout kernel, wgrad_finish_kernel in train.cu, which also folds the hi/lo halves of the MS stem.)
template <int CO, int CI, int TAPS, int ROLES, int NP>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap dz_map,
                                                          const __grid_constant__ CUtensorMap a_map, const WgradParams P) {
    constexpr int MCH = CO / 8, KCH = CI / 8;
    constexpr int TPR = TAPS / ROLES;                           // taps per role
    constexpr int TH = TAPS == 9 ? 16 / NP : 0;
    constexpr uint32_t DZ_PLANE = 128u * 16u;
    constexpr uint32_t DZ_TILE = MCH * DZ_PLANE;
    constexpr uint32_t A_PLANE = TAPS == 9 ? (uint32_t)(TH + 2) * NP * kPitch * 16 : 128u * 16u;
    constexpr uint32_t A_TILE = KCH * A_PLANE;
    constexpr uint32_t STAGE = DZ_TILE + A_TILE;
    constexpr uint32_t LBO_A = TAPS == 9 ? (uint32_t)kPitch * 16 : 128u;     // next 8 pixels of the contraction
    constexpr uint32_t TMEM_USED = TPR * CI;
    constexpr uint32_t TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static_assert(TMEM_USED <= 512 && CI % 16 == 0 && (CO == 64 || CO == 128) && TAPS % ROLES == 0, "wgrad configuration");
    static_assert(STAGE % 128 == 0 && DZ_TILE % 128 == 0, "TMA destinations must stay 128-byte aligned");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* st_s = smem;                                   // n_stage x [dZ tile | A halo tile], + one dZ tile of slack
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.n_stage * STAGE + DZ_TILE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t done_bar = bar0 + 8u * 16;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stage; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // CTA -> (role = kernel row of taps, slot among the CTAs of that role)
    const int per_role = (int)gridDim.x / ROLES;
    const int role = (int)blockIdx.x % ROLES, slot = (int)blockIdx.x / ROLES;
    const int n_local = slot < per_role ? (P.n_tiles - slot + per_role - 1) / per_role : 0;

    if (warp == 0) {
        const bool leader = elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&dz_map) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&a_map) : "memory");
        }
        int st = 0;
        uint32_t ph = 1;
        for (int i = 0; i < n_local; ++i) {
            const int tile = slot + i * per_role;
            mbar_wait(empty_bar(st), ph);
            if (leader) {
                mbar_expect_tx(full_bar(st), STAGE);
                const int grp = tile >> P.tpg_l2, t = tile & ((1 << P.tpg_l2) - 1);
                const uint32_t dst = smem_u32(st_s) + (uint32_t)st * STAGE;
                if (TAPS == 9) {
                    const int ty = t >> P.tiles_x_l2, tx = t & ((1 << P.tiles_x_l2) - 1);
                    tma_load_4d(dst, &dz_map, full_bar(st), tx * 64, grp * NP, ty * TH, 0);
                    tma_load_4d(dst + DZ_TILE, &a_map, full_bar(st), (tx * 8 - 1) * 8, grp * NP, ty * TH - 1, 0);
                } else {
                    tma_load_4d(dst, &dz_map, full_bar(st), 0, (t << P.PX_l2) >> 5, grp << P.NP_l2, 0);
                    tma_load_4d(dst + DZ_TILE, &a_map, full_bar(st), 0, (t << P.PX_l2) >> 5, grp << P.NP_l2, 0);
                }
            }
            __syncwarp();
            if (++st == P.n_stage) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16_mn(128, CI);
        const bool leader = elect_one();
        int st = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_local; ++i) {
            mbar_wait(full_bar(st), ph);
            tc_fence_after();
            const uint32_t base = smem_u32(st_s) + (uint32_t)st * STAGE;
            // K-direction stride -> LBO, channel-chunk (M/N) stride -> SBO
            const uint64_t dz_desc0 = P.swap_lbo_sbo ? umma_desc(base, DZ_PLANE, 128) : umma_desc(base, 128, DZ_PLANE);
            const uint64_t a_desc0 = P.swap_lbo_sbo ? umma_desc(base + DZ_TILE, A_PLANE, LBO_A) : umma_desc(base + DZ_TILE, LBO_A, A_PLANE);
            if (leader) {
#pragma unroll
                for (int tp = 0; tp < TPR; ++tp) {
                    const int tap = role * TPR + tp;       // role = dy when the taps are split by kernel row
                    const uint32_t dy = TAPS == 9 ? (ROLES == 3 ? (uint32_t)role : (uint32_t)(tap / 3)) : 0u;
                    const uint32_t dx = TAPS == 9 ? (uint32_t)(tap % 3) : 0u;
                    const uint32_t tap_off = (dy * NP * kPitch + dx) * 16;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {          // 8 x 16 pixels = the 128-pixel tile
                        const uint64_t ad = dz_desc0 + (uint64_t)((uint32_t)(j * 256) >> 4);
                        const uint64_t bd = a_desc0 + (uint64_t)((tap_off + (uint32_t)j * 2 * LBO_A) >> 4);
                        umma_bf16(tmem_base + (uint32_t)(tp * CI), ad, bd, idesc, (i | j) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(st));
                if (i == n_local - 1) umma_commit(done_bar);
            }
            __syncwarp();
            if (++st == P.n_stage) { st = 0; ph ^= 1; }
        }
    } else if (n_local > 0) {
        // ------------------------------------------------ final epilogue: TMEM -> fp32 atomics
        const int q = warp & 3;
        const int co = q * 32 + lane;
        mbar_wait(done_bar, 0);
        tc_fence_after();
        if (co < CO) {       // warp-uniform (CO is 64 or 128)
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int tp = 0; tp < TPR; ++tp) {
                const int tap = role * TPR + tp;
                float* const dst = P.dw + (size_t)tap * CI * CO + co;
                if (CI == 16) {
                    uint32_t v[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(t_row + (uint32_t)(tp * CI))
                                 : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; ++c) atomicAdd(dst + (size_t)c * CO, __uint_as_float(v[c]));
                } else {
#pragma unroll 1
                    for (int c0 = 0; c0 < CI; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(t_row + (uint32_t)(tp * CI + c0), v);
#pragma unroll
                        for (int c = 0; c < 32; ++c) atomicAdd(dst + (size_t)(c0 + c) * CO, __uint_as_float(v[c]));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace dmf
