// K3: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM, operands fed by TMA) for the dense layers of GMFNet.
//
// Activation layout in HBM ("C8 planar"): [N patches][C/8][H][W][8] bf16 — 16-byte channel chunks,
// each chunk a dense H x W plane.  This is exactly the UMMA K-major NO-SWIZZLE operand layout when a
// tile of it is dropped into shared memory: a core matrix is 8 consecutive pixels x 16 bytes = 128
// contiguous bytes, core matrices that are neighbours in K are one plane apart (LBO) and core
// matrices that are neighbours in M are one row-group apart (SBO).  Because nothing is swizzled,
// a 3x3 tap is just a different START ADDRESS into the same halo tile: the (TH+2) x 10 pixel halo
// of a TH x 8 output tile is loaded ONCE by TMA (out-of-bounds coordinates give the conv's zero
// padding for free) and all 9 taps x C_in/16 k-steps read shifted views of it, so L2->SM traffic
// is ~1.4x the activation size instead of 9x.  Weights for all taps stay resident in shared memory
// for the life of the persistent CTA.
//
// GEMM view per tile: D[128 pixels, C_out] += A_tap[128 pixels, 16 ch] * W_tap[C_out, 16 ch]^T,
// M = 128, N = C_out, K = 16 per tcgen05.mma, 9*C_in/16 MMAs per tile, fp32 accumulate in TMEM.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..5 = epilogue (tcgen05.ld -> BN affine -> ReLU -> bf16 -> 2x2 max-pool via warp shuffles
// -> 16-byte stores in the next layer's layout).  TMEM holds two accumulators so the epilogue of tile
// i overlaps the MMAs of tile i+1; the A halo ring has 2..6 stages.
#pragma once
#include "common.cuh"

namespace dmf {
namespace tc {

constexpr int kThreads = 192;
constexpr int kPitch = 10;                 // halo row pitch in pixels (8 + 2)
constexpr uint64_t kSpinLimit = 4000000000ull;   // ~2 s of SM clocks, then trap instead of hanging the GPU

struct ConvParams {
    int S;            // input map is S x S
    int NP;           // patches per tile
    int TH;           // output rows per tile (3x3 path)
    int tiles_x, tiles_y;
    int PX;           // 1x1 path: pixels of one patch per tile
    int tiles_per_group;
    int n_tiles;
    int N;            // patches
    int a_plane;      // bytes of one channel-chunk plane inside an A stage
    int a_stage;      // bytes of one A stage (TMA box bytes, 128-byte multiple)
    int n_stage;      // A ring depth
    int sbo_a;        // bytes between 8-pixel row groups of A
    int out_chunks;   // channel chunks of the output tensor
    int out_chunk0;   // first chunk this layer writes
    const __nv_bfloat16* w;     // packed [tap][C_in/8][C_out][8]
    const float* scale;         // folded BatchNorm scale  [C_out]
    const float* shift;         // folded BatchNorm shift  [C_out]
    __nv_bfloat16* out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if ((uint64_t)(clock64() - t0) > kSpinLimit) __trap();
    }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"((uint64_t)src), "r"(bytes), "r"(bar)
                 : "memory");
}

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE (layout_type 0), sm_100 version bit.
// bits [0,14) start>>4, [16,30) LBO>>4 (K-neighbour core matrix), [32,46) SBO>>4 (M/N-neighbour group)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

template <int C_IN, int C_OUT, int TAPS, bool POOL>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap in_map, const ConvParams P) {
    constexpr int KCH = C_IN / 8;
    constexpr int KSTEPS = C_IN / 16;
    constexpr uint32_t WBYTES = (uint32_t)TAPS * C_IN * C_OUT * 2;
    constexpr uint32_t TMEM_COLS = 2 * C_OUT;          // two accumulators; power of two >= 32
    static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");
    static_assert(C_IN % 16 == 0 && C_OUT % 32 == 0 && C_OUT <= 256, "channel counts");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + WBYTES;
    float* scale_s = reinterpret_cast<float*>(a_s + (size_t)P.n_stage * P.a_stage);
    float* shift_s = scale_s + C_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(shift_s + C_OUT);
    // bars: [0,8) full, [8,16) empty, 16 weights, 17..18 tmem_full, 19..20 tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t w_bar = bar0 + 8u * 16;
    auto tfull_bar = [&](int a) { return bar0 + 8u * (17 + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (19 + a); };

    for (int i = threadIdx.x; i < C_OUT; i += kThreads) {
        scale_s[i] = P.scale[i];
        shift_s[i] = P.shift[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stage; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(w_bar, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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

    const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&in_map) : "memory");
            mbar_expect_tx(w_bar, WBYTES);
            constexpr uint32_t CH = 16384;
            for (uint32_t off = 0; off < WBYTES; off += CH)
                bulk_load(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(P.w) + off, min(CH, WBYTES - off), w_bar);
            for (int i = 0; i < n_local; ++i) {
                const int tile = blockIdx.x + i * gridDim.x;
                const int st = i % P.n_stage;
                mbar_wait(empty_bar(st), ((i / P.n_stage) & 1) ^ 1);
                mbar_expect_tx(full_bar(st), P.a_stage);
                const int grp = tile / P.tiles_per_group, t = tile - grp * P.tiles_per_group;
                const uint32_t dst = smem_u32(a_s + (size_t)st * P.a_stage);
                if (TAPS == 9) {
                    const int ty = t / P.tiles_x, tx = t - ty * P.tiles_x;
                    tma_load_4d(dst, &in_map, full_bar(st), (tx * 8 - 1) * 8, grp * P.NP, ty * P.TH - 1, 0);
                } else {
                    tma_load_4d(dst, &in_map, full_bar(st), 0, t * P.PX, grp * P.NP, 0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, C_OUT);
            mbar_wait(w_bar, 0);
            const uint32_t w_addr = smem_u32(w_s);
            for (int i = 0; i < n_local; ++i) {
                const int st = i % P.n_stage, acc = i & 1;
                mbar_wait(tempty_bar(acc), ((i >> 1) & 1) ^ 1);
                mbar_wait(full_bar(st), (i / P.n_stage) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(a_s + (size_t)st * P.a_stage);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C_OUT);
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const uint32_t tap_off = TAPS == 9 ? (uint32_t)((dy * P.NP * kPitch + dx) * 16) : 0u;
#pragma unroll
                    for (int j = 0; j < KSTEPS; ++j) {
                        const uint64_t ad = umma_desc(a_addr + (uint32_t)(2 * j) * P.a_plane + tap_off, P.a_plane, P.sbo_a);
                        const uint64_t bd = umma_desc(w_addr + (uint32_t)((tap * KCH + 2 * j) * C_OUT * 16), C_OUT * 16, 128);
                        umma_bf16(d_tmem, ad, bd, idesc, (tap | j) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(st));      // halo stage reusable once these MMAs retire
                umma_commit(tfull_bar(acc));     // accumulator ready for the epilogue
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;             // accumulator row = pixel of the tile
        const int So = POOL ? P.S / 2 : P.S;
        for (int i = 0; i < n_local; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int acc = i & 1;
            const int grp = tile / P.tiles_per_group, t = tile - grp * P.tiles_per_group;
            int n, h, w;
            if (TAPS == 9) {
                const int ty = t / P.tiles_x, tx = t - ty * P.tiles_x;
                const int g = m >> 3;
                n = grp * P.NP + g % P.NP;
                h = ty * P.TH + g / P.NP;
                w = tx * 8 + (m & 7);
            } else {
                const int px = t * P.PX + m % P.PX;
                n = grp * P.NP + m / P.PX;
                h = px / P.S;
                w = px - h * P.S;
            }
            const bool valid = n < P.N;
            mbar_wait(tfull_bar(acc), (i >> 1) & 1);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C_OUT);
#pragma unroll 1
            for (int c0 = 0; c0 < C_OUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t_row + c0, v);
                uint32_t pk[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    float a = fmaf(__uint_as_float(v[2 * k]), scale_s[c0 + 2 * k], shift_s[c0 + 2 * k]);
                    float b = fmaf(__uint_as_float(v[2 * k + 1]), scale_s[c0 + 2 * k + 1], shift_s[c0 + 2 * k + 1]);
                    pk[k] = pack_bf16x2(fmaxf(a, 0.f), fmaxf(b, 0.f));
                }
                if (POOL) {
                    const int hx = 8 * P.NP;     // lane distance of the vertical neighbour
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        pk[k] = max_bf16x2(pk[k], __shfl_xor_sync(0xffffffffu, pk[k], 1));
                        pk[k] = max_bf16x2(pk[k], __shfl_xor_sync(0xffffffffu, pk[k], hx));
                    }
                    // the 4 lanes of a 2x2 window now hold the same 32 pooled channels: each stores one 8-channel chunk
                    const int sub = (lane & 1) | (((lane / hx) & 1) << 1);
                    uint4 o;
                    o.x = sub == 0 ? pk[0] : sub == 1 ? pk[4] : sub == 2 ? pk[8] : pk[12];
                    o.y = sub == 0 ? pk[1] : sub == 1 ? pk[5] : sub == 2 ? pk[9] : pk[13];
                    o.z = sub == 0 ? pk[2] : sub == 1 ? pk[6] : sub == 2 ? pk[10] : pk[14];
                    o.w = sub == 0 ? pk[3] : sub == 1 ? pk[7] : sub == 2 ? pk[11] : pk[15];
                    if (valid) {
                        const int64_t chunk = (int64_t)n * P.out_chunks + P.out_chunk0 + (c0 >> 3) + sub;
                        *reinterpret_cast<uint4*>(P.out + ((chunk * So + (h >> 1)) * So + (w >> 1)) * 8) = o;
                    }
                } else if (valid) {
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const int64_t chunk = (int64_t)n * P.out_chunks + P.out_chunk0 + (c0 >> 3) + s4;
                        *reinterpret_cast<uint4*>(P.out + ((chunk * So + h) * So + w) * 8) =
                            make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
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
