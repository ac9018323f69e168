// K3: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM, operands fed by TMA) for the dense layers of GMFNet.
//
// Activation layout in HBM ("C8 planar"): [N patches][C/8][H][W][8] 16-bit floats (fp16 at inference, bf16 in training) — 16-byte channel chunks,
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
// Warp roles (64 + 128*G threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), then G
// epilogue groups of 4 warps (tcgen05.ld -> BN affine -> ReLU -> bf16 -> 2x2 max-pool via warp
// shuffles -> 16-byte stores in the next layer's layout).  TMEM holds G accumulators; group g drains
// the tiles with (tile % G == g), so G epilogues overlap the MMAs of the following tiles.  (The ncu
// profile of the first version, one group, showed the tensor pipe 52 % active: a single epilogue warp
// per scheduler is latency-bound at ~1400 dependent instructions per tile.)  The A halo ring has
// 2..6 stages.
#pragma once
#include "common.cuh"

namespace dmf {
namespace tc {

constexpr int kPitch = 10;                 // halo row pitch in pixels (8 + 2)
constexpr uint64_t kSpinLimit = 4000000000ull;   // ~2 s of SM clocks, then trap instead of hanging the GPU

struct ConvParams {
    // every geometric quantity is a power of two (p in {8,16,32}); *_l2 are the exponents
    int S, S_l2;      // input map is S x S
    int NP, NP_l2;    // patches per tile
    int TH;           // output rows per tile (3x3 path)
    int tiles_x_l2;
    int PX, PX_l2;    // 1x1 path: pixels of one patch per tile
    int tpg_l2;       // tiles per patch group
    int n_tiles;
    int N;            // patches
    int a_plane;      // bytes of one channel-chunk plane inside an A stage
    int a_stage;      // bytes of one A stage (TMA box bytes, 128-byte multiple)
    int n_stage;      // A ring depth
    int sbo_a;        // bytes between 8-pixel row groups of A
    int out_chunks;   // channel chunks of the output tensor
    int out_chunk0;   // first chunk this layer writes
    int dbg;          // diagnostics only: bit0 = skip the A-tile TMA loads, bit1 = skip the epilogue math/stores
    int f16_in;       // MODE 0: both MMA operands are fp16 (the inference activations / weights); 0 = bf16 (the hi/lo-split MS stem operands)
    const __nv_bfloat16* w;     // packed [tap][C_in/8][C_out][8]
    const float* scale;         // folded BatchNorm scale  [C_out]
    const float* shift;         // folded BatchNorm shift  [C_out]
    __nv_bfloat16* out;
    float* gap;                 // GAPOUT layers: per-patch channel sums [N][C_out] fp32 (zeroed by the caller)
    double* stats;              // MODE 2: sum(z) at [c], sum(z^2) at [stat_stride + c] (zeroed by the caller)
    int stat_stride;
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
// INFERENCE runs on FP16 operands (activations and weights; fp32 accumulation in TMEM), training on bf16.  kind::f16 multiplies
// fp16 and bf16 at the same rate, but both operands must have the SAME format (measured: a bf16 A with an fp16 B raises an illegal
// instruction).  Why fp16: on a fitted GMFNet whose predictions vary (tests/test_gpu_parity_fitted.py, tools/agreement_probe.py) bf16
// operands leave the logits off by ~1.3 % of the logit scale, 1.4 % of it from rounding the WEIGHTS to 8 mantissa bits, and the
// argmax agreement with the fp32 reference at 99.5 - 99.85 %, below the 99.9 % bar; fp16's 11 bits bring the error to ~0.2 %.
// The inputs are normalised to [0, 1] and every layer is BatchNorm-scaled, so fp16's range (65504, conversions saturate) is ample.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) { return umma_idesc_bf16(M, N) & ~((7u << 7) | (7u << 10)); }
// the fp16 rounding of a weight, carried in the 16-bit container type of the packed tensors
static inline __nv_bfloat16 w16(float v) {
    v = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);
    const __half h = __float2half_rn(v);
    __nv_bfloat16 r;
    memcpy(&r, &h, 2);
    return r;
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
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
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
}
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
// fp32 pair -> packed fp16 pair, round-to-nearest, saturating at +-65504; the RELU form also clamps negatives (and NaN) to 0:
// one F2FP.SATFINITE.RELU.F16.F32.PACK_AB instead of two FMNMX + F2FP
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16x2_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t max_f16x2(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
// bit pattern of the saturating fp16 rounding of one value
__device__ __forceinline__ uint32_t f16_bits(float v) { return pack_f16x2(v, 0.f) & 0xFFFFu; }
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

// Sum over the 32 lanes of a warp of 32 per-lane values: exchange-halving (31 shuffles instead of 160);
// afterwards f[0] of lane l is the warp total of element l.
__device__ __forceinline__ void warp_channel_sums(float (&f)[32], int lane) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const bool up = lane & d;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < d) {
                const float send = up ? f[k] : f[k + d], keep = up ? f[k + d] : f[k];
                f[k] = keep + __shfl_xor_sync(0xffffffffu, send, d);
            }
        }
    }
}

// GAPOUT (1x1 layers whose 128-pixel tile holds whole patches): instead of storing the activation, the
// epilogue reduces it over the pixels of each patch (global average pool fused in) with an
// exchange-halving warp reduction and emits per-patch channel sums; the activation never goes to HBM.
//
// MODE (training path, train.cu): 0 = inference epilogue (folded BN + ReLU, optional pool / GAP);
// 1 = RAW: the bf16 rounding of the fp32 accumulator is stored unpooled (dgrad, and the pre-BatchNorm
// tensor Z of a training forward); 2 = RAW + batch statistics: additionally sum(z) and sum(z^2) per
// output channel over everything this launch computes are added (double atomics) to P.stats[2][C_OUT]
// — train-mode BatchNorm needs them before anything can be normalised.
template <int C_IN, int C_OUT, int TAPS, bool POOL, int G, int NP, bool GAPOUT = false, int MODE = 0>
__global__ void __launch_bounds__(64 + 128 * G, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap in_map, const ConvParams P) {
    constexpr int kThreads = 64 + 128 * G;
    constexpr int KCH = C_IN / 8;
    constexpr int KSTEPS = C_IN / 16;
    constexpr uint32_t WBYTES = (uint32_t)TAPS * C_IN * C_OUT * 2;
    constexpr uint32_t TMEM_USED = G * C_OUT;          // G accumulators ...
    constexpr uint32_t TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static_assert(G >= 1 && G <= 4 && TMEM_USED <= 512, "epilogue groups / TMEM columns");   // ... in a power-of-two allocation
    static_assert(C_IN % 16 == 0 && C_OUT % 32 == 0 && C_OUT <= 256, "channel counts");
    // A-tile geometry is compile-time so that every UMMA descriptor is "base + immediate"
    constexpr int TH = TAPS == 9 ? 16 / NP : 0;
    constexpr uint32_t A_PLANE = TAPS == 9 ? (uint32_t)(TH + 2) * NP * kPitch * 16 : 128u * 16u;
    constexpr uint32_t SBO_A = TAPS == 9 ? (uint32_t)kPitch * 16 : 128u;
    constexpr uint32_t A_STAGE = KCH * A_PLANE;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + WBYTES;
    float* scale_s = reinterpret_cast<float*>(a_s + (size_t)P.n_stage * P.a_stage);
    float* shift_s = scale_s + C_OUT;
    float* stat_s = shift_s + C_OUT;                  // [2][C_OUT] per-CTA partial statistics (MODE 2)
    uint64_t* bars = reinterpret_cast<uint64_t*>(stat_s + 2 * C_OUT);
    // bars: [0,8) full, [8,16) empty, 16 weights, 17..20 tmem_full, 21..24 tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t w_bar = bar0 + 8u * 16;
    auto tfull_bar = [&](int a) { return bar0 + 8u * (17 + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (21 + a); };

    for (int i = threadIdx.x; i < C_OUT; i += kThreads) {
        if (MODE == 0) {
            scale_s[i] = P.scale[i];
            shift_s[i] = P.shift[i];
        }
        stat_s[i] = 0.f;
        stat_s[C_OUT + i] = 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stage; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(w_bar, 1);
        for (int a = 0; a < G; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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
        // ------------------------------------------------ TMA producer (whole warp converged, one lane issues)
        const bool leader = elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&in_map) : "memory");
            mbar_expect_tx(w_bar, WBYTES);
            constexpr uint32_t CH = 16384;
            for (uint32_t off = 0; off < WBYTES; off += CH)
                bulk_load(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(P.w) + off, min(CH, WBYTES - off), w_bar);
        }
        __syncwarp();
        int st = 0;
        uint32_t ph = 1;                          // parity to wait on for "stage free"; flips when the ring wraps
        for (int i = 0; i < n_local; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            mbar_wait(empty_bar(st), ph);
            if (leader) {
                if (P.dbg & 1) {
                    mbar_arrive(full_bar(st));
                } else {
                    mbar_expect_tx(full_bar(st), A_STAGE);
                    const int grp = tile >> P.tpg_l2, t = tile & ((1 << P.tpg_l2) - 1);
                    const uint32_t dst = smem_u32(a_s) + (uint32_t)st * A_STAGE;
                    if (TAPS == 9) {
                        const int ty = t >> P.tiles_x_l2, tx = t & ((1 << P.tiles_x_l2) - 1);
                        tma_load_4d(dst, &in_map, full_bar(st), (tx * 8 - 1) * 8, grp * NP, ty * TH - 1, 0);
                    } else {
                        // dims: (8 ch x 32 px merged, 32-px block, patch, chunk)
                        tma_load_4d(dst, &in_map, full_bar(st), 0, (t << P.PX_l2) >> 5, grp << P.NP_l2, 0);
                    }
                }
            }
            __syncwarp();
            if (++st == P.n_stage) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (whole warp converged, one lane issues)
        const uint32_t idesc = (MODE == 0 && P.f16_in) ? umma_idesc_f16(128, C_OUT) : umma_idesc_bf16(128, C_OUT);
        const bool leader = elect_one();
        mbar_wait(w_bar, 0);
        // descriptor = base + (byte offset >> 4): the offsets below are compile-time immediates
        const uint64_t w_desc0 = umma_desc(smem_u32(w_s), C_OUT * 16, 128);
        int st = 0;
        uint32_t ph = 0;
        // (Tried and measured on B200: interleaving two tiles, or split-K over two accumulators per
        // tile, lifts the bare MMA chain from ~94 to ~75 clk per M128xN128xK16 instruction — the gap
        // is the accumulate read-after-write — but the first needs more A stages than fit beside the
        // resident weights and the second doubles the epilogue's TMEM reads; both lost end to end.)
        for (int i = 0; i < n_local; ++i) {
            const int acc = i % G;
            mbar_wait(tempty_bar(acc), ((i / G) & 1) ^ 1);
            mbar_wait(full_bar(st), ph);
            tc_fence_after();
            const uint64_t a_desc0 = umma_desc(smem_u32(a_s) + (uint32_t)st * A_STAGE, A_PLANE, SBO_A);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C_OUT);
            if (leader) {
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    const uint32_t tap_off = TAPS == 9 ? (uint32_t)(((tap / 3) * NP * kPitch + (tap % 3)) * 16) : 0u;
#pragma unroll
                    for (int j = 0; j < KSTEPS; ++j) {
                        const uint64_t ad = a_desc0 + (uint64_t)(((uint32_t)(2 * j) * A_PLANE + tap_off) >> 4);
                        const uint64_t bd = w_desc0 + (uint64_t)((uint32_t)((tap * KCH + 2 * j) * C_OUT * 16) >> 4);
                        umma_bf16(d_tmem, ad, bd, idesc, (tap | j) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(st));      // halo stage reusable once these MMAs retire
                umma_commit(tfull_bar(acc));     // accumulator ready for the epilogue
            }
            __syncwarp();
            if (++st == P.n_stage) { st = 0; ph ^= 1; }
        }
    } else {
        // ------------------------------------------------ epilogue (G groups of 4 warps)
        const int eg = (warp - 2) >> 2;          // epilogue group = accumulator this warp drains
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;             // accumulator row = pixel of the tile
        const int So_l2 = POOL ? P.S_l2 - 1 : P.S_l2;
        constexpr int hx = 8 * NP;               // lane distance of the vertical pooling neighbour
        const int sub = (lane & 1) | (((lane / hx) & 1) << 1);
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * C_OUT);
        for (int i = eg; i < n_local; i += G) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int grp = tile >> P.tpg_l2, t = tile & ((1 << P.tpg_l2) - 1);
            int n, h, w;
            if (TAPS == 9) {
                const int ty = t >> P.tiles_x_l2, tx = t & ((1 << P.tiles_x_l2) - 1);
                const int g = m >> 3;
                n = grp * NP + (g % NP);
                h = ty * TH + (g / NP);
                w = tx * 8 + (m & 7);
            } else {
                const int px = (t << P.PX_l2) + (m & (P.PX - 1));
                n = (grp << P.NP_l2) + (m >> P.PX_l2);
                h = px >> P.S_l2;
                w = px & (P.S - 1);
            }
            const bool valid = n < P.N;
            // element offset of (n, chunk 0, h', w', 0); one channel chunk is So*So*8 elements further
            const int64_t chunk0 = (int64_t)n * P.out_chunks + P.out_chunk0;
            const int hh = POOL ? h >> 1 : h, ww = POOL ? w >> 1 : w;
            __nv_bfloat16* const obase = P.out + (((chunk0 << So_l2) + hh) << So_l2) * 8 + ww * 8;
            const int64_t cstride = (int64_t)8 << (2 * So_l2);
            mbar_wait(tfull_bar(eg), (i / G) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < ((P.dbg & 2) ? 0 : C_OUT); c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t_row + c0, v);
                uint32_t pk[16];
                if (MODE != 0) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) pk[k] = pack_bf16x2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
                    if (valid) {
#pragma unroll
                        for (int s4 = 0; s4 < 4; ++s4)
                            *reinterpret_cast<uint4*>(obase + ((c0 >> 3) + s4) * cstride) =
                                make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]);
                    }
                    if (MODE == 2) {
                        // rows of patches beyond N come from out-of-bounds TMA boxes (train.cu encodes its maps per batch size): all zeros
                        float f[32], q2[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) { f[k] = __uint_as_float(v[k]); q2[k] = f[k] * f[k]; }
                        warp_channel_sums(f, lane);
                        warp_channel_sums(q2, lane);
                        atomicAdd(stat_s + c0 + lane, f[0]);
                        atomicAdd(stat_s + C_OUT + c0 + lane, q2[0]);
                    }
                    continue;
                }
                const float4* sc4 = reinterpret_cast<const float4*>(scale_s + c0);
                const float4* sh4 = reinterpret_cast<const float4*>(shift_s + c0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 sc = sc4[k], sh = sh4[k];
                    const float a0 = fmaf(__uint_as_float(v[4 * k]), sc.x, sh.x);
                    const float a1 = fmaf(__uint_as_float(v[4 * k + 1]), sc.y, sh.y);
                    const float a2 = fmaf(__uint_as_float(v[4 * k + 2]), sc.z, sh.z);
                    const float a3 = fmaf(__uint_as_float(v[4 * k + 3]), sc.w, sh.w);
                    pk[2 * k] = pack_f16x2_relu(a0, a1);          // ReLU + fp16 rounding in one instruction
                    pk[2 * k + 1] = pack_f16x2_relu(a2, a3);
                }
                if (GAPOUT) {
                    // fp32 post-ReLU values of this pixel for channels c0..c0+31 -> sums over the patch's pixels.
                    // PX >= 32: the warp's 32 lanes belong to one patch; after 5 exchange-halving steps lane l owns
                    // channel c0 + l.  PX == 16: two patches per warp, 4 steps, lane l owns channels c0 + 2(l&15), +1.
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 sc = sc4[k], sh = sh4[k];
                        f[4 * k] = fmaxf(fmaf(__uint_as_float(v[4 * k]), sc.x, sh.x), 0.f);
                        f[4 * k + 1] = fmaxf(fmaf(__uint_as_float(v[4 * k + 1]), sc.y, sh.y), 0.f);
                        f[4 * k + 2] = fmaxf(fmaf(__uint_as_float(v[4 * k + 2]), sc.z, sh.z), 0.f);
                        f[4 * k + 3] = fmaxf(fmaf(__uint_as_float(v[4 * k + 3]), sc.w, sh.w), 0.f);
                    }
                    const bool wide = P.PX >= 32;
                    if (wide) {
                        const bool up = lane & 16;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const float send = up ? f[k] : f[k + 16], keep = up ? f[k + 16] : f[k];
                            f[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
                    }
                    // live values per lane before a step with lane distance d: 2d (wide) or 4d (two patches per warp)
#pragma unroll
                    for (int d = 8; d >= 1; d >>= 1) {
                        const bool up = lane & d;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            if (k < 2 * d && (k < d || !wide)) {
                                const float fh = wide ? f[k + d] : f[k + 2 * d];
                                const float send = up ? f[k] : fh, keep = up ? fh : f[k];
                                f[k] = keep + __shfl_xor_sync(0xffffffffu, send, d);
                            }
                        }
                    }
                    if (valid) {
                        if (wide) atomicAdd(P.gap + (int64_t)n * C_OUT + c0 + lane, f[0]);     // PX/32 warps add: 2 addends at p=16 -> order-independent
                        else *reinterpret_cast<float2*>(P.gap + (int64_t)n * C_OUT + c0 + 2 * (lane & 15)) = make_float2(f[0], f[1]);
                    }
                } else if (POOL) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        pk[k] = max_f16x2(pk[k], __shfl_xor_sync(0xffffffffu, pk[k], 1));
                        pk[k] = max_f16x2(pk[k], __shfl_xor_sync(0xffffffffu, pk[k], hx));
                    }
                    // the 4 lanes of a 2x2 window now hold the same 32 pooled channels: each stores one 8-channel chunk
                    uint4 o;
                    o.x = sub == 0 ? pk[0] : sub == 1 ? pk[4] : sub == 2 ? pk[8] : pk[12];
                    o.y = sub == 0 ? pk[1] : sub == 1 ? pk[5] : sub == 2 ? pk[9] : pk[13];
                    o.z = sub == 0 ? pk[2] : sub == 1 ? pk[6] : sub == 2 ? pk[10] : pk[14];
                    o.w = sub == 0 ? pk[3] : sub == 1 ? pk[7] : sub == 2 ? pk[11] : pk[15];
                    if (valid) *reinterpret_cast<uint4*>(obase + ((c0 >> 3) + sub) * cstride) = o;
                } else if (valid) {
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4)
                        *reinterpret_cast<uint4*>(obase + ((c0 >> 3) + s4) * cstride) =
                            make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(eg));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    if (MODE == 2)
        for (int i = threadIdx.x; i < 2 * C_OUT; i += kThreads)
            atomicAdd(P.stats + (i < C_OUT ? i : P.stat_stride + i - C_OUT), (double)stat_s[i]);
}

// ------------------------------------------------------------------------------------------------
// Row-pair variant for thin layers (C_OUT = 64): a tcgen05.mma costs ~30 clk + operand bytes / 128 B/clk
// (measured), so N = 64 instructions run the pipe at ~40 %.  Here one accumulator row stands for TWO
// vertically adjacent output pixels (h, w), (h+1, w) with h even: the N dimension is (s, co) = 2 x 64 and
// the filter becomes a 4 x 3 tap window, W'[(s, co)][dy', dx] = W[co][dy' - s, dx] (zero outside).
// 24 M128 x N128 x K16 MMAs per 256 output pixels instead of 36 M128 x N64 ones, and the vertical half of
// the 2x2 max-pool is thread-local (columns c and 64 + c of the same TMEM lane).
// Tile = 16 row pairs x 8 columns (32 image rows x 8 columns); halo 34 x 10 pixels; row-group stride of
// the A descriptor = two halo rows.
template <int C_IN, int G>
__global__ void __launch_bounds__(64 + 128 * G, 1) conv_rowpair_kernel(const __grid_constant__ CUtensorMap in_map, const ConvParams P) {
    constexpr int kThreads = 64 + 128 * G;
    constexpr int C_OUT = 64, N2 = 2 * C_OUT, TAPS = 12;
    constexpr int KCH = C_IN / 8, KSTEPS = C_IN / 16;
    constexpr int TH = 32;
    constexpr uint32_t WBYTES = (uint32_t)TAPS * C_IN * N2 * 2;
    constexpr uint32_t TMEM_USED = G * N2;
    constexpr uint32_t TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    constexpr uint32_t A_PLANE = (uint32_t)(TH + 2) * kPitch * 16;
    constexpr uint32_t SBO_A = 2u * kPitch * 16;
    constexpr uint32_t A_STAGE = KCH * A_PLANE;
    static_assert(G >= 1 && G <= 4 && C_IN % 16 == 0, "row-pair conv configuration");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + WBYTES;
    float* scale_s = reinterpret_cast<float*>(a_s + (size_t)P.n_stage * A_STAGE);
    float* shift_s = scale_s + C_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(shift_s + C_OUT);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t w_bar = bar0 + 8u * 16;
    auto tfull_bar = [&](int a) { return bar0 + 8u * (17 + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (21 + a); };

    for (int i = threadIdx.x; i < C_OUT; i += kThreads) {
        scale_s[i] = P.scale[i];
        shift_s[i] = P.shift[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stage; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(w_bar, 1);
        for (int a = 0; a < G; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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
    const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // tile -> (patch, ty, tx): tiles_x = S/8 column strips, tiles_y = S/32 row blocks, both powers of two
    const int tx_mask = (1 << P.tiles_x_l2) - 1;

    if (warp == 0) {
        const bool leader = elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&in_map) : "memory");
            mbar_expect_tx(w_bar, WBYTES);
            constexpr uint32_t CH = 16384;
            for (uint32_t off = 0; off < WBYTES; off += CH)
                bulk_load(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(P.w) + off, min(CH, WBYTES - off), w_bar);
        }
        __syncwarp();
        int st = 0;
        uint32_t ph = 1;
        for (int i = 0; i < n_local; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            mbar_wait(empty_bar(st), ph);
            if (leader) {
                mbar_expect_tx(full_bar(st), A_STAGE);
                const int n = tile >> P.tpg_l2, t = tile & ((1 << P.tpg_l2) - 1);
                const int ty = t >> P.tiles_x_l2, tx = t & tx_mask;
                tma_load_4d(smem_u32(a_s) + (uint32_t)st * A_STAGE, &in_map, full_bar(st), (tx * 8 - 1) * 8, n, ty * TH - 1, 0);
            }
            __syncwarp();
            if (++st == P.n_stage) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_f16(128, N2);             // inference only: fp16 operands
        const bool leader = elect_one();
        mbar_wait(w_bar, 0);
        const uint64_t w_desc0 = umma_desc(smem_u32(w_s), N2 * 16, 128);
        int st = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_local; ++i) {
            const int acc = i % G;
            mbar_wait(tempty_bar(acc), ((i / G) & 1) ^ 1);
            mbar_wait(full_bar(st), ph);
            tc_fence_after();
            const uint64_t a_desc0 = umma_desc(smem_u32(a_s) + (uint32_t)st * A_STAGE, A_PLANE, SBO_A);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N2);
            if (leader) {
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    const uint32_t tap_off = (uint32_t)(((tap / 3) * kPitch + (tap % 3)) * 16);
#pragma unroll
                    for (int j = 0; j < KSTEPS; ++j) {
                        const uint64_t ad = a_desc0 + (uint64_t)(((uint32_t)(2 * j) * A_PLANE + tap_off) >> 4);
                        const uint64_t bd = w_desc0 + (uint64_t)((uint32_t)((tap * KCH + 2 * j) * N2 * 16) >> 4);
                        umma_bf16(d_tmem, ad, bd, idesc, (tap | j) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(st));
                umma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++st == P.n_stage) { st = 0; ph ^= 1; }
        }
    } else {
        const int eg = (warp - 2) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int So_l2 = P.S_l2 - 1;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * N2);
        const int half = lane & 1;                       // the two lanes of a horizontal pair share the stores
        for (int i = eg; i < n_local; i += G) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int n = tile >> P.tpg_l2, t = tile & ((1 << P.tpg_l2) - 1);
            const int ty = t >> P.tiles_x_l2, tx = t & tx_mask;
            const int prow = ty * (TH / 2) + (m >> 3);   // pooled row = row-pair index
            const int pcol = (tx * 8 + (m & 7)) >> 1;
            const bool valid = n < P.N;
            const int64_t chunk0 = (int64_t)n * P.out_chunks + P.out_chunk0;
            __nv_bfloat16* const obase = P.out + (((chunk0 << So_l2) + prow) << So_l2) * 8 + pcol * 8;
            const int64_t cstride = (int64_t)8 << (2 * So_l2);
            mbar_wait(tfull_bar(eg), (i / G) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < C_OUT; c0 += 32) {
                uint32_t v0[32], v1[32];
                tmem_ld32_nowait(t_row + c0, v0);                // row h     (s = 0)
                tmem_ld32_nowait(t_row + C_OUT + c0, v1);        // row h + 1 (s = 1)
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t pk[16];
                const float4* sc4 = reinterpret_cast<const float4*>(scale_s + c0);
                const float4* sh4 = reinterpret_cast<const float4*>(shift_s + c0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 sc = sc4[k], sh = sh4[k];
                    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
                    float r[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        r[e] = fmaxf(fmaf(__uint_as_float(v0[4 * k + e]), scv[e], shv[e]), fmaf(__uint_as_float(v1[4 * k + e]), scv[e], shv[e]));
                    pk[2 * k] = pack_f16x2_relu(r[0], r[1]);
                    pk[2 * k + 1] = pack_f16x2_relu(r[2], r[3]);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) pk[k] = max_f16x2(pk[k], __shfl_xor_sync(0xffffffffu, pk[k], 1));
                if (valid) {   // even lane: chunks 0,1 of this 32-channel group; odd lane: chunks 2,3
                    const int cb = (c0 >> 3) + 2 * half;
                    *reinterpret_cast<uint4*>(obase + cb * cstride) =
                        half ? make_uint4(pk[8], pk[9], pk[10], pk[11]) : make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(obase + (cb + 1) * cstride) =
                        half ? make_uint4(pk[12], pk[13], pk[14], pk[15]) : make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(eg));
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
