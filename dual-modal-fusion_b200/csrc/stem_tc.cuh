// PAN stem on the tensor cores: conv3x3 1->32 (zero padding at the patch border) + BN + ReLU +
// maxpool2, straight from the fp32 PAN window of each patch (the K1 gather is fused in).
//
// A 1-channel 3x3 convolution has K = 9, far too thin for tcgen05 as a plain implicit GEMM, and the
// first CUDA-core version of this stem was the most expensive stage of the network (23 % of the
// time for 2 % of the FLOPs).  Here the operand is BUILT in shared memory, one row per POOLED pixel:
//   * a loader warp streams the fp32 window rows of the NEXT patch into a zero-bordered,
//     double-buffered staging tile with cp.async.bulk (one 16-byte-aligned row per copy);
//   * builder warps convert the window IN PLACE to packed {hi, lo} bf16 pairs (x = hi + lo to ~2^-16)
//     and then write, per pooled pixel, the 4x4 input region that its 2x2 pooling window touches as
//     one K = 48 row:  [ (hi_i, lo_i) i=0..15 | (hi_0,hi_1) ... (hi_14,hi_15) ];
//   * the B operand has N = 128 rows = 4 window positions x 32 channels: row (q, co) holds the 3x3
//     kernel of channel co placed at offset (qy, qx) inside the 4x4 region, as
//     [ (w_hi, w_hi) | w_lo ], i.e. x_hi*w_hi + x_lo*w_hi + x_hi*w_lo: fp32-grade products.  The folded
//     BatchNorm scale is multiplied into the fp32 weights before the split;
//   * three M128 x N128 x K16 MMAs per 128 pooled pixels put the four candidates of every pooled
//     pixel into four 32-column blocks of its own TMEM lane: the epilogue is max -> + shift -> ReLU ->
//     bf16 -> 16-byte stores, no shuffles.
// (An earlier layout with one K = 32 row per output pixel and 8 N = 32 MMAs per tile was shared-memory
// bandwidth bound: ~860 smem wavefronts per tile against ~420 here.)
// Roles (64 + 256 + 128*G threads): warp 0 = MMA issuer / TMEM owner, warp 1 = window loader,
// warps 2..9 = builders in NT teams (team k builds the tiles with tile % NT == k, so NT tiles are in
// flight and the LDS -> PRMT -> STS -> proxy-fence -> arrive latency chain of one tile overlaps the
// others; the ncu profile of the single-team version showed exactly that chain exposed), then G
// epilogue groups of 4 warps.  Persistent over patches; NT..4 stage A ring of 32 KB.
#pragma once
#include "conv_tc.cuh"

namespace dmf {
namespace tc {

constexpr int kStemCout = 32;
constexpr int kStemMaxStages = 4;
constexpr int kStemKch = 6;                      // K = 48 = 6 chunks of 8 bf16
constexpr int kStemStage = kStemKch * 128 * 16;  // A tile: 6 k-chunks x 128 rows x 16 B = 12 KB
constexpr int kStemWBytes = kStemKch * 128 * 16; // B: 6 k-chunks x (4 positions x 32 channels) x 16 B

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
    const __nv_bfloat16* w;        // packed [6 k-chunks][128 = (q, co)][8]
    const float* shift;            // folded BatchNorm shift (the scale lives in the weights)
    __nv_bfloat16* out;            // [N][4][2p][2p][8]
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}

template <int G, int NT>
__global__ void __launch_bounds__(320 + 128 * G, 1) stem_pan_tc_kernel(const StemPanParams P) {
    constexpr int kBuilders = 256;
    constexpr int kTeam = kBuilders / NT;            // threads per builder team
    constexpr int kRows = 128 / kTeam;               // tile rows per builder thread
    static_assert(NT == 2 || NT == 4, "builder teams");
    constexpr uint32_t TMEM_USED = G * 128;
    constexpr uint32_t TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int PW = 4 * P.p, S = 2 * P.p, RP = P.raw_pitch;
    const int raw_floats = (PW + 2) * RP;                                 // one staging buffer
    uint8_t* a_s = smem;                                                  // n_stage x kStemStage
    uint8_t* w_s = a_s + P.n_stage * kStemStage;                          // 12 KB
    float* shift_s = reinterpret_cast<float*>(w_s + kStemWBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(shift_s + 2 * kStemCout);    // full[4] empty[4] tfull[4] tempty[4] rawfull[2] rawempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* raw = reinterpret_cast<float*>(tmem_slot + 4);                 // 2 x (PW+2) x RP fp32, zero border

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (8 + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (12 + a); };
    auto rawfull_bar = [&](int b) { return bar0 + 8u * (16 + b); };
    auto rawempty_bar = [&](int b) { return bar0 + 8u * (18 + b); };

    for (int i = threadIdx.x; i < kStemWBytes / 4; i += blockDim.x)      // 12 KB of weights
        reinterpret_cast<uint32_t*>(w_s)[i] = reinterpret_cast<const uint32_t*>(P.w)[i];
    for (int i = threadIdx.x; i < 2 * raw_floats; i += blockDim.x) raw[i] = 0.f;   // borders stay zero for good
    if (threadIdx.x < kStemCout) shift_s[threadIdx.x] = P.shift[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stage; ++s) { mbar_init(full_bar(s), kTeam / 32); mbar_init(empty_bar(s), 1); }
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
        constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
        const bool leader = elect_one();
        const uint64_t w_desc0 = umma_desc(smem_u32(w_s), 128 * 16, 128);
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
                for (int j = 0; j < kStemKch / 2; ++j)
                    umma_bf16(d0, a_desc0 + (uint64_t)((2 * j * 128 * 16) >> 4), w_desc0 + (uint64_t)((2 * j * 128 * 16) >> 4), idesc,
                              j ? 1u : 0u);
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
        const int team = bt / kTeam, r0 = bt % kTeam;
        for (int64_t pl = 0; pl < n_patches; ++pl) {
            const int b = (int)(pl & 1);
            uint32_t* win = reinterpret_cast<uint32_t*>(raw + b * raw_floats);
            mbar_wait(rawfull_bar(b), (uint32_t)((pl >> 1) & 1));
            // fp32 -> packed {hi, lo} bf16, in place, interior only (the zero border is already 0 = {0, 0})
            for (int i = bt; i < PW * (PW / 4); i += kBuilders) {
                const int r = i / (PW / 4), c4 = i - r * (PW / 4);
                uint4* cell = reinterpret_cast<uint4*>(win + (r + 1) * RP + 4) + c4;
                const uint4 u = *cell;
                const float f[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float hi = __bfloat162float(__float2bfloat16_rn(f[e]));
                    o[e] = pack_bf16x2(hi, f[e] - hi);
                }
                *cell = make_uint4(o[0], o[1], o[2], o[3]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int t = team; t < tpp; t += NT) {
                const int64_t i = (pl << P.tpp_l2) + t;         // tile counter of this CTA -> ring slot and phase
                const int st = (int)(i % P.n_stage);
                mbar_wait(empty_bar(st), (uint32_t)((i / P.n_stage) & 1) ^ 1);
                uint8_t* stage = a_s + (size_t)st * kStemStage;
#pragma unroll
                for (int rr = 0; rr < kRows; ++rr) {
                    const int m = r0 + rr * kTeam;              // tile row = pooled pixel t*128 + m
                    const int pp = t * 128 + m;
                    const int prow = pp >> P.S_l2, pcol = pp & (S - 1);
                    // 4x4 input region of the pooled pixel: rows 2 prow - 1 .. 2 prow + 2, cols 2 pcol - 1 .. 2 pcol + 2
                    // = staging rows 2 prow .. 2 prow + 3, staging columns 2 pcol + 3 .. 2 pcol + 6
                    const uint32_t* wp = win + (2 * prow) * RP + 2 * pcol + 3;
                    uint32_t R[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) R[e] = wp[(e >> 2) * RP + (e & 3)];
                    uint4* dst = reinterpret_cast<uint4*>(stage) + m;
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[c * 128] = make_uint4(R[4 * c], R[4 * c + 1], R[4 * c + 2], R[4 * c + 3]);
                    dst[4 * 128] = make_uint4(__byte_perm(R[0], R[1], 0x5410), __byte_perm(R[2], R[3], 0x5410),
                                              __byte_perm(R[4], R[5], 0x5410), __byte_perm(R[6], R[7], 0x5410));
                    dst[5 * 128] = make_uint4(__byte_perm(R[8], R[9], 0x5410), __byte_perm(R[10], R[11], 0x5410),
                                              __byte_perm(R[12], R[13], 0x5410), __byte_perm(R[14], R[15], 0x5410));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(st));       // one arrival per builder warp of the team
            }
            mbar_arrive(rawempty_bar(b));                       // this thread no longer touches staging buffer b
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
            for (int hf = 0; hf < 2; ++hf) {                  // 16 channels at a time
                uint32_t v0[16], v1[16], v2[16], v3[16];
                tmem_ld16(t_row + hf * 16, v0);
                tmem_ld16(t_row + 32 + hf * 16, v1);
                tmem_ld16(t_row + 64 + hf * 16, v2);
                tmem_ld16(t_row + 96 + hf * 16, v3);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t pk[8];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 sh = *reinterpret_cast<const float4*>(shift_s + hf * 16 + 4 * k4);
                    const float shv[4] = {sh.x, sh.y, sh.z, sh.w};
                    float r[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = 4 * k4 + e;
                        const float mx = fmaxf(fmaxf(__uint_as_float(v0[k]), __uint_as_float(v1[k])),
                                               fmaxf(__uint_as_float(v2[k]), __uint_as_float(v3[k])));
                        r[e] = mx + shv[e];
                    }
                    pk[2 * k4] = pack_f16x2_relu(r[0], r[1]);            // ReLU + fp16 (the inference activation format)
                    pk[2 * k4 + 1] = pack_f16x2_relu(r[2], r[3]);
                }
                *reinterpret_cast<uint4*>(obase + (int64_t)(2 * hf) * (S * S) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(obase + (int64_t)(2 * hf + 1) * (S * S) * 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
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
