// Scene-dense layers on tcgen05 — the tensor-core kernels of the whole-scene ("dense") inference path.
//
// Solver.color() / test() classify EVERY pixel of a scene (solver/mainsolver.py:167-185): patches at stride 1.  The
// value a layer produces at patch-relative position (i, j) of the patch anchored at (x, y) depends only on the
// ABSOLUTE position (x + i, y + j) and on how close (i, j) is to the patch border (the convs zero-pad at the patch
// border, not at the scene border): a 3x3/pad-1 conv output has 3 border variants per axis {first, interior, last},
// and a conv + 2x2 max-pool on such an input again has 3.  So every layer is 9 scene-level maps ("planes"), each
// computed ONCE per scene position instead of once per patch.  Two kernels:
//   fuse_rowsum_kernel      the 1x1 fusion conv fused with the row sums of the global average pool
//   conv_pool4_kernel       conv3x3 + BN + ReLU + 2x2 max-pool, 9 pooled planes out of 9 (x 4 phases) input planes
// Both use the C8-planar no-swizzle operand layout and the warp roles of conv_tc.cuh.
#pragma once
#include "conv_tc.cuh"

namespace dmf {
namespace tc {

// ------------------------------------------------------------------------------------------------
// 1x1 fusion conv + BN + ReLU fused with the inner sums of the global average pool: F never goes to HBM.
//
//     S[a][X][y] = sum_{l < P2} F[a, cls(l)][X][y + 2l],    F[a, b] = fp16(relu(bn(W . CAT[a, b])))    (cls = first / interior / last)
//
// One tile = one map row X of one row class a, 128 consecutive columns: three accumulators (column classes b = 0, 1, 2) of
// M128 x N128, K = 256 in 3 x 4 pipeline steps of 8 channel chunks.  128 columns of one row of one chunk plane are 2 KB contiguous
// in HBM, so a step is 8 plain bulk copies (cp.async.bulk) that land directly in the K-major no-swizzle layout (a 5-D tensor-map
// box with a 16-byte inner dimension did the same at a third of the rate).  Epilogue, 8 warps: (1) thread = (column, channel half):
// tcgen05.ld, affine, ReLU, bf16, into a shared-memory F tile [b][col][128 ch] (row pitch 272 B: conflict-free 16-byte
// accesses); (2) thread = (anchor column y, channel half): the P2 strided samples are added in fp32, l ascending, S is written.  A tile yields 128 - 2 (P2 - 1)
// anchors (the tiles overlap by the window width).  Same rounding points and summation order as a separate F tensor + row-sum
// pass: the results are bit-identical to that formulation.
struct FuseRowsParams {
    int rows, W;                          // map rows of this band, anchor columns
    int tiles_x, n_tiles;
    int s_rows;                           // row dimension of S ([3][16][s_rows][W][8] fp16)
    int row_lo[3], row_n[3];              // map rows [lo, lo + n) of each row class that some anchor of the band uses
    int R1, C1;                           // CAT plane geometry: [9][32 chunks][R1][C1][8] bf16 (+ 2 KB of slack behind the tensor)
    const __nv_bfloat16* cat;
    int dbg;
    const __nv_bfloat16* w;               // packed [C_in/8][C_out][8]
    const float* scale;
    const float* shift;
    uint4* S;                              // [3][16][s_rows][W][8] fp16: row means of F (l ascending, fp32 accumulation, one fp16 rounding)
};

constexpr int kFrKQ = 8, kFrStages = 7, kFrPitch = 272;      // 7 x 16 KB of CAT rows in flight per SM (the kernel is bound by memory-level parallelism)
constexpr uint32_t kFrStage = kFrKQ * 128 * 16;                                        // 16 KB
constexpr size_t kFrSmem = (size_t)256 * 128 * 2 + kFrStages * kFrStage + 128 * kFrPitch + 2 * 128 * 4 + 26 * 8;

template <int P2>
__global__ void __launch_bounds__(320, 1) fuse_rowsum_kernel(const __grid_constant__ FuseRowsParams P) {
    constexpr int C_IN = 256, C_OUT = 128, KCH = C_IN / 8, NSTEP = KCH / kFrKQ;
    constexpr int VALID = 128 - 2 * (P2 - 1);
    constexpr uint32_t WBYTES = (uint32_t)C_IN * C_OUT * 2;
    constexpr uint32_t A_PLANE = 128u * 16u;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + WBYTES;
    uint8_t* f_s = a_s + kFrStages * kFrStage;                       // [128 cols][272 B]: the F tile of ONE column class
    float* scale_s = reinterpret_cast<float*>(f_s + 128 * kFrPitch);
    float* shift_s = scale_s + C_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(shift_s + C_OUT);
    // bars: [0,S) full, [S,2S) empty, 2S weights, 2S+1.. tmem_full[4], 2S+5.. tmem_empty[4]   (S = kFrStages)
    // TMEM is a ring of four 128-column accumulators, one per (tile, column class): the MMA warp runs up to four classes ahead of the
    // epilogue, which hands every accumulator back as soon as it has read it (with one barrier pair per TILE the MMA warp waited for the
    // drain a third of its time and the epilogue warps for the MMAs 39 % of theirs: ncu source view, round 2).
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kFrStages + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kFrStages + s); };
    const uint32_t w_bar = bar0 + 8u * (2 * kFrStages);
    auto tfull_bar = [&](uint32_t slot) { return bar0 + 8u * (2 * kFrStages + 1 + slot); };
    auto tempty_bar = [&](uint32_t slot) { return bar0 + 8u * (2 * kFrStages + 5 + slot); };

    for (int i = threadIdx.x; i < C_OUT; i += 320) {
        scale_s[i] = P.scale[i];
        shift_s[i] = P.shift[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kFrStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(w_bar, 1);
        for (uint32_t k = 0; k < 4; ++k) { mbar_init(tfull_bar(k), 1); mbar_init(tempty_bar(k), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // tile -> (map row X, row class a, column tile tx), walked incrementally by every role
    int tx = (int)blockIdx.x % P.tiles_x, a = ((int)blockIdx.x / P.tiles_x) % 3, X = ((int)blockIdx.x / P.tiles_x) / 3;
    auto next_tile = [&]() {
        tx += (int)gridDim.x;
        while (tx >= P.tiles_x) {
            tx -= P.tiles_x;
            if (++a == 3) { a = 0; ++X; }
        }
    };
    auto live = [&]() { return X >= P.row_lo[a] && X < P.row_lo[a] + P.row_n[a]; };    // rows no anchor uses are skipped by every role alike

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(w_bar, WBYTES);
            constexpr uint32_t CH = 16384;
            for (uint32_t off = 0; off < WBYTES; off += CH)
                bulk_load(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(P.w) + off, CH, w_bar);
        }
        __syncwarp();
        int st = 0;
        uint32_t ph = 1;
        for (int i = 0; i < n_local; ++i, next_tile()) {
            if (!live()) continue;
            for (int step = 0; step < 3 * NSTEP; ++step) {
                mbar_wait(empty_bar(st), ph);
                if (leader) {
                    if (P.dbg & 1) {
                        mbar_arrive(full_bar(st));
                    } else {
                        mbar_expect_tx(full_bar(st), kFrStage);
                        const int64_t chunk0 = (int64_t)(a * 3 + step / NSTEP) * KCH + (step % NSTEP) * kFrKQ;
                        const __nv_bfloat16* src = P.cat + ((chunk0 * P.R1 + X) * P.C1 + tx * VALID) * 8;
                        const int64_t cstride = (int64_t)P.R1 * P.C1 * 8;
                        for (int ck = 0; ck < kFrKQ; ++ck)
                            bulk_load(smem_u32(a_s) + (uint32_t)st * kFrStage + (uint32_t)ck * A_PLANE, src + ck * cstride, A_PLANE, full_bar(st));
                    }
                }
                __syncwarp();
                if (++st == kFrStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(128, C_OUT);          // fp16 activations and weights
        const bool leader = elect_one();
        mbar_wait(w_bar, 0);
        const uint64_t w_desc0 = umma_desc(smem_u32(w_s), C_OUT * 16, 128);
        int st = 0;
        uint32_t ph = 0;
        uint32_t unit = 0;                             // (tile, class) accumulators started
        for (int i = 0; i < n_local; ++i, next_tile()) {
            if (!live()) continue;
            for (int step = 0; step < 3 * NSTEP; ++step) {
                const int kq = step % NSTEP;
                const uint32_t slot = unit & 3u;
                if (kq == 0) mbar_wait(tempty_bar(slot), ((unit >> 2) & 1u) ^ 1u);
                mbar_wait(full_bar(st), ph);
                tc_fence_after();
                const uint64_t a_desc0 = umma_desc(smem_u32(a_s) + (uint32_t)st * kFrStage, A_PLANE, 128);
                if (leader) {
#pragma unroll
                    for (int j = 0; j < kFrKQ / 2; ++j) {
                        const uint64_t ad = a_desc0 + (uint64_t)(((uint32_t)(2 * j) * A_PLANE) >> 4);
                        const uint64_t bd = w_desc0 + (uint64_t)((uint32_t)((kq * kFrKQ + 2 * j) * C_OUT * 16) >> 4);
                        umma_bf16(tmem_base + slot * (uint32_t)C_OUT, ad, bd, idesc, (kq | j) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(st));
                    if (kq == NSTEP - 1) umma_commit(tfull_bar(slot));
                }
                __syncwarp();
                if (kq == NSTEP - 1) ++unit;
                if (++st == kFrStages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------ epilogue (8 warps: TMEM lane quarter q, channel half hc)
        const int q = warp & 3, hc = (warp - 2) >> 2;
        const int m = q * 32 + lane;                      // phase 1: column of the tile; phase 2: anchor column
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t unit = 0;                             // (tile, class) accumulators drained
        for (int i = 0; i < n_local; ++i, next_tile()) {
            if (!live()) continue;
            // One F tile (one column class) in shared memory at a time; the row sums live in 64 fp32 registers per thread
            // (thread = anchor column m, channel half hc).  Class 0 contributes F[a,0][X][y] = the thread's OWN TMEM lane: no
            // exchange; class 1 contributes the columns y + 2l, l = 1 .. P2-2, class 2 the column y + 2(P2-1): through the tile.
            // Same fp16 rounding points and the same fp32 summation order (l ascending) as a materialised F tensor.
            float acc[64];
            const bool act = !(P.dbg & 2);
            // next accumulator of the ring: wait for its MMAs, read it (when `act`), hand it back
            auto drain = [&](bool to_regs) {
                const uint32_t slot = unit & 3u;
                mbar_wait(tfull_bar(slot), (unit >> 2) & 1u);
                ++unit;
                tc_fence_after();
                if (act) {
#pragma unroll
                for (int cg = 0; cg < 2; ++cg) {
                    const int c0 = hc * 64 + cg * 32;
                    uint32_t v[32];
                    tmem_ld32(t_row + slot * (uint32_t)C_OUT + (uint32_t)c0, v);
                    const float4* sc4 = reinterpret_cast<const float4*>(scale_s + c0);
                    const float4* sh4 = reinterpret_cast<const float4*>(shift_s + c0);
                    uint32_t pk[16];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 sc = sc4[k], sh = sh4[k];
                        const float a0 = fmaf(__uint_as_float(v[4 * k]), sc.x, sh.x);
                        const float a1 = fmaf(__uint_as_float(v[4 * k + 1]), sc.y, sh.y);
                        const float a2 = fmaf(__uint_as_float(v[4 * k + 2]), sc.z, sh.z);
                        const float a3 = fmaf(__uint_as_float(v[4 * k + 3]), sc.w, sh.w);
                        pk[2 * k] = pack_f16x2_relu(a0, a1);
                        pk[2 * k + 1] = pack_f16x2_relu(a2, a3);
                    }
                    if (to_regs) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pk[k]));
                            acc[cg * 32 + 2 * k] = f.x;
                            acc[cg * 32 + 2 * k + 1] = f.y;
                        }
                    } else {
                        uint4* dst = reinterpret_cast<uint4*>(f_s + (uint32_t)m * kFrPitch + (uint32_t)c0 * 2);
#pragma unroll
                        for (int s4 = 0; s4 < 4; ++s4) dst[s4] = make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]);
                    }
                }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(slot));
            };
            auto add_col = [&](int col) {                              // acc += F tile row `col`, this thread's 64 channels
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint4 v = *reinterpret_cast<const uint4*>(f_s + (uint32_t)col * kFrPitch + (uint32_t)(hc * 8 + ch) * 16);
                    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[h]));
                        acc[ch * 8 + 2 * h] += f.x;
                        acc[ch * 8 + 2 * h + 1] += f.y;
                    }
                }
            };
            const int y = tx * VALID + m;
            const bool mine = m < VALID && y < P.W && act;
            drain(true);                                                  // class 0: l = 0
            drain(false);                                                 // class 1
            asm volatile("bar.sync 1, 256;" ::: "memory");                // class-1 tile complete
            if (mine) {
#pragma unroll
                for (int l = 1; l < P2 - 1; ++l) add_col(m + 2 * l);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");                // class-1 tile consumed
            drain(false);                                                 // class 2
            asm volatile("bar.sync 1, 256;" ::: "memory");                // class-2 tile complete
            if (mine) {
                add_col(m + 2 * (P2 - 1));
                uint4* o = P.S + (((int64_t)a * 16 + hc * 8) * P.s_rows + X) * P.W + y;
                const int64_t cstride = (int64_t)P.s_rows * P.W;
                constexpr float inv = 1.0f / (float)P2;                    // S holds row MEANS (an exact power-of-two scaling: no fp16 overflow of the sums)
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    o[ch * cstride] = make_uint4(pack_f16x2(acc[ch * 8] * inv, acc[ch * 8 + 1] * inv), pack_f16x2(acc[ch * 8 + 2] * inv, acc[ch * 8 + 3] * inv),
                                                 pack_f16x2(acc[ch * 8 + 4] * inv, acc[ch * 8 + 5] * inv), pack_f16x2(acc[ch * 8 + 6] * inv, acc[ch * 8 + 7] * inv));
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");                // F tile free again
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Fused conv3x3 + 2x2 max-pool: the pooled maps are the only thing written.
//
// The pooled output cell (X, Y) of border class (a, b) is the max of 4 conv outputs ("sub-positions" (s, t)); conv output
// (s, t) reads inputs at offsets o = s + dy, t + dx in {-1, 0, 1, 2} from the cell origin, and the border variant of an
// input depends on its offset only (first cell: {outside, first, interior, interior}; interior: all interior; last cell:
// {interior, interior, last, outside}).
//   * ALIGNED pooling (pan2, on the pooled-once PAN grid, which moves 2 cells per pixel): the input maps are stored
//     PHASE-SEPARATED ([variant][row phase, col phase][chunk][R][C][8]); offset o lives in phase o & 1 at cell shift
//     floor(o / 2), so every (sub-position, tap) pair is a unit-stride 16 x 8 window of some (variant, phase) plane.
//   * STRIDE-1 pooling (ms2, pan3: the pooling windows of neighbouring anchors overlap): offset o is simply cell shift o.
// A "source" of an axis = a distinct (variant, phase); a TMA box = (row source) x (column source), BR x BC cells, placed at
// the smallest shift any of its users needs.  One pooled tile = 16 x 8 cells = 4 accumulators of N = C_OUT in TMEM and
// (C_IN / 8) / KQ pipeline steps of KQ channel chunks (all boxes of a step land on one mbarrier); MMAs are issued
// tap-major so that consecutive instructions hit different accumulators (no accumulate read-after-write stall).
// NBUF = 2 (4 x C_OUT <= 256 columns): two tiles in flight, two epilogue groups of 4 warps.  NBUF = 1 (C_OUT = 128: the
// 4 accumulators fill TMEM): 8 epilogue warps split the channels, the next tile's MMAs wait for them.
// Epilogue: thread-local maximum of the 4 raw accumulators (the weights carry sign(BN scale)), then |scale| * max + shift, ReLU,
// bf16, 16-byte stores in the C8-planar layout.
//
// SHARE (stride-1 pooling only).  With stride-1 pooling the sub-position s = 1 of cell X is the conv output at position X + 1.
// On an axis whose border class is INTERIOR that is exactly the s = 0 output of the NEXT cell (same input variants, same
// taps): evaluating both sub-positions in every cell computes each conv output twice (four times in the interior / interior
// class).  For such an axis the tile evaluates ONE sub-position per cell (cls.ns / cls.nt = 1) and the epilogue takes the
// maximum with the neighbouring cell's value instead: columns by a lane shuffle (lane + 1), rows by a lane shuffle (lane + 8)
// and, across the 4-row lane quarters, through a 1 KB shared-memory hand-over between the epilogue warps of a channel slice.
// The max is taken on the packed fp16 values AFTER the affine + ReLU + rounding, which is exact because all three are monotone
// (the weights carry sign(scale)).  A tile's last row / column then only feeds its neighbour, so such tiles advance by 15 rows /
// 7 columns.  Tap evaluations per position over the 9 classes: (5 + 3 + 5)^2 = 169 instead of (5 + 6 + 5)^2 = 256.
constexpr int kP4MaxBoxes = 9;

// tcgen05.mma with a collector hint for the A operand: USAGE 0 = none, 1 = fill (keep A for the next instruction), 2 = use (A is
// the one kept; keep it), 3 = lastuse.  SASS: UTCHMMA gdesc[..].A_KEEP / .A_REUSE.A_KEEP / .A_REUSE.
template <int USAGE>
__device__ __forceinline__ void umma_bf16_a(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (USAGE == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if (USAGE == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if (USAGE == 3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
}

struct Pool4Cls {
    int16_t out_plane, n_boxes;
    int16_t box_plane[kP4MaxBoxes];
    int8_t box_drow[kP4MaxBoxes], box_dcol[kP4MaxBoxes];
    int16_t win[16];                      // [row offset + 1][col offset + 1]: (byte offset of that A window inside a stage) >> 4, or -1 = outside the patch
    int16_t ns, nt;                       // sub-position rows / columns the tile evaluates itself: 2, or 1 = shared with the neighbouring cell (see below)
    int16_t ca, cb;                       // row / column border class (the table is walked in an order that keeps equal tile sizes together)
};

struct Pool4Params {
    int rows, cols;                       // grid of the pooled output = grid of every input plane
    int tiles_x, tiles_y, n_tiles;
    int trow[3], tcol[3];                 // tile pitch per border class: 16 rows / 8 columns, or 15 / 7 where the tile's last row / column only feeds its neighbour
    int out_chunks, out_chunk0;
    int dbg;
    // rows / columns of each border class (first, interior, last) that some anchor of the band actually uses: [lo, lo + n)
    int row_lo[3], row_n[3], col_lo[3], col_n[3];
    const __nv_bfloat16* w;               // packed [C_in/8][tap][C_out][8], output channel co multiplied by sign(BN scale[co])
    const float* scale;                   // |BN scale| (its sign is folded into w)
    const float* shift;
    __nv_bfloat16* out;                   // [9][out_chunks][rows][cols][8]
    Pool4Cls cls[9];
};

template <int C_IN, int C_OUT, int KQ, int STAGES, int BR, int BC, int NBUF, bool SHARE = false>
struct Pool4Cfg {
    static constexpr int KCH = C_IN / 8, NSTEP = KCH / KQ;
    static constexpr uint32_t PLANE = BR * BC * 16;                          // bytes of one channel-chunk plane of a box
    static constexpr uint32_t BOX_BYTES = KQ * PLANE;
    static constexpr uint32_t BOX_SLOT = (BOX_BYTES + 127) / 128 * 128;      // TMA destinations are 128-byte aligned
    static constexpr int MAX_BOXES = BC == 9 ? 9 : 4;                        // aligned: 3 x 3 sources, stride-1: 2 x 2
    static constexpr uint32_t STAGE = MAX_BOXES * BOX_SLOT;
    static constexpr uint32_t WBYTES = 9u * C_IN * C_OUT * 2;
    static constexpr uint32_t CLS_BYTES = (9 * sizeof(Pool4Cls) + 15) / 16 * 16;
    static constexpr uint32_t XCHG = SHARE ? 2u * 64u * C_OUT : 0u;         // 2 buffers x 8 lanes x C_OUT channels x fp16 x 4 lane quarters
    static constexpr int GRAN = SHARE ? 1 : MAX_BOXES;                       // boxes per slot of the A ring: SHARE allocates it box by box (see conv_pool4_kernel)
    static constexpr int NSLOT = STAGES * MAX_BOXES / GRAN;
    static constexpr size_t SMEM = WBYTES + (size_t)STAGES * STAGE + 2 * C_OUT * 4 + (2 * NSLOT + 12) * 8 + CLS_BYTES + XCHG;
    static_assert(NSLOT <= 32, "slot phase bits live in one 32-bit word");
    static_assert(KCH % KQ == 0 && KQ % 2 == 0 && NBUF * 4 * C_OUT <= 512 && WBYTES % 128 == 0 && (!SHARE || NBUF == 1), "pool4 configuration");
};


// The MMAs of one pipeline step of one tile, for NS x NT sub-positions (see conv_pool4_kernel).  Window-major issue order with PAIRED
// column sub-positions.  The A window at offset (orow, ocol) from the cell origin feeds every (sub-position (s, t), tap (dy, dx)) with
// s + dy = orow, t + dx = ocol.  NT = 2: for ocol in {0, 1} both column sub-positions take part, with the horizontally adjacent taps
// dx = ocol - 1 (t = 1) and dx = ocol (t = 0): ONE instruction with N = 2 * C_out whose B operand spans the two tap blocks and whose
// accumulator spans the column blocks of (s, 1) and (s, 0) — TMEM block of (s, t) = 2 s + (1 - t) (NT = 1: block s).  Twice the math per A fetch and per
// instruction issue (an M128 x N128 x K16 instruction is bound by its shared-memory operand reads, ~94 clk against a 64-clk math
// floor; at N = 256 the math takes 128 clk and hides them).  ocol = -1 / 2 feed one column sub-position each (N = C_out).  Within a
// row offset the windows ocol = 0, 1 are issued first, so that the first write of a tile into a block (both halves of a pair) is an
// overwriting instruction.  Instructions that share an A window are issued back to back with A kept in the collector (UTCHMMA
// ...A_KEEP / A_REUSE).  NT = 1: t = 0 only, every window feeds one N = C_out instruction; NS = 1: s = 0 only.
template <int C_OUT, int KQ, uint32_t PLANE, int NS, int NT>
__device__ __forceinline__ void pool4_issue_step(const int (&win)[16], uint64_t a_desc0, uint64_t w_desc0, uint32_t d_tmem, int kq, uint32_t& started) {
    constexpr uint32_t idesc1 = umma_idesc_f16(128, C_OUT), idesc2 = umma_idesc_f16(128, 2 * C_OUT);      // fp16 activations and weights
#pragma unroll
    for (int orow = -1; orow <= NS; ++orow) {
        int nu = 0;                                    // sub-position rows with a tap dy = orow - s
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) nu += (orow - s2 >= -1 && orow - s2 <= 1) ? 1 : 0;
#pragma unroll
        for (int oi = 0; oi < 2 + NT; ++oi) {
            constexpr int kOcol[4] = {0, 1, -1, 2};
            const int ocol = kOcol[oi];
            const int o = win[(orow + 1) * 4 + ocol + 1];
            if (o >= 0) {
#pragma unroll
                for (int j = 0; j < KQ / 2; ++j) {
                    const uint64_t ad = a_desc0 + (uint64_t)(uint32_t)o + (uint64_t)(((uint32_t)(2 * j) * PLANE) >> 4);
                    int ui = 0;
#pragma unroll
                    for (int s2 = 0; s2 < NS; ++s2) {
                        if (orow - s2 >= -1 && orow - s2 <= 1) {
                            const int dy = orow - s2;
                            const bool lead = ocol == 0 || ocol == 1;                        // the windows that open a block
                            const bool pair = NT == 2 && lead;
                            const int t2 = ocol == 2 ? 1 : 0;                                // single: the one column sub-position
                            const int tap = pair ? (dy + 1) * 3 + ocol : (dy + 1) * 3 + (ocol - t2) + 1;
                            const uint64_t bd = w_desc0 + (uint64_t)((uint32_t)(((kq * KQ + 2 * j) * 9 + tap) * C_OUT * 16) >> 4);
                            const uint32_t dt = d_tmem + (uint32_t)((pair ? 2 * s2 : NT * s2 + (NT == 2 ? 1 - t2 : 0)) * C_OUT);   // block of (s, t) = NT s + (NT - 1)(1 - t)
                            const uint32_t accf = lead ? (((started >> s2) & 1u) | (uint32_t)j) : 1u;
                            const uint32_t id = pair ? idesc2 : idesc1;
                            if (nu == 1) umma_bf16_a<0>(dt, ad, bd, id, accf);
                            else if (ui == 0) umma_bf16_a<1>(dt, ad, bd, id, accf);
                            else umma_bf16_a<3>(dt, ad, bd, id, accf);
                            ++ui;
                        }
                    }
                }
                if (ocol == 0 || ocol == 1) {
#pragma unroll
                    for (int s2 = 0; s2 < NS; ++s2)
                        if (orow - s2 >= -1 && orow - s2 <= 1) started |= 1u << s2;
                }
            }
        }
    }
}

template <int C_IN, int C_OUT, int KQ, int STAGES, int BR, int BC, int NBUF, int EW, bool SHARE = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1) conv_pool4_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ Pool4Params P) {
    using Cfg = Pool4Cfg<C_IN, C_OUT, KQ, STAGES, BR, BC, NBUF, SHARE>;
    constexpr int kThreads = 64 + 32 * EW;             // EW epilogue warps: NBUF = 2 -> two groups of EW / 2; NBUF = 1 -> EW / 4 channel slices
    static_assert(EW % 4 == 0 && (NBUF == 1 || EW == 8), "epilogue warps");
    constexpr int KCH = Cfg::KCH, NSTEP = Cfg::NSTEP;
    constexpr uint32_t WBYTES = Cfg::WBYTES, SBO_A = BC * 16;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + WBYTES;
    float* scale_s = reinterpret_cast<float*>(a_s + (size_t)STAGES * Cfg::STAGE);
    float* shift_s = scale_s + C_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(shift_s + C_OUT);
    // SHARE: the A ring is a ring of NSLOT box slots, not of stages: a pipeline step takes as many consecutive slots as its class has
    // boxes (1 .. 4 of the 8; it restarts at slot 0 rather than wrap), so the classes with few input sources keep up to 8 steps in
    // flight instead of 2 — their tiles are short (9 instructions per step in the interior / interior class) and two steps did not
    // cover the load latency.  Per-slot mbarriers: `full` of a step's FIRST slot, `empty` of every slot it used.  Otherwise (pan2: one
    // step per tile, 4 .. 9 boxes) a slot is a whole stage: per-box barriers cost more there than the extra depth gives (4.05 -> 4.4 ms).
    // bars: [0,NS) full; [NS,2NS) empty; 2NS weights; 2NS+1.. tmem_full[4]; 2NS+5.. tmem_empty[4]   (NBUF = 2 uses two of each)
    constexpr int NSLOT = Cfg::NSLOT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSLOT + 10);
    Pool4Cls* cls_s = reinterpret_cast<Pool4Cls*>(bars + 2 * NSLOT + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (NSLOT + s); };
    const uint32_t w_bar = bar0 + 8u * (2 * NSLOT);
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NSLOT + 1 + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NSLOT + 5 + a); };
    auto take_boxes = [](int& head, int n) {           // first slot of a step with n boxes
        const int h = head + n > NSLOT ? 0 : head;
        head = h + n == NSLOT ? 0 : h + n;
        return h;
    };
    // NBUF = 1: TMEM is a ring of 4 slots of C_OUT columns; a tile takes ns * nt (1, 2 or 4) slots, aligned to its size, so the
    // tiles that share sub-positions with their neighbours leave room for the next tile's MMAs while they are drained.
    constexpr int kSlots = NBUF == 1 ? 4 : NBUF;
    auto take_slots = [](int& ring, int n) {
        int b = (ring + n - 1) & ~(n - 1);
        if (b + n > 4) b = 0;
        ring = (b + n) & 3;
        return b;
    };

    for (int i = threadIdx.x; i < C_OUT; i += kThreads) {
        scale_s[i] = P.scale[i];
        shift_s[i] = P.shift[i];
    }
    for (int i = threadIdx.x; i < 9 * (int)(sizeof(Pool4Cls) / 4); i += kThreads)
        reinterpret_cast<uint32_t*>(cls_s)[i] = reinterpret_cast<const uint32_t*>(P.cls)[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(w_bar, 1);
        for (int a = 0; a < kSlots; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NBUF == 2 ? 4 : EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // tile -> (row strip ty, class c, column tile tx), walked incrementally by every role
    int tx = (int)blockIdx.x % P.tiles_x, c = ((int)blockIdx.x / P.tiles_x) % 9, ty = ((int)blockIdx.x / P.tiles_x) / 9;
    auto next_tile = [&]() {
        tx += (int)gridDim.x;
        while (tx >= P.tiles_x) {
            tx -= P.tiles_x;
            if (++c == 9) { c = 0; ++ty; }
        }
    };
    // tiles outside the rows / columns that any anchor of the band uses for this border class are skipped by every role alike
    auto live = [&]() {
        const int a = cls_s[c].ca, b = cls_s[c].cb;
        return ty * P.trow[a] < P.row_n[a] && tx * P.tcol[b] < P.col_n[b];
    };

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: all boxes of a step land on one barrier
        const bool leader = elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&in_map) : "memory");
            mbar_expect_tx(w_bar, WBYTES);
            constexpr uint32_t CH = 12288;
            for (uint32_t off = 0; off < WBYTES; off += CH)
                bulk_load(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(P.w) + off, min(CH, WBYTES - off), w_bar);
        }
        __syncwarp();
        int head = 0;                                  // next free box slot
        uint32_t eph = 0;                              // phase bit of every slot's empty barrier
        for (int i = 0; i < n_local; ++i, next_tile()) {
            if (!live()) continue;
            const int nbx = cls_s[c].n_boxes;
            const int ca = cls_s[c].ca, cb = cls_s[c].cb;
            const int row0 = P.row_lo[ca] + ty * P.trow[ca], col0 = P.col_lo[cb] + tx * P.tcol[cb];
            const int nsl = SHARE ? nbx : 1;               // ring slots of a step
            for (int kq = 0; kq < NSTEP; ++kq) {
                const int h = take_boxes(head, nsl);
                for (int k = h; k < h + nsl; ++k) mbar_wait(empty_bar(k), ((eph >> k) & 1u) ^ 1u);
                eph ^= ((1u << nsl) - 1u) << h;
                if (leader) {
                    if (P.dbg & 1) {
                        mbar_arrive(full_bar(h));
                    } else {
                        mbar_expect_tx(full_bar(h), (uint32_t)nbx * Cfg::BOX_BYTES);
                        const uint32_t dst = smem_u32(a_s) + (uint32_t)h * (Cfg::GRAN * Cfg::BOX_SLOT);
                        for (int b = 0; b < nbx; ++b)
                            tma_load_4d(dst + (uint32_t)b * Cfg::BOX_SLOT, &in_map, full_bar(h), (col0 + cls_s[c].box_dcol[b]) * 8,
                                        cls_s[c].box_plane[b], row0 + cls_s[c].box_drow[b], kq * KQ);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        // The packed weights are [C_in/8][tap][C_out][8]: for one channel chunk the 9 taps are consecutive C_out-row blocks, so a
        // B descriptor that starts at tap t and spans N = 2 * C_out rows covers taps t and t + 1 = (dy, dx) and (dy, dx + 1).
        const bool leader = elect_one();
        mbar_wait(w_bar, 0);
        const uint64_t w_desc0 = umma_desc(smem_u32(w_s), 9 * C_OUT * 16, 128);
        int head = 0;                                  // next box slot of the A ring
        uint32_t aph = 0;                              // phase bit of every slot's full barrier
        int it = 0;                                    // tiles actually processed
        int ring = 0;                                  // NBUF = 1: next free TMEM slot
        uint32_t eph = 0;                              // NBUF = 1: phase bit of every slot's tmem_empty barrier
        for (int i = 0; i < n_local; ++i, next_tile()) {
            if (!live()) continue;
            const int ns = SHARE ? cls_s[c].ns : 2, nt = SHARE ? cls_s[c].nt : 2;
            int buf;
            uint32_t d_tmem;
            if (NBUF == 2) {
                buf = it & 1;
                mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1);
                d_tmem = tmem_base + (uint32_t)(buf * 4 * C_OUT);
            } else {
                const int n = ns * nt;
                buf = take_slots(ring, n);
                for (int k = buf; k < buf + n; ++k) mbar_wait(tempty_bar(k), ((eph >> k) & 1u) ^ 1u);
                eph ^= ((1u << n) - 1u) << buf;
                d_tmem = tmem_base + (uint32_t)(buf * C_OUT);
            }
            ++it;
            int win[16];                                   // window table of the class in registers: no shared-memory load on the issue path
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t w2 = reinterpret_cast<const uint32_t*>(cls_s[c].win)[k];
                win[2 * k] = (int)(int16_t)(w2 & 0xffffu);
                win[2 * k + 1] = (int)(int16_t)(w2 >> 16);
            }
            uint32_t started = 0;                          // bit s: the accumulator blocks of sub-position row s have been written
            const int nsl = SHARE ? cls_s[c].n_boxes : 1;  // ring slots of a step
            for (int kq = 0; kq < NSTEP; ++kq) {
                const int h = take_boxes(head, nsl);
                mbar_wait(full_bar(h), (aph >> h) & 1u);
                aph ^= 1u << h;
                tc_fence_after();
                const uint64_t a_desc0 = umma_desc(smem_u32(a_s) + (uint32_t)h * (Cfg::GRAN * Cfg::BOX_SLOT), Cfg::PLANE, SBO_A);
                if (leader) {
                    if (!SHARE || (ns == 2 && nt == 2)) pool4_issue_step<C_OUT, KQ, Cfg::PLANE, 2, 2>(win, a_desc0, w_desc0, d_tmem, kq, started);
                    else if (ns == 2) pool4_issue_step<C_OUT, KQ, Cfg::PLANE, 2, 1>(win, a_desc0, w_desc0, d_tmem, kq, started);
                    else if (nt == 2) pool4_issue_step<C_OUT, KQ, Cfg::PLANE, 1, 2>(win, a_desc0, w_desc0, d_tmem, kq, started);
                    else pool4_issue_step<C_OUT, KQ, Cfg::PLANE, 1, 1>(win, a_desc0, w_desc0, d_tmem, kq, started);
                    for (int k = h; k < h + nsl; ++k) umma_commit(empty_bar(k));
                    if (kq == NSTEP - 1) umma_commit(tfull_bar(buf));
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------ epilogue: BN + ReLU on the 4 sub-position accumulators, thread-local max
        const int eg = (warp - 2) >> 2;                    // NBUF = 2: tile parity this group drains; NBUF = 1: channel slice
        const int q = warp & 3;
        const int m = q * 32 + lane;
        constexpr int C_SPAN = NBUF == 2 ? C_OUT : C_OUT / (EW / 4);
        const int c_lo = NBUF == 2 ? 0 : eg * C_SPAN;
        const uint32_t t_row0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(NBUF == 2 ? eg * 4 * C_OUT : 0);
        const int64_t cstride = (int64_t)P.rows * P.cols * 8;
        constexpr int NG = C_SPAN / 32;                      // 32-channel groups per thread
        // SHARE: hand-over of the first cell row of every lane quarter to the quarter above it: [buffer][channel slice][quarter][group][lane 0..7][4] x 16 B
        uint4* const xchg = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(cls_s) + Cfg::CLS_BYTES);
        int it = -1;                                   // tiles actually processed
        int it_x = 0;                                  // tiles that used the hand-over
        int ring = 0;                                  // NBUF = 1: next free TMEM slot
        uint32_t fph = 0;                              // NBUF = 1: phase bit of every slot's tmem_full barrier
        for (int i = 0; i < n_local; ++i, next_tile()) {
            if (!live()) continue;
            ++it;
            if (NBUF == 2 && (it & 1) != eg) continue;
            const int ns = SHARE ? cls_s[c].ns : 2, nt = SHARE ? cls_s[c].nt : 2;
            const int nblk = ns * nt;                  // accumulator blocks of this tile, consecutive in TMEM
            const int buf = NBUF == 2 ? eg : take_slots(ring, nblk);
            const uint32_t t_row = t_row0 + (uint32_t)(NBUF == 2 ? 0 : buf * C_OUT);
            const int ca = cls_s[c].ca, cb = cls_s[c].cb;
            const int row = P.row_lo[ca] + ty * P.trow[ca] + (m >> 3), col = P.col_lo[cb] + tx * P.tcol[cb] + (m & 7);
            // a tile that shares a sub-position with its neighbours produces 15 rows / 7 columns: its last row / column only feeds them
            const bool valid = row < P.rows && col < P.cols && (!SHARE || ((ns == 2 || (m >> 3) < 15) && (nt == 2 || (m & 7) < 7)));
            __nv_bfloat16* const obase =
                P.out + ((((int64_t)cls_s[c].out_plane * P.out_chunks + P.out_chunk0) * P.rows + row) * P.cols + col) * 8;
            if (NBUF == 2) {
                mbar_wait(tfull_bar(buf), (it >> 1) & 1);
            } else {
                mbar_wait(tfull_bar(buf), (fph >> buf) & 1u);
                fph ^= 1u << buf;
            }
            tc_fence_after();
            // The packed weights carry sign(BN scale) and P.scale holds |scale|, so max_s relu(scale * z_s + shift) =
            // relu(|scale| * max_s z'_s + shift): one FMNMX per accumulator value, the affine + ReLU + fp16 rounding once.
            // Phase 1 reduces the accumulator blocks of the tile's own sub-positions into registers and hands their TMEM slots back
            // to the MMA warp; phase 2 (affine, ReLU, fp16, neighbour maxima, stores) then runs under the next tiles' MMAs.
            uint32_t mx[NG][32];
            if (!(P.dbg & 2)) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int c0 = c_lo + 32 * g;
                    uint32_t m1[32];
                    auto fold = [&](int blk) {
                        tmem_ld32_nowait(t_row + (uint32_t)(blk * C_OUT + c0), m1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int k = 0; k < 32; ++k) mx[g][k] = __float_as_uint(fmaxf(__uint_as_float(mx[g][k]), __uint_as_float(m1[k])));
                    };
                    tmem_ld32_nowait(t_row + (uint32_t)c0, mx[g]);
                    if (!SHARE || nblk >= 2) {
                        fold(1);
                    } else {
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                    if (!SHARE || nblk == 4) {
                        fold(2);
                        fold(3);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (NBUF == 2) {
                    mbar_arrive(tempty_bar(buf));
                } else {
                    for (int k = buf; k < buf + nblk; ++k) mbar_arrive(tempty_bar(k));
                }
            }
            if (!(P.dbg & 2)) {
                uint32_t pk[NG][16];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int c0 = c_lo + 32 * g;
                    const float4* sc4 = reinterpret_cast<const float4*>(scale_s + c0);
                    const float4* sh4 = reinterpret_cast<const float4*>(shift_s + c0);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 sc = sc4[k], sh = sh4[k];
                        const float a0 = fmaf(__uint_as_float(mx[g][4 * k]), sc.x, sh.x);
                        const float a1 = fmaf(__uint_as_float(mx[g][4 * k + 1]), sc.y, sh.y);
                        const float a2 = fmaf(__uint_as_float(mx[g][4 * k + 2]), sc.z, sh.z);
                        const float a3 = fmaf(__uint_as_float(mx[g][4 * k + 3]), sc.w, sh.w);
                        pk[g][2 * k] = pack_f16x2_relu(a0, a1);              // ReLU + saturating fp16 rounding: one F2FP
                        pk[g][2 * k + 1] = pack_f16x2_relu(a2, a3);
                    }
                }
                if (SHARE && nt == 1) {                                      // the next cell's column (lane + 1; column 7 has none: not stored)
#pragma unroll
                    for (int g = 0; g < NG; ++g)
#pragma unroll
                        for (int k = 0; k < 16; ++k) pk[g][k] = max_f16x2(pk[g][k], __shfl_down_sync(0xffffffffu, pk[g][k], 1));
                }
                if (SHARE && ns == 1) {                                      // the next cell's row: lane + 8, or the first row of the quarter above
                    uint4* const xw = xchg + (size_t)(((it_x & 1) * (EW / 4) + eg) * 4 + q) * (NG * 32);
                    ++it_x;
                    if (lane < 8) {
#pragma unroll
                        for (int g = 0; g < NG; ++g)
#pragma unroll
                            for (int s4 = 0; s4 < 4; ++s4)
                                xw[(g * 8 + lane) * 4 + s4] = make_uint4(pk[g][4 * s4], pk[g][4 * s4 + 1], pk[g][4 * s4 + 2], pk[g][4 * s4 + 3]);
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
                    const uint4* const xr = xw + NG * 32;                  // quarter q + 1 of the same channel slice
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        uint32_t up[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) up[k] = __shfl_down_sync(0xffffffffu, pk[g][k], 8);
                        if (lane >= 24 && q < 3) {
#pragma unroll
                            for (int s4 = 0; s4 < 4; ++s4) {
                                const uint4 v = xr[(g * 8 + lane - 24) * 4 + s4];
                                up[4 * s4] = v.x; up[4 * s4 + 1] = v.y; up[4 * s4 + 2] = v.z; up[4 * s4 + 3] = v.w;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 16; ++k) pk[g][k] = max_f16x2(pk[g][k], up[k]);
                    }
                }
                if (valid) {
#pragma unroll
                    for (int g = 0; g < NG; ++g)
#pragma unroll
                        for (int s4 = 0; s4 < 4; ++s4)
                            *reinterpret_cast<uint4*>(obase + (((c_lo + 32 * g) >> 3) + s4) * cstride) =
                                make_uint4(pk[g][4 * s4], pk[g][4 * s4 + 1], pk[g][4 * s4 + 2], pk[g][4 * s4 + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace tc
}  // namespace dmf
