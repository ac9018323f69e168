// Scene-dense whole-scene inference: GMFNet evaluated as scene-level maps instead of per-patch tensors.
//
// Replaces the two whole-scene loader passes of Solver.color() and the full-loader Solver.test()
// (solver/mainsolver.py:104-141, 167-185), where the patches of neighbouring pixels overlap in all but one row /
// column.  Patch (x, y) covers MS positions X = x + i, Y = y + j (i, j in [0, p)) and PAN pooled-once positions
// U = 2x + u, V = 2y + v (u, v in [0, 2p)).  A layer's value at a patch-relative position depends on the absolute
// position and on the position's border class only (see dense_tc.cuh), so per row band [r0, r1) of anchors:
//
//   scene --ms_stem_map--> A  [9 ][ 8][R][C][8]        9 = {first, interior, last}^2 of the 3x3/pad-1 stem conv
//   A  --conv_pool4 (stride-1 pool)--> CAT[9][0..15]     conv 64->128 + BN + ReLU + 2x2 max over (X..X+1, Y..Y+1), 9 pooled classes
//   scene --pan_stem_map--> B1 [9 ][4 phases][4][R][C][8]  stem conv + pool on the pooled-once grid, stored phase-separated
//   B1 --conv_pool4 (aligned pool)--> B2 [9][8][R][C][8]  conv 32->64 + aligned 2x2 max (the pooled-once grid moves 2 cells per pixel)
//   B2 --conv_pool4 (stride-1 pool)--> CAT[9][16..31]
//   CAT --fuse_rowsum (1x1 conv + row sums of the average pool)--> S [3][16][R][W][8] fp32;  F = relu(bn(W . CAT)) stays on chip
//   S  --head_dense--> per pixel: mean over the (p/2)^2 strided samples F[cls(k),cls(l)][x+2k][y+2l], 2 linears,
//                      argmax, confusion matrix, label map
// with R = (r1 - r0) + p - 1 rows and C = W + p - 1 columns.  Positions a band never needs hold don't-care values
// (they only ever feed other don't-care positions).  Results equal the per-patch path up to fp32 summation order.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "dense_tc.cuh"
#include "net_types.cuh"

namespace dmf {

struct DenseWs {
    int W = 0, band = 0, p = 0, R1 = 0, C1 = 0;
    __nv_bfloat16 *A = nullptr, *CAT = nullptr, *B1 = nullptr, *B2 = nullptr;
    __half* S = nullptr;                             // row means of F for the separable average pool [3][16][R][W][8], fp16
    float *w_ms1 = nullptr, *w_pan1 = nullptr;       // fp32 stem conv weights in torch layout
    // conv + pool layers (ms2, pan2, pan3): bf16 weights [C_in/8][tap][C_out][8] with sign(BN scale) folded into every output channel, |scale|,
    // shift — conv_pool4_kernel takes the max over the pooling window BEFORE the affine
    __nv_bfloat16* w_cp[3] = {nullptr, nullptr, nullptr};
    float *sc_cp[3] = {nullptr, nullptr, nullptr}, *sh_cp[3] = {nullptr, nullptr, nullptr};
    __nv_bfloat16* w_fc1 = nullptr;                  // fc1 weight as bf16 hi / lo parts [2][16][64][8] (B operand of the head's MMA)
    CUtensorMap mapA, mapB1, mapB2s;
    cudaEvent_t ev[12] = {};
    float stage_ms[12] = {};
    size_t bytes = 0;
};

// ---------------------------------------------------------------------------------------------- stem maps
// masked tap sums of one output position: t[k], k = (dy+1)*3 + (dx+1)  ->  s[rc*3 + cc]
// rc: 0 = first row of a patch (no dy = -1 taps), 1 = interior, 2 = last row (no dy = +1); cc likewise for dx.
__device__ __forceinline__ void masked_sums(const float (&t)[9], float (&s)[9]) {
    float u[3][3];                    // u[dy][cc]
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        u[d][0] = t[3 * d + 1] + t[3 * d + 2];
        u[d][1] = (t[3 * d] + t[3 * d + 1]) + t[3 * d + 2];
        u[d][2] = t[3 * d] + t[3 * d + 1];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        s[0 + c] = u[1][c] + u[2][c];
        s[3 + c] = (u[0][c] + u[1][c]) + u[2][c];
        s[6 + c] = u[0][c] + u[1][c];
    }
}

// MS stem (conv3x3 4->64 + BN + ReLU) at every position of the band: ms = padded scene [Hp][Wp] float4, band-local
// row Xl <-> scene row r0 + Xl.  w: fp32 [64][4][3][3].  out: A[9][8][R1][C1][8].  One thread = one position, ALL 8 channel chunks: the 3x3
// neighbourhood (9 float4 loads, bounds tests, index arithmetic: more than half of the instructions of a one-chunk thread) is fetched once.
__global__ void __launch_bounds__(256, 2) ms_stem_map_kernel(const float4* __restrict__ ms, int Hp, int Wp, int r0, int rows, int R1, int C1,
                                                          const float* __restrict__ w, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, __nv_bfloat16* __restrict__ A) {
    __shared__ float w_s[C_MS1][36], sc_s[C_MS1], sh_s[C_MS1];
    for (int i = threadIdx.x; i < C_MS1 * 36; i += blockDim.x) w_s[i / 36][i % 36] = w[i];
    if (threadIdx.x < C_MS1) { sc_s[threadIdx.x] = scale[threadIdx.x]; sh_s[threadIdx.x] = shift[threadIdx.x]; }
    __syncthreads();
    const int total = rows * C1;                       // <= 4111 x (W + 31): fits 32 bits
    const int64_t plane = (int64_t)R1 * C1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int Xl = i / C1, Y = i - Xl * C1;
        const int X = r0 + Xl;
        float4 nb[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int xx = X + k / 3 - 1, yy = Y + k % 3 - 1;
            nb[k] = (xx >= 0 && xx < Hp && yy >= 0 && yy < Wp) ? __ldg(ms + (int64_t)xx * Wp + yy) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint4* o = reinterpret_cast<uint4*>(A) + (int64_t)Xl * C1 + Y;
#pragma unroll 1
        for (int ch = 0; ch < C_MS1 / 8; ++ch, o += plane) {
            uint32_t out[9][4];
            float even[9];                   // affine results of the even channel of a pair: ReLU + fp16 rounding + packing is ONE F2FP per pair
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* wj = w_s[ch * 8 + j];          // [band][k]
                const float sc = sc_s[ch * 8 + j], sh = sh_s[ch * 8 + j];
                float t[9], s[9];
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    t[k] = fmaf(nb[k].w, wj[27 + k], fmaf(nb[k].z, wj[18 + k], fmaf(nb[k].y, wj[9 + k], nb[k].x * wj[k])));
                masked_sums(t, s);
#pragma unroll
                for (int v = 0; v < 9; ++v) {
                    const float y = fmaf(s[v], sc, sh);
                    if (j & 1) out[v][j >> 1] = tc::pack_f16x2_relu(even[v], y);
                    else even[v] = y;
                }
            }
#pragma unroll
            for (int v = 0; v < 9; ++v) o[(int64_t)v * 8 * plane] = make_uint4(out[v][0], out[v][1], out[v][2], out[v][3]);
        }
    }
}

// PAN stem (conv3x3 1->32 + BN + ReLU + maxpool2) on the pooled-once grid (2 cells per pixel and axis; the pooling grid is aligned to
// every patch origin because 4x is even).  One thread = the 2 x 2 pooled cells of one pixel (I, J) = all four (row, column) parity
// phases, one 6 x 6 window of PAN (fetched once), all 4 channel chunks in turn.  A cell on an even pooled row can only be the FIRST pooled row of a patch (u = 0) or an
// interior one, a cell on an odd row only the LAST (u = 2p-1) or interior; columns likewise: 2 x 2 of the 9 border variants exist per
// phase, the others are never read (conv_pool4_kernel's class table) and are neither computed nor stored.  "First" masks the dy = -1
// taps of the cell's upper conv row, "last" the dy = +1 taps of its lower one; everything is static per phase.  BN after the max: the
// weights carry sign(scale), sc_s holds |scale| (max commutes with a non-negative scale).
// w: fp32 [32][1][3][3].  out: B1[9][phase 4][4][R1][C1][8], PHASE-SEPARATED: pooled cell (U, V) is stored at
// (U >> 1, V >> 1) of phase plane (U & 1) * 2 + (V & 1).  i0 = first pixel row of the band, rows = pixel rows to produce.
__global__ void __launch_bounds__(256, 2) pan_stem_map_kernel(const float* __restrict__ pan, int H4p, int W4p, int pitch, int i0, int rows,
                                                           int R1, int C1, const float* __restrict__ w, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, __nv_bfloat16* __restrict__ B1) {
    __shared__ float w_all[C_PAN1][9], sc_all[C_PAN1], sh_all[C_PAN1];
    for (int i = threadIdx.x; i < C_PAN1 * 9; i += blockDim.x) w_all[i / 9][i % 9] = scale[i / 9] < 0.f ? -w[i] : w[i];
    if (threadIdx.x < C_PAN1) { sc_all[threadIdx.x] = fabsf(scale[threadIdx.x]); sh_all[threadIdx.x] = shift[threadIdx.x]; }
    __syncthreads();
    const int total = rows * C1;                       // fits 32 bits (see ms_stem_map_kernel)
    const int64_t plane = (int64_t)R1 * C1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int Il = i / C1, J = i - Il * C1;
        const int I = i0 + Il;
        float xw[6][6];                  // PAN rows 4I-1 .. 4I+4, cols 4J-1 .. 4J+4
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const int xx = 4 * I - 1 + a, yy = 4 * J - 1 + b;
                xw[a][b] = (xx >= 0 && xx < H4p && yy >= 0 && yy < W4p) ? __ldg(pan + (int64_t)xx * pitch + yy) : 0.f;
            }
#pragma unroll 1
        for (int ch = 0; ch < C_PAN1 / 8; ++ch) {
        const float (*w_s)[9] = w_all + ch * 8;
        const float *sc_s = sc_all + ch * 8, *sh_s = sh_all + ch * 8;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr)
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
                uint32_t out[4][4];      // [edge/interior row variant][edge/interior column variant] -> 8 packed channels
                float even[4];           // affine results of the even channel of a pair: ReLU + fp16 rounding + packing is ONE F2FP per pair
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // sv[a][b][re][ce]: conv position (a, b) of this pooled cell with its row (column) taps masked as on the patch edge
                    // that position can touch (a = 0: first row, a = 1: last row) when re (ce) = 0, unmasked when 1
                    float sv[2][2][2][2];
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            float u[3][2];                   // u[dy][ce]
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                const float t0 = xw[2 * pr + a + d][2 * pc + b] * w_s[j][3 * d], t1 = xw[2 * pr + a + d][2 * pc + b + 1] * w_s[j][3 * d + 1],
                                            t2 = xw[2 * pr + a + d][2 * pc + b + 2] * w_s[j][3 * d + 2];
                                u[d][0] = b == 0 ? t1 + t2 : t0 + t1;          // edge column class: no dx = -1 (left) / no dx = +1 (right)
                                u[d][1] = b == 0 ? u[d][0] + t0 : u[d][0] + t2;
                            }
#pragma unroll
                            for (int ce = 0; ce < 2; ++ce) {
                                sv[a][b][0][ce] = a == 0 ? u[1][ce] + u[2][ce] : u[0][ce] + u[1][ce];   // edge row class: no dy = -1 / no dy = +1
                                sv[a][b][1][ce] = a == 0 ? sv[a][b][0][ce] + u[0][ce] : sv[a][b][0][ce] + u[2][ce];
                            }
                        }
#pragma unroll
                    for (int er = 0; er < 2; ++er)
#pragma unroll
                        for (int ec = 0; ec < 2; ++ec) {
                            // er = 0: this phase's edge row variant (first if pr == 0: the upper conv row is masked; last if pr == 1: the lower)
                            const int r0e = (er == 0 && pr == 0) ? 0 : 1, r1e = (er == 0 && pr == 1) ? 0 : 1;
                            const int c0e = (ec == 0 && pc == 0) ? 0 : 1, c1e = (ec == 0 && pc == 1) ? 0 : 1;
                            const float m = fmaxf(fmaxf(sv[0][0][r0e][c0e], sv[0][1][r0e][c1e]), fmaxf(sv[1][0][r1e][c0e], sv[1][1][r1e][c1e]));
                            const float y = fmaf(m, sc_s[j], sh_s[j]);
                            if (j & 1) out[er * 2 + ec][j >> 1] = tc::pack_f16x2_relu(even[er * 2 + ec], y);
                            else even[er * 2 + ec] = y;
                        }
                }
                uint4* o = reinterpret_cast<uint4*>(B1) + (((int64_t)(pr * 2 + pc) * 4 + ch) * R1 + Il) * C1 + J;
#pragma unroll
                for (int er = 0; er < 2; ++er)
#pragma unroll
                    for (int ec = 0; ec < 2; ++ec) {
                        const int v = (er ? 1 : (pr ? 2 : 0)) * 3 + (ec ? 1 : (pc ? 2 : 0));
                        o[(int64_t)v * 16 * plane] = make_uint4(out[er * 2 + ec][0], out[er * 2 + ec][1], out[er * 2 + ec][2], out[er * 2 + ec][3]);
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- head
// The global average pool of pixel (x, y) is a strided gather from the 9 F planes,
//     g = 1/(p/2)^2 * sum_{k,l} F[cls(k), cls(l)][x + 2k][y + 2l],      cls = first / interior / last cell,
// evaluated separably: fuse_rowsum_kernel (dense_tc.cuh) forms the inner sums S[a][X][y] = sum_l F[a, cls(l)][X][y + 2l] once
// per map row and row class a (fp32, [3][16][rows][W][8]) in the epilogue of the fusion conv; head_dense_kernel adds the p/2 rows S[cls(k)][x + 2k][y] of a pixel and
// runs the two linears (the first one on the tensor pipe), the first-maximum argmax, the confusion matrix (per-block shared
// histogram -> 64-bit global atomics) and the label map.  Summation order: l ascending inside a row, then k ascending.
// One block = 128 consecutive pixels of one anchor row.
//   phase 1 (8 warps): warp w = channel chunks w and w + 8, lane = pixel (4 passes of 32): the p/2 row sums are added, scaled and written
//     to shared memory as the A operand of the first linear, split into bf16 hi + lo parts (g = hi + lo to ~2^-17);
//   phase 2 (one elected thread): Linear 128->64 on the tensor pipe, 8 k-steps x 3 tcgen05.mma M128 x N64 x K16
//     (g_hi.w_hi + g_lo.w_hi + g_hi.w_lo: fp32-grade products, fp32 accumulation in 64 TMEM columns);
//   phase 3 (4 warps, thread = pixel = TMEM lane): + bias, ReLU, Linear 64->C from shared-memory weights (k ascending),
//     first-maximum argmax, logits / label map / confusion matrix (warp-aggregated shared histogram).
constexpr int kDenseHeadThreads = 256, kHeadPx = 128;        // 8 warps x <= 128 registers: two blocks per SM
constexpr uint32_t kHeadAPlane = kHeadPx * 16;                       // one 8-channel chunk of the A operand
constexpr uint32_t kHeadABytes = (C_FUSE / 8) * kHeadAPlane;         // 32 KB per hi / lo part
constexpr uint32_t kHeadWBytes = (C_FUSE / 8) * C_HID * 16;          // 16 KB per hi / lo part
static size_t dense_head_smem(int C) {
    const int Cp = (C + 3) & ~3;
    return 2 * kHeadABytes + 2 * kHeadWBytes + sizeof(float) * (C_HID + C_HID * Cp + Cp) + sizeof(unsigned int) * C * C + 64;
}

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

template <int P2>
__global__ void __launch_bounds__(kDenseHeadThreads, 2) head_dense_kernel(const uint4* __restrict__ S, int rows, int nb, int W, int C,
                                                                       const __nv_bfloat16* __restrict__ w1_hilo /* [2][16][64][8] */,
                                                                       const float* __restrict__ fc1b, const float* __restrict__ fc2t,
                                                                       const float* __restrict__ fc2b,
                                                                       int64_t pix0 /* flat index of the band's first pixel */,
                                                                       const uint8_t* __restrict__ label, float* __restrict__ logits_out,
                                                                       unsigned long long* __restrict__ cm, uint8_t* __restrict__ pred_map) {
    extern __shared__ __align__(1024) uint8_t hsm[];
    uint8_t* a_hi = hsm;                                       // [16 chunks][128 px][8] bf16 — K-major, no swizzle
    uint8_t* a_lo = a_hi + kHeadABytes;
    uint8_t* w_hi = a_lo + kHeadABytes;                        // [16 chunks][64 out][8] bf16
    uint8_t* w_lo = w_hi + kHeadWBytes;
    const int Cp = (C + 3) & ~3;
    float* b1 = reinterpret_cast<float*>(w_lo + kHeadWBytes);
    float* w2 = b1 + C_HID;                                    // [64][Cp]
    float* b2 = w2 + C_HID * Cp;
    unsigned int* hist = reinterpret_cast<unsigned int*>(b2 + Cp);
    uint64_t* bar = reinterpret_cast<uint64_t*>(hist + C * C + ((C * C) & 1));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (int)(2 * kHeadWBytes / 16); i += blockDim.x)
        reinterpret_cast<uint4*>(w_hi)[i] = __ldg(reinterpret_cast<const uint4*>(w1_hilo) + i);
    if (threadIdx.x < C_HID) b1[threadIdx.x] = fc1b[threadIdx.x];
    for (int i = threadIdx.x; i < C_HID * Cp; i += blockDim.x) {
        const int k = i / Cp, c = i - k * Cp;
        w2[i] = c < C ? fc2t[k * C + c] : 0.f;
    }
    if (threadIdx.x < Cp) b2[threadIdx.x] = threadIdx.x < C ? fc2b[threadIdx.x] : 0.f;
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    const uint32_t bar_a = tc::smem_u32(bar);
    if (threadIdx.x == 0) {
        tc::mbar_init(bar_a, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the weight copies above feed the tensor pipe
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const float inv = 1.0f / (float)P2;                 // S holds row means: the column mean over the P2 rows is left
    const int segs = (W + kHeadPx - 1) / kHeadPx;
    const int n_seg = nb * segs;
    uint32_t phase = 0;
    for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        const int xl = seg / segs, y0 = (seg - xl * segs) * kHeadPx;
        // ---- phase 1: column sums of the row sums -> A operand (hi / lo)
#pragma unroll 1
        for (int it = 0; it < 2 * (kHeadPx / 32); ++it) {
            const int chunk = warp + 8 * (it & 1), pass = it >> 1;
            const int px = pass * 32 + lane, y = y0 + px;
            float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (y < W) {
#pragma unroll
                for (int k0 = 0; k0 < P2; k0 += 4) {
                    uint4 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = k0 + e, a = k == 0 ? 0 : (k == P2 - 1 ? 2 : 1);
                        v[e] = __ldg(S + (((int64_t)a * 16 + chunk) * rows + xl + 2 * k) * W + y);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t u[4] = {v[e].x, v[e].y, v[e].z, v[e].w};
#pragma unroll
                        for (int h2 = 0; h2 < 4; ++h2) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[h2]));
                            s[2 * h2] += f.x;
                            s[2 * h2 + 1] += f.y;
                        }
                    }
                }
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float g0 = s[2 * j] * inv, g1 = s[2 * j + 1] * inv;
                const __nv_bfloat16 h0 = __float2bfloat16_rn(g0), h1 = __float2bfloat16_rn(g1);
                hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lo[j] = tc::pack_bf16x2(g0 - __bfloat162float(h0), g1 - __bfloat162float(h1));
            }
            *reinterpret_cast<uint4*>(a_hi + (uint32_t)chunk * kHeadAPlane + (uint32_t)px * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(a_lo + (uint32_t)chunk * kHeadAPlane + (uint32_t)px * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async-proxy (tensor pipe) reads
        __syncthreads();
        // ---- phase 2: Linear 128 -> 64 on the tensor pipe
        if (warp == 1) {
            tc::tc_fence_after();
            if (tc::elect_one()) {
                constexpr uint32_t idesc = tc::umma_idesc_bf16(128, C_HID);
                const uint64_t ah = tc::umma_desc(tc::smem_u32(a_hi), kHeadAPlane, 128), al = tc::umma_desc(tc::smem_u32(a_lo), kHeadAPlane, 128);
                const uint64_t wh = tc::umma_desc(tc::smem_u32(w_hi), C_HID * 16, 128), wl = tc::umma_desc(tc::smem_u32(w_lo), C_HID * 16, 128);
#pragma unroll
                for (int j = 0; j < C_FUSE / 16; ++j) {
                    const uint64_t ao = (uint64_t)((2u * j * kHeadAPlane) >> 4), wo = (uint64_t)((2u * j * C_HID * 16) >> 4);
                    tc::umma_bf16(tmem_base, ah + ao, wh + wo, idesc, j ? 1u : 0u);
                    tc::umma_bf16(tmem_base, al + ao, wh + wo, idesc, 1u);
                    tc::umma_bf16(tmem_base, ah + ao, wl + wo, idesc, 1u);
                }
                tc::umma_commit(bar_a);
            }
            __syncwarp();
        }
        // ---- phase 3: thread = pixel
        if (warp < 4) {
            tc::mbar_wait(bar_a, phase);
            tc::tc_fence_after();
            const int px = warp * 32 + lane, y = y0 + px;
            const bool live = y < W;
            float hid[C_HID];
            {
                uint32_t v[32];
                const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
                tc::tmem_ld32(t_row, v);
#pragma unroll
                for (int k = 0; k < 32; ++k) hid[k] = fmaxf(__uint_as_float(v[k]) + b1[k], 0.f);
                tc::tmem_ld32(t_row + 32, v);
#pragma unroll
                for (int k = 0; k < 32; ++k) hid[32 + k] = fmaxf(__uint_as_float(v[k]) + b1[32 + k], 0.f);
            }
            tc::tc_fence_before();
            const int64_t n = (int64_t)xl * W + y;                          // pixel index inside the band
            float bv = -INFINITY;
            int bi = 0;
            for (int c4 = 0; c4 < C; c4 += 4) {
                float4 a = *reinterpret_cast<const float4*>(b2 + c4);
#pragma unroll
                for (int k = 0; k < C_HID; ++k) {
                    const float4 w = *reinterpret_cast<const float4*>(w2 + k * Cp + c4);
                    a.x = fmaf(hid[k], w.x, a.x); a.y = fmaf(hid[k], w.y, a.y); a.z = fmaf(hid[k], w.z, a.z); a.w = fmaf(hid[k], w.w, a.w);
                }
                const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (c4 + e < C) {
                        if (live && logits_out) logits_out[n * C + c4 + e] = av[e];
                        if (av[e] > bv) { bv = av[e]; bi = c4 + e; }           // strict >: the first maximum (torch.max semantics)
                    }
                }
            }
            int key = 0;
            bool count = false;
            if (live) {
                const int64_t kflat = pix0 + n;
                if (pred_map) pred_map[kflat] = (uint8_t)bi;
                if (cm) {
                    const int lab = label[kflat];
                    count = lab < C;
                    key = bi * C + lab;
                }
            }
            if (cm) hist_add_warp(hist, key, count);
        }
        phase ^= 1;
        __syncthreads();                                                   // A operand and TMEM columns are free again
    }
    __syncthreads();
    if (cm)
        for (int i = threadIdx.x; i < C * C; i += blockDim.x)
            if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
static int make_dense_map(CUtensorMap* m, const void* base, int planes, int kch, int rows, int cols, int box_cols, int box_rows, int box_kch) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return DMF_ERR_CUDA; }
    cuuint64_t dims[4] = {8ull * cols, (cuuint64_t)planes, (cuuint64_t)rows, (cuuint64_t)kch};
    cuuint64_t strides[3] = {(cuuint64_t)kch * rows * cols * 16, (cuuint64_t)cols * 16, (cuuint64_t)rows * cols * 16};
    cuuint32_t box[4] = {(cuuint32_t)(8 * box_cols), 1, (cuuint32_t)box_rows, (cuuint32_t)box_kch}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (dense map %dx%d, %d planes) failed: CUresult %d", rows, cols, planes, (int)r); return DMF_ERR_CUDA; }
    return DMF_OK;
}

// Pooled class (a, b) of a fused conv + pool layer (conv_pool4_kernel): per axis, input offset o = s + dy in {-1, 0, 1, 2}
// from the cell origin has border variant kOffVariant[a][o + 1] (-1 = outside the patch).  ALIGNED pooling (phase-separated
// input): offset o lives in phase o & 1 at cell shift floor(o / 2); stride-1 pooling: phase 0, shift o.  Sources = distinct
// (variant, phase) pairs of an axis; a box = row source x column source, its origin the smallest shift any user needs.
static const int kOffVariant[3][4] = {{-1, 0, 1, 1}, {1, 1, 1, 1}, {1, 1, 2, -1}};
struct AxisSrc { int variant, phase, origin; };
static int axis_sources(int a, bool aligned, AxisSrc* src, int* src_of_off, int* shift_of_off) {
    int n = 0;
    for (int o = -1; o <= 2; ++o) {
        src_of_off[o + 1] = -1;
        const int v = kOffVariant[a][o + 1];
        if (v < 0) continue;
        const int ph = aligned ? (o & 1) : 0, sh = aligned ? (o < 0 ? -1 : o / 2) : o;
        shift_of_off[o + 1] = sh;
        int k = 0;
        while (k < n && !(src[k].variant == v && src[k].phase == ph)) ++k;
        if (k == n) { src[n].variant = v; src[n].phase = ph; src[n].origin = sh; ++n; }
        src[k].origin = std::min(src[k].origin, sh);
        src_of_off[o + 1] = k;
    }
    return n;
}
static int build_pool4_cls(tc::Pool4Cls& c, int a, int b, bool aligned, int box_rows, int box_cols, int max_boxes, uint32_t box_slot) {
    memset(&c, 0, sizeof(c));
    AxisSrc rs[4], cs[4];
    int r_of[4], c_of[4], r_sh[4], c_sh[4];
    const int nr = axis_sources(a, aligned, rs, r_of, r_sh), nc = axis_sources(b, aligned, cs, c_of, c_sh);
    if (nr * nc > max_boxes) { set_error("pool4 class (%d,%d) needs %d boxes", a, b, nr * nc); return DMF_ERR_STATE; }
    c.out_plane = (int16_t)(a * 3 + b);
    c.n_boxes = (int16_t)(nr * nc);
    // stride-1 pooling, interior class on an axis: sub-position 1 of a cell IS sub-position 0 of the next cell -> evaluated once (dense_tc.cuh, SHARE)
    c.ns = (int16_t)(!aligned && a == 1 ? 1 : 2);
    c.nt = (int16_t)(!aligned && b == 1 ? 1 : 2);
    c.ca = (int16_t)a;
    c.cb = (int16_t)b;
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) {
            const int k = i * nc + j;
            const int variant = rs[i].variant * 3 + cs[j].variant;
            c.box_plane[k] = (int16_t)(aligned ? variant * 4 + rs[i].phase * 2 + cs[j].phase : variant);
            c.box_drow[k] = (int8_t)rs[i].origin;
            c.box_dcol[k] = (int8_t)cs[j].origin;
        }
    // A window at offset (orow, ocol) from the cell origin, shared by every (sub-position, tap) pair that reads it
    for (int orow = -1; orow <= 2; ++orow)
        for (int ocol = -1; ocol <= 2; ++ocol) {
            const int wi = (orow + 1) * 4 + ocol + 1;
            const int i = r_of[orow + 1], j = c_of[ocol + 1];
            if (i < 0 || j < 0) { c.win[wi] = -1; continue; }
            const int dr = r_sh[orow + 1] - rs[i].origin, dc = c_sh[ocol + 1] - cs[j].origin;
            if (dr < 0 || dr + 16 > box_rows || dc < 0 || dc + 8 > box_cols) { set_error("pool4 class (%d,%d): window outside its box", a, b); return DMF_ERR_STATE; }
            c.win[wi] = (int16_t)(((uint32_t)(i * nc + j) * box_slot + (uint32_t)(dr * box_cols + dc) * 16) >> 4);
        }
    return DMF_OK;
}

// conv + pool of one layer over the band: in = 9 (x 4 phases) planes, out = 9 pooled planes
template <int CI, int CO, int KQ, int STAGES, int BR, int BC, int NBUF, int EW, bool ALIGNED, bool SHARE>
static int launch_pool4(const CUtensorMap& map, const __nv_bfloat16* w, const float* scale, const float* shift, __nv_bfloat16* out,
                        int out_chunks, int out_chunk0, int nb, int W, int p, int R1, int C1, cudaStream_t st) {
    constexpr bool aligned = ALIGNED;             // pan2 pools on the aligned (phase-separated) grid; the stride-1 layers can share sub-positions
    static_assert(!(ALIGNED && SHARE), "sub-positions are shared between the cells of a stride-1 pooling grid only");
    using Cfg = tc::Pool4Cfg<CI, CO, KQ, STAGES, BR, BC, NBUF, SHARE>;
    static_assert(Cfg::SMEM <= (size_t)kSmemLimit, "conv_pool4_kernel does not fit in shared memory");
    tc::Pool4Params P{};
    P.rows = R1; P.cols = C1;
    // Output cells per patch axis: aligned pooling (pan2) p cells at X = x + k; stride-1 pooling p/2 cells at X = x + 2k.  A border
    // class (first / interior / last cell) is only ever read at the rows the band's anchors x in [0, nb) put it on.
    const int cells = aligned ? p : p / 2, step = aligned ? 1 : 2;
    const int k_lo[3] = {0, 1, cells - 1}, k_hi[3] = {0, cells - 2, cells - 1};
    P.tiles_x = P.tiles_y = 0;
    for (int a = 0; a < 3; ++a) {
        P.row_lo[a] = P.col_lo[a] = step * k_lo[a];
        P.row_n[a] = nb + step * (k_hi[a] - k_lo[a]);
        P.col_n[a] = W + step * (k_hi[a] - k_lo[a]);
        P.trow[a] = SHARE && a == 1 ? 15 : 16;    // the interior class hands its last tile row / column to the next tile
        P.tcol[a] = SHARE && a == 1 ? 7 : 8;
        P.tiles_y = std::max(P.tiles_y, cdiv(P.row_n[a], P.trow[a]));
        P.tiles_x = std::max(P.tiles_x, cdiv(P.col_n[a], P.tcol[a]));
    }
    P.n_tiles = P.tiles_x * P.tiles_y * 9;
    P.out_chunks = out_chunks; P.out_chunk0 = out_chunk0;
    P.w = w; P.scale = scale; P.shift = shift; P.out = out;
    static const int dbg = getenv("DMF_DENSE_DBG") ? atoi(getenv("DMF_DENSE_DBG")) : 0;      // timing diagnostics (results are wrong when set)
    P.dbg = dbg;
    // table order: interior / interior (1 TMEM slot per tile), the four classes with one interior axis (2 slots), the corners (4): tiles
    // of equal size follow each other, so that the slot ring keeps two or more of the small ones in flight
    static const int kOrder[9] = {4, 1, 7, 3, 5, 0, 2, 6, 8};
    for (int i = 0; i < 9; ++i) {
        DMF_TRY(build_pool4_cls(P.cls[i], kOrder[i] / 3, kOrder[i] % 3, aligned, BR, BC, Cfg::MAX_BOXES, Cfg::BOX_SLOT));
        if (!SHARE) P.cls[i].ns = P.cls[i].nt = 2;
    }
    auto kern = tc::conv_pool4_kernel<CI, CO, KQ, STAGES, BR, BC, NBUF, EW, SHARE>;
    static bool attr_set = false;
    if (!attr_set) {
        DMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    kern<<<std::min(P.n_tiles, num_sms()), 64 + 32 * EW, Cfg::SMEM, st>>>(map, P);
    DMF_LAUNCHED();
    return DMF_OK;
}

int dense_pack(dmf_net* n) {
    if (!n->dense) n->dense = new DenseWs();
    DenseWs* d = n->dense;
    auto* w1 = param(n, "ms1.0.weight", (size_t)C_MS1 * 4 * 9);
    auto* wp = param(n, "pan1.0.weight", (size_t)C_PAN1 * 9);
    if (!w1 || !wp) return DMF_ERR_STATE;
    DMF_TRY(to_device(&d->w_ms1, *w1));
    DMF_TRY(to_device(&d->w_pan1, *wp));
    const char* blk[3] = {"ms2", "pan2", "pan3"};
    const int cin[3] = {C_MS1, C_PAN1, C_PAN2}, cout[3] = {C_MS2, C_PAN2, C_PAN3};
    for (int l = 0; l < 3; ++l) {
        auto* w = param(n, std::string(blk[l]) + ".0.weight", (size_t)cout[l] * cin[l] * 9);
        if (!w) return DMF_ERR_STATE;
        std::vector<float> sc, sh;
        DMF_TRY(fold_bn(n, blk[l], cout[l], sc, sh));
        std::vector<__nv_bfloat16> pk((size_t)9 * cin[l] * cout[l]);
        for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < cin[l]; ++ci)
                for (int co = 0; co < cout[l]; ++co) {
                    const float v = (*w)[((size_t)co * cin[l] + ci) * 9 + tap];
                    pk[(((size_t)(ci / 8) * 9 + tap) * cout[l] + co) * 8 + ci % 8] = tc::w16(sc[co] < 0.f ? -v : v);   // [ci/8][tap][co][8], fp16
                }
        for (auto& v : sc) v = fabsf(v);
        DMF_TRY(to_device(&d->w_cp[l], pk));
        DMF_TRY(to_device(&d->sc_cp[l], sc));
        DMF_TRY(to_device(&d->sh_cp[l], sh));
    }
    auto* f1 = param(n, "fc1.weight", (size_t)C_HID * C_FUSE);
    if (!f1) return DMF_ERR_STATE;
    std::vector<__nv_bfloat16> hl((size_t)2 * C_FUSE * C_HID);
    for (int o = 0; o < C_HID; ++o)
        for (int k = 0; k < C_FUSE; ++k) {
            const float w = (*f1)[(size_t)o * C_FUSE + k];
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const size_t at = ((size_t)(k / 8) * C_HID + o) * 8 + k % 8;
            hl[at] = hi;
            hl[(size_t)C_FUSE * C_HID + at] = __float2bfloat16_rn(w - __bfloat162float(hi));
        }
    DMF_TRY(to_device(&d->w_fc1, hl));
    return DMF_OK;
}

static void dense_free_ws(DenseWs* d) {
    __nv_bfloat16* bs[] = {d->A, d->CAT, d->B1, d->B2};
    for (auto* b : bs) cudaFree(b);
    cudaFree(d->S);
    d->S = nullptr;
    d->A = d->CAT = d->B1 = d->B2 = nullptr;
    d->W = d->band = 0;
    d->bytes = 0;
}

void dense_release(dmf_net* n) {
    if (!n->dense) return;
    DenseWs* d = n->dense;
    dense_free_ws(d);
    cudaFree(d->w_ms1); cudaFree(d->w_pan1); cudaFree(d->w_fc1);
    for (int l = 0; l < 3; ++l) { cudaFree(d->w_cp[l]); cudaFree(d->sc_cp[l]); cudaFree(d->sh_cp[l]); }
    for (auto& e : d->ev) if (e) cudaEventDestroy(e);
    delete d;
    n->dense = nullptr;
}

// workspace for bands of `band` anchor rows of a scene W pixels wide
static int dense_prepare(dmf_net* n, int W, int band) {
    DenseWs* d = n->dense;
    if (d->A && d->W == W && band <= d->band && d->p == n->p) return DMF_OK;      // a smaller band runs inside the existing planes
    DMF_CUDA(cudaDeviceSynchronize());
    dense_free_ws(d);
    const int p = n->p;
    const size_t R1 = band + p - 1, C1 = W + p - 1, px = R1 * C1;
    const size_t sA = px * 9 * C_MS1 * 2, sCAT = px * 9 * C_CAT * 2 + 2048 /* fuse_rowsum_kernel's last 2 KB row copy may run past the tensor */, sB1 = px * 4 * 9 * C_PAN1 * 2,
                 sB2 = px * 9 * C_PAN2 * 2;
    DMF_CUDA(cudaMalloc(&d->A, sA));
    DMF_CUDA(cudaMalloc(&d->CAT, sCAT));
    DMF_CUDA(cudaMalloc(&d->B1, sB1));
    DMF_CUDA(cudaMalloc(&d->B2, sB2));
    const size_t sS = R1 * (size_t)W * 3 * C_FUSE * sizeof(__half);
    DMF_CUDA(cudaMalloc(&d->S, sS));
    d->bytes = sA + sCAT + sB1 + sB2 + sS;
    // positions a band never writes are only ever read into don't-care outputs; zero them once so that runs are reproducible
    DMF_CUDA(cudaMemset(d->A, 0, sA));
    DMF_CUDA(cudaMemset(d->CAT, 0, sCAT));
    DMF_CUDA(cudaMemset(d->B1, 0, sB1));
    DMF_CUDA(cudaMemset(d->B2, 0, sB2));
    d->W = W; d->band = band; d->p = p; d->R1 = (int)R1; d->C1 = (int)C1;
    DMF_TRY(make_dense_map(&d->mapA, d->A, 9, C_MS1 / 8, (int)R1, (int)C1, 11, 19, 2));
    DMF_TRY(make_dense_map(&d->mapB1, d->B1, 36, C_PAN1 / 8, (int)R1, (int)C1, 9, 17, C_PAN1 / 8));
    DMF_TRY(make_dense_map(&d->mapB2s, d->B2, 9, C_PAN2 / 8, (int)R1, (int)C1, 11, 19, 2));
    for (auto& e : d->ev) if (!e) DMF_CUDA(cudaEventCreate(&e));
    DMF_CUDA(cudaFuncSetAttribute(tc::fuse_rowsum_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    DMF_CUDA(cudaFuncSetAttribute(tc::fuse_rowsum_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    DMF_CUDA(cudaFuncSetAttribute(tc::fuse_rowsum_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    DMF_CUDA(cudaFuncSetAttribute(head_dense_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DMF_CUDA(cudaFuncSetAttribute(head_dense_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DMF_CUDA(cudaFuncSetAttribute(head_dense_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    return DMF_OK;
}

static int grid_for(int64_t work_items, int threads, int per_sm) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((work_items + threads - 1) / threads, (int64_t)num_sms() * per_sm));
}

int dense_infer(dmf_net* n, const dmf_scene* s, int row0, int row1, float* logits_dev, uint8_t* pred_map_dev, int64_t* cm_dev, cudaStream_t st) {
    DenseWs* d = n->dense;
    if (!d || !d->w_ms1) { set_error("dense path: weights not packed"); return DMF_ERR_STATE; }
    const int p = n->p, W = s->W;
    const int band = std::max(1, std::min(n->dense_band, row1 - row0));
    DMF_TRY(dense_prepare(n, W, band));
    const int R1 = d->R1, C1 = d->C1;
    const bool tm = n->timing;
    int evi = 0;
    auto mark = [&]() { if (tm && evi < 12) cudaEventRecord(d->ev[evi++], st); };

    for (int b0 = row0; b0 < row1; b0 += band) {
        const int nb = std::min(band, row1 - b0);
        const int rows = nb + p - 1;                      // map rows this band needs (<= R1)
        evi = 0;
        mark();
        // ---- MS branch
        ms_stem_map_kernel<<<grid_for((int64_t)rows * C1, 256, 8), 256, 0, st>>>(
            reinterpret_cast<const float4*>(s->ms), s->Hp, s->Wp, b0, rows, R1, C1, d->w_ms1, n->L[4].scale, n->L[4].shift, d->A);
        DMF_LAUNCHED();
        mark();
        // DMF_DENSE_SHARE=0 (diagnostics): every cell of the stride-1 layers evaluates all four sub-positions itself, as pan2 must
        static const bool share = !(getenv("DMF_DENSE_SHARE") && atoi(getenv("DMF_DENSE_SHARE")) == 0);
        if (share) DMF_TRY((launch_pool4<C_MS1, C_MS2, 2, 2, 19, 11, 1, 8, false, true>(d->mapA, d->w_cp[0], d->sc_cp[0], d->sh_cp[0], d->CAT, C_CAT / 8, 0, nb, W, p, R1, C1, st)));
        else DMF_TRY((launch_pool4<C_MS1, C_MS2, 2, 3, 19, 11, 1, 8, false, false>(d->mapA, d->w_cp[0], d->sc_cp[0], d->sh_cp[0], d->CAT, C_CAT / 8, 0, nb, W, p, R1, C1, st)));
        mark();
        mark();          // (stage slot of the former separate pooling pass)
        // ---- PAN branch
        pan_stem_map_kernel<<<grid_for((int64_t)rows * C1, 256, 8), 256, 0, st>>>(
            n->use_mspan ? s->mspan : s->pan, s->H4p, s->W4p, s->pan_pitch, b0, rows, R1, C1, d->w_pan1, n->sc_pan1, n->sh_pan1, d->B1);
        DMF_LAUNCHED();
        mark();
        DMF_TRY((launch_pool4<C_PAN1, C_PAN2, 4, 2, 17, 9, 2, 8, true, false>(d->mapB1, d->w_cp[1], d->sc_cp[1], d->sh_cp[1], d->B2, C_PAN2 / 8, 0, nb, W, p, R1, C1, st)));
        mark();
        mark();
        if (share) DMF_TRY((launch_pool4<C_PAN2, C_PAN3, 2, 2, 19, 11, 1, 8, false, true>(d->mapB2s, d->w_cp[2], d->sc_cp[2], d->sh_cp[2], d->CAT, C_CAT / 8, C_MS2 / 8, nb, W, p, R1, C1, st)));
        else DMF_TRY((launch_pool4<C_PAN2, C_PAN3, 2, 3, 19, 11, 1, 8, false, false>(d->mapB2s, d->w_cp[2], d->sc_cp[2], d->sh_cp[2], d->CAT, C_CAT / 8, C_MS2 / 8, nb, W, p, R1, C1, st)));
        mark();
        mark();
        // ---- fusion conv (1x1) on the 9 pooled planes + row sums of the global average pool
        {
            tc::FuseRowsParams P{};
            const int valid = 128 - 2 * (p / 2 - 1);
            P.rows = rows; P.W = W; P.tiles_x = cdiv(W, valid); P.n_tiles = rows * 3 * P.tiles_x; P.s_rows = rows;
            static const int dbg = getenv("DMF_DENSE_DBG") ? atoi(getenv("DMF_DENSE_DBG")) : 0;      // timing diagnostics (results are wrong when set)
            P.dbg = dbg;
            P.w = n->L[3].w; P.scale = n->L[3].scale; P.shift = n->L[3].shift; P.S = reinterpret_cast<uint4*>(d->S);
            P.cat = d->CAT; P.R1 = R1; P.C1 = C1;
            {   // pooled cells k of a patch sit at X = x + 2k: first k = 0, interior 1 .. p/2 - 2, last p/2 - 1
                const int P2 = p / 2, k_lo[3] = {0, 1, P2 - 1}, k_hi[3] = {0, P2 - 2, P2 - 1};
                for (int a = 0; a < 3; ++a) { P.row_lo[a] = 2 * k_lo[a]; P.row_n[a] = nb + 2 * (k_hi[a] - k_lo[a]); }
            }
            auto fk = p == 8 ? tc::fuse_rowsum_kernel<4> : p == 16 ? tc::fuse_rowsum_kernel<8> : tc::fuse_rowsum_kernel<16>;
            fk<<<std::min(P.n_tiles, num_sms()), 320, tc::kFrSmem, st>>>(P);
            DMF_LAUNCHED();
        }
        mark();
        // ---- head
        {
            const int n_seg = nb * ((W + kHeadPx - 1) / kHeadPx);
            const int64_t off = (int64_t)(b0 - row0) * W;
            auto kern = p == 8 ? head_dense_kernel<4> : p == 16 ? head_dense_kernel<8> : head_dense_kernel<16>;
            kern<<<std::min(n_seg, 2 * num_sms()), kDenseHeadThreads, dense_head_smem(n->C), st>>>(      // 2 blocks per SM (~100 KB each)
                reinterpret_cast<const uint4*>(d->S), rows, nb, W, n->C, d->w_fc1, n->fc1b, n->fc2t, n->fc2b, (int64_t)b0 * W, s->label,
                logits_dev ? logits_dev + off * n->C : nullptr, reinterpret_cast<unsigned long long*>(cm_dev), pred_map_dev);
            DMF_LAUNCHED();
        }
        mark();
        if (tm) {
            DMF_CUDA(cudaEventSynchronize(d->ev[evi - 1]));
            for (int i = 0; i + 1 < evi; ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, d->ev[i], d->ev[i + 1]);
                d->stage_ms[i] += ms;
            }
            float tot = 0.f;
            cudaEventElapsedTime(&tot, d->ev[0], d->ev[evi - 1]);
            d->stage_ms[11] += tot;
        }
    }
    return DMF_OK;
}

}  // namespace dmf

using namespace dmf;

extern "C" {

int dmf_net_set_dense(dmf_net* n, int enabled, int band_rows) {
    DMF_REQUIRE(n, "net_set_dense: null");
    DMF_REQUIRE(band_rows == 0 || (band_rows >= 1 && band_rows <= 4096), "net_set_dense: band_rows must be in [1, 4096] (0 keeps the current value)");
    n->dense_mode = enabled ? 1 : 0;
    if (band_rows) n->dense_band = band_rows;
    return DMF_OK;
}

int dmf_infer_scene_dense(dmf_net* n, const dmf_scene* s, int row0, int row1, float* logits_out_dev, uint8_t* pred_map_dev, int64_t* cm_dev,
                          void* stream) {
    if (!n || !n->ready) { set_error("net: call dmf_net_finalize after loading all parameters"); return DMF_ERR_STATE; }
    DMF_REQUIRE(s && row0 >= 0 && row1 >= row0 && row1 <= s->H, "infer_scene_dense: bad row band [%d,%d)", row0, row1);
    DMF_REQUIRE(s->p == n->p, "infer_scene_dense: scene patch size %d != net patch size %d", s->p, n->p);
    DMF_REQUIRE(!cm_dev || s->label, "infer_scene_dense: confusion matrix needs dmf_scene_set_labels");
    DMF_REQUIRE(!n->use_mspan || s->mspan, "infer_scene_dense: the IHS product was selected as input but the scene has none (dmf_scene_set_mspan)");
    if (row1 == row0) return DMF_OK;
    return dense_infer(n, s, row0, row1, logits_out_dev, pred_map_dev, cm_dev, (cudaStream_t)stream);
}

/* test hook (host only, no GPU): the class table conv_pool4_kernel gets for pooled border class (a, b).  aligned = 1: pan2 (phase
 * planes, boxes of 17 x 9 cells, 4 chunks per step); 0: the stride-1 layers (19 x 11, 2 chunks).  win[16], box_plane[9], box_drow[9],
 * box_dcol[9]; returns the box slot size in bytes through *slot_bytes and the box count through *n_boxes. */
int dmf_dense_class_table(int a, int b, int aligned, int16_t* win, int16_t* box_plane, int8_t* box_drow, int8_t* box_dcol, int32_t* n_boxes,
                          int32_t* slot_bytes) {
    DMF_REQUIRE(a >= 0 && a < 3 && b >= 0 && b < 3 && win && box_plane && box_drow && box_dcol && n_boxes && slot_bytes, "dense_class_table: bad argument");
    tc::Pool4Cls c;
    using CfgA = tc::Pool4Cfg<C_PAN1, C_PAN2, 4, 2, 17, 9, 2>;
    using CfgS = tc::Pool4Cfg<C_MS1, C_MS2, 2, 2, 19, 11, 1, true>;
    DMF_TRY(aligned ? build_pool4_cls(c, a, b, true, 17, 9, CfgA::MAX_BOXES, CfgA::BOX_SLOT) : build_pool4_cls(c, a, b, false, 19, 11, CfgS::MAX_BOXES, CfgS::BOX_SLOT));
    memcpy(win, c.win, sizeof(c.win));
    memcpy(box_plane, c.box_plane, sizeof(c.box_plane));
    memcpy(box_drow, c.box_drow, sizeof(c.box_drow));
    memcpy(box_dcol, c.box_dcol, sizeof(c.box_dcol));
    *n_boxes = c.n_boxes;
    *slot_bytes = (int32_t)(aligned ? CfgA::BOX_SLOT : CfgS::BOX_SLOT);
    return DMF_OK;
}

int dmf_net_get_dense_timing(dmf_net* n, float out_ms[12], int reset) {
    DMF_REQUIRE(n && out_ms, "net_get_dense_timing: null");
    if (!n->dense) { memset(out_ms, 0, 12 * sizeof(float)); return DMF_OK; }
    memcpy(out_ms, n->dense->stage_ms, 12 * sizeof(float));
    if (reset) memset(n->dense->stage_ms, 0, sizeof(n->dense->stage_ms));
    return DMF_OK;
}

/* test hook: device pointer + byte size of a dense-path map ("A", "CAT", "B1", "B2", "S"); dims[0..1] = R, C of the MS-resolution grid */
int dmf_net_dense_buffer(dmf_net* n, const char* name, void** ptr_out, int64_t* bytes_out, int32_t dims[2]) {
    DMF_REQUIRE(n && name && ptr_out && bytes_out && dims, "net_dense_buffer: null");
    DMF_REQUIRE(n->dense && n->dense->A, "net_dense_buffer: the dense path has not run yet");
    DenseWs* d = n->dense;
    const size_t px = (size_t)d->R1 * d->C1;
    const std::string k(name);
    if (k == "A") { *ptr_out = d->A; *bytes_out = px * 9 * C_MS1 * 2; }
    else if (k == "CAT") { *ptr_out = d->CAT; *bytes_out = px * 9 * C_CAT * 2; }
    else if (k == "B1") { *ptr_out = d->B1; *bytes_out = px * 4 * 9 * C_PAN1 * 2; }
    else if (k == "B2") { *ptr_out = d->B2; *bytes_out = px * 9 * C_PAN2 * 2; }
    else if (k == "S") { *ptr_out = d->S; *bytes_out = (int64_t)d->R1 * d->W * 3 * C_FUSE * sizeof(__half); }
    else { set_error("net_dense_buffer: unknown map '%s'", name); return DMF_ERR_ARG; }
    dims[0] = d->R1; dims[1] = d->C1;
    return DMF_OK;
}

}  // extern "C"
