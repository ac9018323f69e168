// GMFNet on sm_100a: weight packing, the two CUDA-core stems (fused with the patch gather), the
// tcgen05 conv layers (conv_tc.cuh), the head (GAP + 2 linears + argmax + confusion matrix) and the
// chunked whole-scene driver.  Replaces model.gmfnet.Net behind solver/mainsolver.py:30-38,52,109,169
// and the loops of Solver.test()/color() (solver/mainsolver.py:104-141, 167-185).
//
// Data flow for a chunk of NB pixels (activations in the C8-planar bf16 layout, see conv_tc.cuh):
//   scene --stem_ms--> A1[NB][8][p][p][8]   --conv ms2 (+pool)--> CAT[NB][0..15][p/2][p/2][8]
//   scene --stem_pan-> B1[NB][4][2p][2p][8] --conv pan2 (+pool)-> B2[NB][8][p][p][8]
//                                            --conv pan3 (+pool)-> CAT[NB][16..31][p/2][p/2][8]
//   CAT --conv 1x1 fuse--> F[NB][16][p/2][p/2][8] --head--> logits / pred / confusion matrix
// The stems read the patch windows straight from the padded scene (K1 is never materialised on the
// inference path).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "conv_tc.cuh"
#include "net_geom.cuh"
#include "net_types.cuh"
#include "stem_tc.cuh"

using namespace dmf;

namespace dmf {

// ------------------------------------------------------------------------------------ stems
struct PatchSrc {
    // either windows of a scene (idx == null -> consecutive pixels from `first`) ...
    dmf_scene scene;
    const int64_t* idx;
    int64_t first;
    // ... or materialised patches [N][4][p][p] / [N][1][4p][4p]
    const float* patches;
};

// MS stem, step 1 (step 2 is a tcgen05 layer): the 4-band fp32 window of every patch is split into
// bf16 hi + lo parts (x = hi + lo to ~2^-16) and written as a 16-channel C8-planar tensor
//   chunk 0 = [hi0..hi3, lo0..lo3]    chunk 1 = [hi0..hi3, 0, 0, 0, 0]
// The stem weights are packed to match ([w_hi, w_hi, w_lo, 0], net_finalize), so that one K=16
// tcgen05.mma per tap evaluates x_hi*w_hi + x_lo*w_hi + x_hi*w_lo: fp32-grade products on the bf16
// tensor pipe.  One thread per pixel; 16-byte coalesced stores.
template <bool FROM_SCENE>
__global__ void __launch_bounds__(256) ms_prep_kernel(PatchSrc src, int p, int64_t N, __nv_bfloat16* __restrict__ out) {
    const int pp = p * p;
    const int64_t total = N * pp;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = t / pp;
        const int px = (int)(t - n * pp);
        const int r = px / p, c = px - r * p;
        float4 v;
        if (FROM_SCENE) {
            const int64_t k = src.idx ? src.idx[n] : src.first + n;
            const int x = (int)(k / src.scene.W), y = (int)(k % src.scene.W);
            v = __ldg(reinterpret_cast<const float4*>(src.scene.ms) + (int64_t)(x + r) * src.scene.Wp + y + c);
        } else {
            const float* b = src.patches + n * 4 * pp + px;
            v = make_float4(b[0], b[pp], b[2 * pp], b[3 * pp]);
        }
        const float f[4] = {v.x, v.y, v.z, v.w};
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            hi[i] = __bfloat162float(__float2bfloat16_rn(f[i]));
            lo[i] = f[i] - hi[i];
        }
        const uint32_t h01 = tc::pack_bf16x2(hi[0], hi[1]), h23 = tc::pack_bf16x2(hi[2], hi[3]);
        uint4* o = reinterpret_cast<uint4*>(out + (n * 2 * pp + px) * 8);
        o[0] = make_uint4(h01, h23, tc::pack_bf16x2(lo[0], lo[1]), tc::pack_bf16x2(lo[2], lo[3]));
        o[pp] = make_uint4(h01, h23, 0u, 0u);
    }
}

// ------------------------------------------------------------------------------------ head
// F[N][16][px][8] bf16 -> global average pool -> Linear 128->64 + ReLU -> Linear 64->C -> logits,
// argmax (first maximum), confusion matrix, prediction map.  One WARP per patch, no block barriers:
// coalesced 512-byte reads of F, an exchange-halving warp reduction for the pooling (9 shuffles per
// 8 channels instead of 40), the two small linears from shared-memory weights, shuffle argmax.

// GAPIN: the pooled sums come from the fusion conv's epilogue (gap [N][128] fp32) instead of F.
template <bool GAPIN>
__global__ void __launch_bounds__(32 * kHeadWarps) head_kernel(const __nv_bfloat16* __restrict__ F, const float* __restrict__ gap,
                                                              int64_t N, int npx, int C,
                                                              const float* __restrict__ fc1t, const float* __restrict__ fc1b,
                                                              const float* __restrict__ fc2t, const float* __restrict__ fc2b,
                                                              const dmf_scene scene, const int64_t* __restrict__ idx, int64_t first,
                                                              float* __restrict__ logits_out, uint8_t* __restrict__ pred_out,
                                                              unsigned long long* __restrict__ cm, uint8_t* __restrict__ pred_map) {
    extern __shared__ __align__(16) float hs[];
    float* w1 = hs;                                   // [128][64]
    float* w2 = w1 + C_FUSE * C_HID;                  // [64][C]
    float* b1 = w2 + C_HID * C;
    float* b2 = b1 + C_HID;
    float* gbuf = b2 + ((C + 3) & ~3);                // per warp: g[128] + hid[64]
    unsigned int* hist = reinterpret_cast<unsigned int*>(gbuf + kHeadWarps * (C_FUSE + C_HID));   // [C*C]
    for (int i = threadIdx.x; i < C_FUSE * C_HID; i += blockDim.x) w1[i] = fc1t[i];
    for (int i = threadIdx.x; i < C_HID * C; i += blockDim.x) w2[i] = fc2t[i];
    if (threadIdx.x < C_HID) b1[threadIdx.x] = fc1b[threadIdx.x];
    if (threadIdx.x < C) b2[threadIdx.x] = fc2b[threadIdx.x];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* g = gbuf + warp * (C_FUSE + C_HID);
    float* hid = g + C_FUSE;
    const float inv = 1.0f / (float)npx;
    const int64_t wstride = (int64_t)gridDim.x * kHeadWarps;
    for (int64_t n = (int64_t)blockIdx.x * kHeadWarps + warp; n < N; n += wstride) {
        if (GAPIN) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(gap + n * C_FUSE) + lane);
            *reinterpret_cast<float4*>(g + 4 * lane) = make_float4(s4.x * inv, s4.y * inv, s4.z * inv, s4.w * inv);
        }
        const uint4* base = reinterpret_cast<const uint4*>(F) + n * 16 * npx;
        for (int chunk = 0; chunk < (GAPIN ? 0 : 16); ++chunk) {
            // the 32 lanes sweep the npx pixels of this 8-channel chunk, 16 bytes each (512 B per sweep)
            float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int px = lane; px < npx; px += 32) {
                const uint4 v = __ldg(base + chunk * npx + px);
                const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[k]));
                    s[2 * k] += f.x; s[2 * k + 1] += f.y;
                }
            }
            // exchange-halving reduction: 8 -> 4 -> 2 -> 1 values per lane, then two butterfly steps
            float t4[4], t2[2], t1;
            {
                const bool up = lane & 16;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float send = up ? s[k] : s[k + 4], keep = up ? s[k + 4] : s[k];
                    t4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            {
                const bool up = lane & 8;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = up ? t4[k] : t4[k + 2], keep = up ? t4[k + 2] : t4[k];
                    t2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
            }
            {
                const bool up = lane & 4;
                const float send = up ? t2[0] : t2[1], keep = up ? t2[1] : t2[0];
                t1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
            t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
            // lane bits (4,3,2) select the channel: bit4 -> +4, bit3 -> +2, bit2 -> +1
            if ((lane & 3) == 0) g[chunk * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = t1 * inv;
        }
        __syncwarp();
        float h0 = b1[lane], h1 = b1[lane + 32];
#pragma unroll 4
        for (int k4 = 0; k4 < C_FUSE; k4 += 4) {
            const float4 gv = *reinterpret_cast<const float4*>(g + k4);
            const float gk[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                h0 = fmaf(gk[e], w1[(k4 + e) * C_HID + lane], h0);
                h1 = fmaf(gk[e], w1[(k4 + e) * C_HID + lane + 32], h1);
            }
        }
        hid[lane] = fmaxf(h0, 0.f);
        hid[lane + 32] = fmaxf(h1, 0.f);
        __syncwarp();
        // logits: classes lane and lane + 32 (C <= 64)
        float l0 = -INFINITY, l1 = -INFINITY;
        if (lane < C) {
            float a = b2[lane];
#pragma unroll 8
            for (int k = 0; k < C_HID; ++k) a = fmaf(hid[k], w2[k * C + lane], a);
            l0 = a;
            if (logits_out) logits_out[n * C + lane] = a;
        }
        if (lane + 32 < C) {
            float a = b2[lane + 32];
#pragma unroll 8
            for (int k = 0; k < C_HID; ++k) a = fmaf(hid[k], w2[k * C + lane + 32], a);
            l1 = a;
            if (logits_out) logits_out[n * C + lane + 32] = a;
        }
        // argmax with torch.max semantics: the first (lowest) index among equal maxima
        float bv = l0;
        int bi = lane;
        if (l1 > bv) { bv = l1; bi = lane + 32; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            if (pred_out) pred_out[n] = (uint8_t)bi;
            if (cm || pred_map) {
                const int64_t k = idx ? idx[n] : first + n;
                if (pred_map) pred_map[k] = (uint8_t)bi;
                if (cm) {
                    const int lab = scene.label[k];
                    if (lab < C) atomicAdd(&hist[bi * C + lab], 1u);
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (cm)
        for (int i = threadIdx.x; i < C * C; i += blockDim.x)
            if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

// ------------------------------------------------------------------------------------ device-side debug conv
// Plain CUDA-core convolution over the same layouts and packed weights; fp32 accumulation, same
// epilogue.  Exists only so tests can localise a fault to one layer on the GPU; the product path
// never calls it.
__global__ void direct_conv_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ w,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   __nv_bfloat16* __restrict__ out, int64_t N, int S, int cin, int cout, int taps,
                                   int pool, int out_chunks, int out_chunk0) {
    const int So = pool ? S / 2 : S;
    const int64_t total = N * So * So * cout;
    const int kch = cin / 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % cout);
        int64_t r = i / cout;
        const int ow = (int)(r % So); r /= So;
        const int oh = (int)(r % So);
        const int64_t n = r / So;
        float best = -INFINITY;
        const int reps = pool ? 2 : 1;
        for (int oy = 0; oy < reps; ++oy)
            for (int ox = 0; ox < reps; ++ox) {
                const int h = pool ? 2 * oh + oy : oh, wv = pool ? 2 * ow + ox : ow;
                float acc = 0.f;
                for (int tap = 0; tap < taps; ++tap) {
                    const int dy = taps == 9 ? tap / 3 - 1 : 0, dx = taps == 9 ? tap % 3 - 1 : 0;
                    const int hh = h + dy, ww = wv + dx;
                    if (hh < 0 || hh >= S || ww < 0 || ww >= S) continue;
                    for (int ci = 0; ci < cin; ++ci) {
                        const float a = __half2float(reinterpret_cast<const __half*>(in)[(((n * kch + ci / 8) * S + hh) * S + ww) * 8 + ci % 8]);
                        const float b = __half2float(reinterpret_cast<const __half*>(w)[(((int64_t)tap * kch + ci / 8) * cout + co) * 8 + ci % 8]);
                        acc = fmaf(a, b, acc);
                    }
                }
                const float yv = fmaxf(fmaf(acc, scale[co], shift[co]), 0.f);
                best = fmaxf(best, __half2float(__float2half_rn(fminf(yv, 65504.f))));
            }
        reinterpret_cast<__half*>(out)[(((n * out_chunks + out_chunk0 + co / 8) * So + oh) * So + ow) * 8 + co % 8] = __float2half_rn(best);
    }
}

// ------------------------------------------------------------------------------------ host side
const std::vector<float>* param(const dmf_net* n, const std::string& k, size_t numel) {
    auto it = n->params.find(k);
    if (it == n->params.end()) { set_error("net: parameter '%s' was not loaded", k.c_str()); return nullptr; }
    if (it->second.size() != numel) {
        set_error("net: parameter '%s' has %zu elements, expected %zu", k.c_str(), it->second.size(), numel);
        return nullptr;
    }
    return &it->second;
}

int fold_bn(const dmf_net* n, const std::string& blk, int cout, std::vector<float>& scale, std::vector<float>& shift) {
    auto *cb = param(n, blk + ".0.bias", cout), *gw = param(n, blk + ".1.weight", cout), *gb = param(n, blk + ".1.bias", cout),
         *mu = param(n, blk + ".1.running_mean", cout), *var = param(n, blk + ".1.running_var", cout);
    if (!cb || !gw || !gb || !mu || !var) return DMF_ERR_STATE;
    scale.resize(cout); shift.resize(cout);
    for (int c = 0; c < cout; ++c) {
        const float s = (*gw)[c] / sqrtf((*var)[c] + BN_EPS);
        scale[c] = s;
        shift[c] = (*gb)[c] + ((*cb)[c] - (*mu)[c]) * s;
    }
    return DMF_OK;
}

static int pack_conv(dmf_net* n, ConvLayer& L, const std::string& blk) {
    const int cin = L.g.cin, cout = L.g.cout, taps = L.g.taps, kch = cin / 8;
    auto* w = param(n, blk + ".0.weight", (size_t)cout * cin * taps);
    if (!w) return DMF_ERR_STATE;
    std::vector<__nv_bfloat16> pk((size_t)taps * cin * cout);
    for (int tap = 0; tap < taps; ++tap)
        for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
                pk[(((size_t)tap * kch + ci / 8) * cout + co) * 8 + ci % 8] = tc::w16((*w)[((size_t)co * cin + ci) * taps + tap]);
    L.f16 = true;
    std::vector<float> sc, sh;
    DMF_TRY(fold_bn(n, blk, cout, sc, sh));
    DMF_TRY(to_device(&L.w, pk));
    DMF_TRY(to_device(&L.scale, sc));
    DMF_TRY(to_device(&L.shift, sh));
    return DMF_OK;
}

// row-pair packing: tap' = dy'*3 + dx over a 4 x 3 window, row n = s*cout + co holds W[co][dy' - s][dx]
static int pack_conv_rowpair(dmf_net* n, ConvLayer& L, const std::string& blk) {
    const int cin = L.g.cin, cout = L.g.cout, kch = cin / 8, N2 = 2 * cout;
    auto* w = param(n, blk + ".0.weight", (size_t)cout * cin * 9);
    if (!w) return DMF_ERR_STATE;
    std::vector<__nv_bfloat16> pk((size_t)12 * cin * N2, __float2bfloat16_rn(0.f));
    for (int s2 = 0; s2 < 2; ++s2)
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx)
                for (int ci = 0; ci < cin; ++ci)
                    for (int co = 0; co < cout; ++co) {
                        const int tap = (dy + s2) * 3 + dx;
                        pk[(((size_t)tap * kch + ci / 8) * N2 + s2 * cout + co) * 8 + ci % 8] = tc::w16((*w)[((size_t)co * cin + ci) * 9 + dy * 3 + dx]);
                    }
    L.f16 = true;
    std::vector<float> sc, sh;
    DMF_TRY(fold_bn(n, blk, cout, sc, sh));
    DMF_TRY(to_device(&L.w, pk));
    DMF_TRY(to_device(&L.scale, sc));
    DMF_TRY(to_device(&L.shift, sh));
    return DMF_OK;
}

// MS stem weights for the hi/lo-split input: per tap k = [w_hi(4), w_hi(4), w_lo(4), 0(4)]
static int pack_ms_stem(dmf_net* n, ConvLayer& L) {
    const int cout = C_MS1;
    auto* w = param(n, "ms1.0.weight", (size_t)cout * 4 * 9);
    if (!w) return DMF_ERR_STATE;
    std::vector<__nv_bfloat16> pk((size_t)9 * 16 * cout, __float2bfloat16_rn(0.f));
    for (int tap = 0; tap < 9; ++tap)
        for (int co = 0; co < cout; ++co)
            for (int ci = 0; ci < 4; ++ci) {
                const float wv = (*w)[((size_t)co * 4 + ci) * 9 + tap];
                const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
                const __nv_bfloat16 lo = __float2bfloat16_rn(wv - __bfloat162float(hi));
                auto at = [&](int k) -> __nv_bfloat16& { return pk[(((size_t)tap * 2 + k / 8) * cout + co) * 8 + k % 8]; };
                at(ci) = hi; at(4 + ci) = hi; at(8 + ci) = lo;
            }
    std::vector<float> sc, sh;
    DMF_TRY(fold_bn(n, "ms1", cout, sc, sh));
    DMF_TRY(to_device(&L.w, pk));
    DMF_TRY(to_device(&L.scale, sc));
    DMF_TRY(to_device(&L.shift, sh));
    return DMF_OK;
}

template <int CI, int CO, int TAPS, bool POOL, int G, int NP, bool GAPOUT = false>
static int launch_conv(const ConvLayer& L, const CUtensorMap& map, __nv_bfloat16* out, int out_chunks, int out_chunk0,
                       int64_t N, cudaStream_t st, int dbg = 0, float* gap = nullptr) {
    tc::ConvParams P;
    const LayerGeom& g = L.g;
    P.S = g.S; P.S_l2 = g.S_l2; P.NP = g.NP; P.NP_l2 = g.NP_l2; P.TH = g.TH; P.tiles_x_l2 = g.tiles_x_l2;
    P.PX = g.PX; P.PX_l2 = g.PX_l2; P.tpg_l2 = g.tpg_l2;
    P.N = (int)N;
    P.n_tiles = (int)((N + g.NP - 1) / g.NP) * g.tiles_per_group;
    P.a_plane = g.a_plane; P.a_stage = g.a_stage; P.n_stage = g.n_stage; P.sbo_a = g.sbo_a;
    P.out_chunks = out_chunks; P.out_chunk0 = out_chunk0; P.dbg = dbg; P.f16_in = L.f16 ? 1 : 0;
    P.stats = nullptr; P.stat_stride = 0;
    P.w = L.w; P.scale = L.scale; P.shift = L.shift; P.out = out; P.gap = gap;
    if (GAPOUT) DMF_CUDA(cudaMemsetAsync(gap, 0, sizeof(float) * (size_t)N * CO, st));
    if (TAPS == 9 && (g.NP != NP || g.TH != 16 / NP)) { set_error("conv geometry/template mismatch"); return DMF_ERR_STATE; }
    auto kern = tc::conv_tc_kernel<CI, CO, TAPS, POOL, G, NP, GAPOUT>;
    static bool attr_set = false;
    if (!attr_set) {
        DMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    const int grid = std::min(P.n_tiles, num_sms());
    kern<<<grid, 64 + 128 * G, g.smem, st>>>(map, P);
    DMF_LAUNCHED();
    return DMF_OK;
}

template <int CI, int G>
static int launch_rowpair(const ConvLayer& L, const CUtensorMap& map, __nv_bfloat16* out, int out_chunks, int out_chunk0,
                          int64_t N, cudaStream_t st) {
    tc::ConvParams P{};
    const LayerGeom& g = L.g;
    P.S = g.S; P.S_l2 = g.S_l2; P.NP = 1; P.NP_l2 = 0; P.TH = g.TH; P.tiles_x_l2 = g.tiles_x_l2; P.tpg_l2 = g.tpg_l2;
    P.N = (int)N; P.n_tiles = (int)N * g.tiles_per_group;
    P.a_plane = g.a_plane; P.a_stage = g.a_stage; P.n_stage = g.n_stage; P.sbo_a = g.sbo_a;
    P.out_chunks = out_chunks; P.out_chunk0 = out_chunk0; P.f16_in = L.f16 ? 1 : 0;
    P.w = L.w; P.scale = L.scale; P.shift = L.shift; P.out = out;
    auto kern = tc::conv_rowpair_kernel<CI, G>;
    static bool attr_set = false;
    if (!attr_set) {
        DMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    kern<<<std::min(P.n_tiles, num_sms()), 64 + 128 * G, g.smem, st>>>(map, P);
    DMF_LAUNCHED();
    return DMF_OK;
}

static int run_layer(dmf_net* n, int layer, const CUtensorMap& map, __nv_bfloat16* out, int64_t N, cudaStream_t st, int dbg) {
    const int ocat = C_CAT / 8;
    const bool np2 = n->L[layer].g.NP == 2;       // 8x8 maps (p = 8): two patches per 128-pixel tile
    switch (layer) {
        case 0: return np2 ? launch_conv<C_MS1, C_MS2, 9, true, 3, 2>(n->L[0], map, out, ocat, 0, N, st, dbg)
                           : launch_conv<C_MS1, C_MS2, 9, true, 3, 1>(n->L[0], map, out, ocat, 0, N, st, dbg);
        case 1: return n->L[1].g.rowpair ? launch_rowpair<C_PAN1, 3>(n->L[1], map, out, C_PAN2 / 8, 0, N, st)
                                         : launch_conv<C_PAN1, C_PAN2, 9, true, 4, 1>(n->L[1], map, out, C_PAN2 / 8, 0, N, st, dbg);
        case 2: return np2 ? launch_conv<C_PAN2, C_PAN3, 9, true, 3, 2>(n->L[2], map, out, ocat, C_MS2 / 8, N, st, dbg)
                           : launch_conv<C_PAN2, C_PAN3, 9, true, 3, 1>(n->L[2], map, out, ocat, C_MS2 / 8, N, st, dbg);
        case 3: return (n->gap && out == n->F) ? launch_conv<C_CAT, C_FUSE, 1, false, 2, 1, true>(n->L[3], map, out, C_FUSE / 8, 0, N, st, dbg, n->gap)
                                               : launch_conv<C_CAT, C_FUSE, 1, false, 2, 1>(n->L[3], map, out, C_FUSE / 8, 0, N, st, dbg);
        case 4: return np2 ? launch_conv<16, C_MS1, 9, false, 4, 2>(n->L[4], map, out, C_MS1 / 8, 0, N, st, dbg)
                           : launch_conv<16, C_MS1, 9, false, 4, 1>(n->L[4], map, out, C_MS1 / 8, 0, N, st, dbg);
    }
    return DMF_ERR_ARG;
}

static int run_layer(dmf_net* n, int layer, const CUtensorMap& map, __nv_bfloat16* out, int64_t N, cudaStream_t st, int dbg);
static int stem_pan_stages(int p) { (void)p; return tc::kStemMaxStages; }
static int stem_pan_raw_pitch(int p) { return 4 * p + 8; }
static size_t stem_pan_smem(int p) {
    const size_t raw = 2ull * (4 * p + 2) * stem_pan_raw_pitch(p) * 4;
    return (size_t)stem_pan_stages(p) * tc::kStemStage + tc::kStemWBytes + 2 * tc::kStemCout * 4 + 20 * 8 + 16 + raw;
}
size_t head_smem(int C) {
    return sizeof(float) * (C_FUSE * C_HID + C_HID * C + C_HID + ((C + 3) & ~3) + kHeadWarps * (C_FUSE + C_HID)) +
           sizeof(unsigned int) * C * C;
}

static int launch_stems(dmf_net* n, const PatchSrc& src, bool from_scene, int64_t N, __nv_bfloat16* A1, __nv_bfloat16* B1,
                        int which, cudaStream_t st) {
    constexpr int GP = 2;                  // epilogue groups of the PAN stem kernel
    static bool attr_set = false;
    if (!attr_set) {
        DMF_CUDA(cudaFuncSetAttribute(tc::stem_pan_tc_kernel<GP, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    if (which & 1) {
        const int grid = (int)std::min<int64_t>((N * n->p * n->p + 255) / 256, (int64_t)num_sms() * 16);
        if (from_scene) ms_prep_kernel<true><<<grid, 256, 0, st>>>(src, n->p, N, n->X0);
        else ms_prep_kernel<false><<<grid, 256, 0, st>>>(src, n->p, N, n->X0);
        DMF_LAUNCHED();
        DMF_TRY(run_layer(n, 4, n->L[4].map, A1, N, st, 0));
    }
    if (which & 2) {
        auto l2 = [](int v) { int e = 0; while ((1 << e) < v) ++e; return e; };
        tc::StemPanParams Q;
        Q.scene_pan = (from_scene && n->use_mspan) ? src.scene.mspan : src.scene.pan; Q.pan_pitch = src.scene.pan_pitch; Q.scene_W = src.scene.W;
        Q.idx = src.idx; Q.first = src.first; Q.patches = src.patches; Q.from_scene = from_scene ? 1 : 0;
        Q.p = n->p; Q.S_l2 = l2(2 * n->p); Q.tpp_l2 = l2(4 * n->p * n->p / 128); Q.N = N;
        Q.n_stage = stem_pan_stages(n->p); Q.raw_pitch = stem_pan_raw_pitch(n->p);
        Q.w = n->w_pan1; Q.shift = n->sh_pan1; Q.out = B1;
        const int grid = (int)std::min<int64_t>(N, num_sms());
        tc::stem_pan_tc_kernel<GP, 4><<<grid, 320 + 128 * GP, stem_pan_smem(n->p), st>>>(Q);
        DMF_LAUNCHED();
    }
    return DMF_OK;
}

// one chunk (N <= NB patches) through the whole network
static int forward_chunk(dmf_net* n, const PatchSrc& src, bool from_scene, int64_t N, float* logits, uint8_t* pred,
                         int64_t* cm, uint8_t* pred_map, cudaStream_t st) {
    const bool tm = n->timing;
    if (tm) cudaEventRecord(n->ev[0], st);
    DMF_TRY(launch_stems(n, src, from_scene, N, n->A1, n->B1, 1, st));
    if (tm) cudaEventRecord(n->ev[1], st);
    DMF_TRY(run_layer(n, 0, n->L[0].map, n->CAT, N, st, 0));
    if (tm) cudaEventRecord(n->ev[2], st);
    DMF_TRY(launch_stems(n, src, from_scene, N, n->A1, n->B1, 2, st));
    if (tm) cudaEventRecord(n->ev[3], st);
    DMF_TRY(run_layer(n, 1, n->L[1].map, n->B2, N, st, 0));
    if (tm) cudaEventRecord(n->ev[4], st);
    DMF_TRY(run_layer(n, 2, n->L[2].map, n->CAT, N, st, 0));
    if (tm) cudaEventRecord(n->ev[5], st);
    DMF_TRY(run_layer(n, 3, n->L[3].map, n->F, N, st, 0));
    if (tm) cudaEventRecord(n->ev[6], st);
    const int npx = (n->p / 2) * (n->p / 2);
    const int grid = (int)std::min<int64_t>((N + kHeadWarps - 1) / kHeadWarps, (int64_t)num_sms() * 4);
    if (n->gap)
        head_kernel<true><<<grid, 32 * kHeadWarps, head_smem(n->C), st>>>(n->F, n->gap, N, npx, n->C, n->fc1t, n->fc1b, n->fc2t, n->fc2b,
                                                                          src.scene, src.idx, src.first, logits, pred,
                                                                          (unsigned long long*)cm, pred_map);
    else
        head_kernel<false><<<grid, 32 * kHeadWarps, head_smem(n->C), st>>>(n->F, nullptr, N, npx, n->C, n->fc1t, n->fc1b, n->fc2t, n->fc2b,
                                                                           src.scene, src.idx, src.first, logits, pred,
                                                                           (unsigned long long*)cm, pred_map);
    DMF_LAUNCHED();
    if (tm) {
        cudaEventRecord(n->ev[7], st);
        DMF_CUDA(cudaEventSynchronize(n->ev[7]));
        for (int i = 0; i < 7; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, n->ev[i], n->ev[i + 1]);
            n->stage_ms[i] += ms;
        }
        float tot = 0.f;
        cudaEventElapsedTime(&tot, n->ev[0], n->ev[7]);
        n->stage_ms[7] += tot;
    }
    return DMF_OK;
}

}  // namespace dmf

extern "C" {

int dmf_net_create(dmf_net** out, int p, int num_classes, int max_batch) {
    DMF_REQUIRE(out, "net_create: null");
    DMF_REQUIRE(p == 8 || p == 16 || p == 32, "net_create: patch_size must be 8, 16 or 32 (got %d)", p);
    DMF_REQUIRE(num_classes >= 2 && num_classes <= 64, "net_create: 2 <= Categories_Number <= 64");
    DMF_REQUIRE(max_batch >= 1 && max_batch <= (1 << 20), "net_create: bad max_batch");
    dmf_net* n = new dmf_net();
    n->p = p; n->C = num_classes; n->NB = max_batch;
    int rc = make_geom(n->L[0].g, p, 9, C_MS1, C_MS2, 1);
    if (rc == DMF_OK) rc = make_geom(n->L[1].g, 2 * p, 9, C_PAN1, C_PAN2, 1, (2 * p) % 32 == 0);
    if (rc == DMF_OK) rc = make_geom(n->L[2].g, p, 9, C_PAN2, C_PAN3, 1);
    if (rc == DMF_OK) rc = make_geom(n->L[3].g, p / 2, 1, C_CAT, C_FUSE, 0);
    if (rc == DMF_OK) rc = make_geom(n->L[4].g, p, 9, 16, C_MS1, 0);
    if (rc != DMF_OK) { delete n; return rc; }
    *out = n;
    return DMF_OK;
}

int dmf_net_destroy(dmf_net* n) {
    if (!n) return DMF_OK;
    cudaFree(n->w_pan1);
    float* fs[] = {n->sc_pan1, n->sh_pan1, n->fc1t, n->fc1b, n->fc2t, n->fc2b};
    for (float* f : fs) cudaFree(f);
    for (auto& L : n->L) { cudaFree(L.w); cudaFree(L.scale); cudaFree(L.shift); }
    __nv_bfloat16* bs[] = {n->X0, n->A1, n->B1, n->B2, n->CAT, n->F};
    for (auto* b : bs) cudaFree(b);
    cudaFree(n->gap);
    for (auto& e : n->ev) if (e) cudaEventDestroy(e);
    dense_release(n);
    delete n;
    return DMF_OK;
}

int dmf_net_load_param(dmf_net* n, const char* name, const float* data_host, int64_t numel) {
    DMF_REQUIRE(n && name && data_host && numel > 0, "net_load_param: bad argument");
    n->params[name].assign(data_host, data_host + numel);
    n->ready = false;
    return DMF_OK;
}

int64_t dmf_net_flops_per_patch(const dmf_net* n) {
    if (!n) return 0;
    const int64_t p = n->p;
    auto conv = [](int64_t cin, int64_t cout, int64_t k, int64_t h) { return 2 * cin * cout * k * k * h * h; };
    return conv(4, C_MS1, 3, p) + conv(C_MS1, C_MS2, 3, p) + conv(1, C_PAN1, 3, 4 * p) + conv(C_PAN1, C_PAN2, 3, 2 * p) +
           conv(C_PAN2, C_PAN3, 3, p) + conv(C_CAT, C_FUSE, 1, p / 2) + 2 * C_FUSE * C_HID + 2 * C_HID * n->C;
}

int dmf_net_finalize(dmf_net* n, void* stream) {
    DMF_REQUIRE(n, "net_finalize: null");
    (void)stream;
    const int p = n->p, C = n->C;
    // --- stems
    DMF_TRY(pack_ms_stem(n, n->L[4]));
    {
        auto* w = param(n, "pan1.0.weight", (size_t)C_PAN1 * 9);
        if (!w) return DMF_ERR_STATE;
        // B operand of the PAN stem (stem_tc.cuh): row n = q*32 + co, K = 48 over the 4x4 input region of a
        // pooled pixel: region pixel i = 4*ry + rx carries tap (ry - qy, rx - qx) of window position q
        // (zero outside the 3x3); k = 2i, 2i+1 -> w_hi, k = 32 + i -> w_lo.
        std::vector<__nv_bfloat16> pk((size_t)tc::kStemKch * 128 * 8, __float2bfloat16_rn(0.f));
        std::vector<float> sc, sh;
        DMF_TRY(fold_bn(n, "pan1", C_PAN1, sc, sh));
        for (int q = 0; q < 4; ++q)
            for (int co = 0; co < C_PAN1; ++co)
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) {
                        const float wv = (*w)[co * 9 + dy * 3 + dx] * sc[co];    // BN scale folded in fp32, before the split
                        const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
                        const __nv_bfloat16 lo = __float2bfloat16_rn(wv - __bfloat162float(hi));
                        const int i = 4 * ((q >> 1) + dy) + (q & 1) + dx, row = q * C_PAN1 + co;
                        auto at = [&](int k) -> __nv_bfloat16& { return pk[((size_t)(k / 8) * 128 + row) * 8 + k % 8]; };
                        at(2 * i) = hi; at(2 * i + 1) = hi; at(32 + i) = lo;
                    }
        DMF_TRY(to_device(&n->w_pan1, pk)); DMF_TRY(to_device(&n->sc_pan1, sc)); DMF_TRY(to_device(&n->sh_pan1, sh));
    }
    // --- tensor-core layers
    DMF_TRY(pack_conv(n, n->L[0], "ms2"));
    DMF_TRY(n->L[1].g.rowpair ? pack_conv_rowpair(n, n->L[1], "pan2") : pack_conv(n, n->L[1], "pan2"));
    DMF_TRY(pack_conv(n, n->L[2], "pan3"));
    DMF_TRY(pack_conv(n, n->L[3], "fuse"));
    // --- head
    {
        auto *w1 = param(n, "fc1.weight", (size_t)C_HID * C_FUSE), *b1 = param(n, "fc1.bias", C_HID),
             *w2 = param(n, "fc2.weight", (size_t)C * C_HID), *b2 = param(n, "fc2.bias", C);
        if (!w1 || !b1 || !w2 || !b2) return DMF_ERR_STATE;
        std::vector<float> t1((size_t)C_FUSE * C_HID), t2((size_t)C_HID * C);
        for (int o = 0; o < C_HID; ++o)
            for (int k = 0; k < C_FUSE; ++k) t1[(size_t)k * C_HID + o] = (*w1)[(size_t)o * C_FUSE + k];
        for (int o = 0; o < C; ++o)
            for (int k = 0; k < C_HID; ++k) t2[(size_t)k * C + o] = (*w2)[(size_t)o * C_HID + k];
        DMF_TRY(to_device(&n->fc1t, t1)); DMF_TRY(to_device(&n->fc1b, *b1));
        DMF_TRY(to_device(&n->fc2t, t2)); DMF_TRY(to_device(&n->fc2b, *b2));
    }
    // --- workspace + tensor maps (once)
    if (!n->A1) {
        const size_t NB = n->NB;
        DMF_CUDA(cudaMalloc(&n->X0, NB * 16 * p * p * 2));
        DMF_CUDA(cudaMalloc(&n->A1, NB * C_MS1 * p * p * 2));
        DMF_CUDA(cudaMalloc(&n->B1, NB * C_PAN1 * 4 * p * p * 2));
        DMF_CUDA(cudaMalloc(&n->B2, NB * C_PAN2 * p * p * 2));
        DMF_CUDA(cudaMalloc(&n->CAT, NB * C_CAT * (p / 2) * (p / 2) * 2));
        DMF_CUDA(cudaMalloc(&n->F, NB * C_FUSE * (p / 2) * (p / 2) * 2));
        if ((p / 2) * (p / 2) <= 64) DMF_CUDA(cudaMalloc(&n->gap, NB * C_FUSE * sizeof(float)));   // whole patches per 128-pixel tile
        DMF_TRY(make_map(&n->L[0].map, n->L[0].g, n->A1, n->NB));
        DMF_TRY(make_map(&n->L[1].map, n->L[1].g, n->B1, n->NB));
        DMF_TRY(make_map(&n->L[2].map, n->L[2].g, n->B2, n->NB));
        DMF_TRY(make_map(&n->L[3].map, n->L[3].g, n->CAT, n->NB));
        DMF_TRY(make_map(&n->L[4].map, n->L[4].g, n->X0, n->NB));
        for (auto& e : n->ev) DMF_CUDA(cudaEventCreate(&e));
        DMF_CUDA(cudaFuncSetAttribute(head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        DMF_CUDA(cudaFuncSetAttribute(head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    DMF_TRY(dense_pack(n));
    DMF_CUDA(cudaDeviceSynchronize());
    n->ready = true;
    return DMF_OK;
}

int dmf_net_set_pan_source(dmf_net* n, int use_mspan) {
    DMF_REQUIRE(n, "net_set_pan_source: null");
    n->use_mspan = use_mspan ? 1 : 0;
    return DMF_OK;
}

int dmf_net_set_timing(dmf_net* n, int enabled) {
    DMF_REQUIRE(n, "net_set_timing: null");
    n->timing = enabled != 0;
    memset(n->stage_ms, 0, sizeof(n->stage_ms));
    return DMF_OK;
}
int dmf_net_get_timing(dmf_net* n, float out_ms[8]) {
    DMF_REQUIRE(n && out_ms, "net_get_timing: null");
    memcpy(out_ms, n->stage_ms, sizeof(n->stage_ms));
    return DMF_OK;
}

#define DMF_NET_READY(n)                                                                             \
    do {                                                                                             \
        if (!(n) || !(n)->ready) { dmf::set_error("net: call dmf_net_finalize after loading all parameters"); return DMF_ERR_STATE; } \
    } while (0)

int dmf_net_forward_patches(dmf_net* n, const float* ms_dev, const float* pan_dev, int64_t N, float* logits_out_dev,
                            void* stream) {
    DMF_NET_READY(n);
    DMF_REQUIRE(ms_dev && pan_dev && logits_out_dev && N >= 0, "net_forward_patches: bad argument");
    const int p = n->p;
    for (int64_t o = 0; o < N; o += n->NB) {
        const int64_t nb = std::min<int64_t>(n->NB, N - o);
        // the two stems read different tensors: run them with their own sources
        PatchSrc sm{}; sm.patches = ms_dev + o * 4 * p * p;
        PatchSrc sp{}; sp.patches = pan_dev + o * 16 * p * p;
        cudaStream_t st = (cudaStream_t)stream;
        DMF_TRY(launch_stems(n, sm, false, nb, n->A1, n->B1, 1, st));
        DMF_TRY(launch_stems(n, sp, false, nb, n->A1, n->B1, 2, st));
        DMF_TRY(run_layer(n, 0, n->L[0].map, n->CAT, nb, st, 0));
        DMF_TRY(run_layer(n, 1, n->L[1].map, n->B2, nb, st, 0));
        DMF_TRY(run_layer(n, 2, n->L[2].map, n->CAT, nb, st, 0));
        DMF_TRY(run_layer(n, 3, n->L[3].map, n->F, nb, st, 0));
        const int npx = (p / 2) * (p / 2);
        const int grid = (int)std::min<int64_t>((nb + kHeadWarps - 1) / kHeadWarps, (int64_t)num_sms() * 4);
        dmf_scene none{};
        if (n->gap)
            head_kernel<true><<<grid, 32 * kHeadWarps, head_smem(n->C), st>>>(n->F, n->gap, nb, npx, n->C, n->fc1t, n->fc1b, n->fc2t, n->fc2b,
                                                                              none, nullptr, 0, logits_out_dev + o * n->C, nullptr, nullptr, nullptr);
        else
            head_kernel<false><<<grid, 32 * kHeadWarps, head_smem(n->C), st>>>(n->F, nullptr, nb, npx, n->C, n->fc1t, n->fc1b, n->fc2t, n->fc2b,
                                                                               none, nullptr, 0, logits_out_dev + o * n->C, nullptr, nullptr, nullptr);
        DMF_LAUNCHED();
    }
    return DMF_OK;
}

int dmf_net_forward_scene(dmf_net* n, const dmf_scene* s, const int64_t* flat_idx_dev, int64_t first, int64_t N,
                          float* logits_out_dev, uint8_t* pred_out_dev, int64_t* cm_dev, uint8_t* pred_map_dev, void* stream) {
    DMF_NET_READY(n);
    DMF_REQUIRE(s && N >= 0, "net_forward_scene: bad argument");
    DMF_REQUIRE(s->p == n->p, "net_forward_scene: scene patch size %d != net patch size %d", s->p, n->p);
    DMF_REQUIRE(!cm_dev || s->label, "net_forward_scene: confusion matrix needs dmf_scene_set_labels");
    DMF_REQUIRE(!n->use_mspan || s->mspan, "net_forward_scene: the IHS product was selected as input but the scene has none (dmf_scene_set_mspan)");
    DMF_REQUIRE(flat_idx_dev || (first >= 0 && first + N <= (int64_t)s->H * s->W), "net_forward_scene: pixel range outside the scene");
    for (int64_t o = 0; o < N; o += n->NB) {
        const int64_t nb = std::min<int64_t>(n->NB, N - o);
        PatchSrc src{};
        src.scene = *s;
        src.idx = flat_idx_dev ? flat_idx_dev + o : nullptr;
        src.first = first + o;
        DMF_TRY(forward_chunk(n, src, true, nb, logits_out_dev ? logits_out_dev + o * n->C : nullptr,
                              pred_out_dev ? pred_out_dev + o : nullptr, cm_dev, pred_map_dev, (cudaStream_t)stream));
    }
    return DMF_OK;
}

int dmf_infer_scene(dmf_net* n, const dmf_scene* s, int row0, int row1, uint8_t* pred_map_dev, int64_t* cm_dev, void* stream) {
    DMF_REQUIRE(s && row0 >= 0 && row1 >= row0 && row1 <= s->H, "infer_scene: bad row band [%d,%d)", row0, row1);
    if (n && n->dense_mode) return dmf_infer_scene_dense(n, s, row0, row1, nullptr, pred_map_dev, cm_dev, stream);
    return dmf_net_forward_scene(n, s, nullptr, (int64_t)row0 * s->W, (int64_t)(row1 - row0) * s->W, nullptr, nullptr, cm_dev,
                                 pred_map_dev, stream);
}

int dmf_net_debug_layer(dmf_net* n, int layer, int impl, const void* in_dev, void* out_dev, int64_t N, void* stream) {
    DMF_NET_READY(n);
    DMF_REQUIRE(layer >= 0 && layer < 5 && in_dev && out_dev && N > 0, "net_debug_layer: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const ConvLayer& L = n->L[layer];
    const int och = layer == 0 || layer == 2 ? C_CAT / 8 : L.g.cout / 8;
    const int oc0 = layer == 2 ? C_MS2 / 8 : 0;
    if (impl == 1 && L.g.rowpair) { set_error("net_debug_layer: the CUDA-core debug conv does not read row-pair weights"); return DMF_ERR_UNSUPPORTED; }
    if (impl == 1) {
        const int So = L.g.pool ? L.g.S / 2 : L.g.S;
        const int64_t total = N * So * So * L.g.cout;
        const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
        direct_conv_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)in_dev, L.w, L.scale, L.shift, (__nv_bfloat16*)out_dev, N,
                                                 L.g.S, L.g.cin, L.g.cout, L.g.taps, L.g.pool, och, oc0);
        DMF_LAUNCHED();
        return DMF_OK;
    }
    CUtensorMap map;
    DMF_TRY(make_map(&map, L.g, in_dev, N));
    return run_layer(n, layer, map, (__nv_bfloat16*)out_dev, N, st, impl >= 2 ? impl - 1 : 0);
}

int dmf_net_debug_stem(dmf_net* n, int which, const float* patches_dev, void* out_dev, int64_t N, void* stream) {
    DMF_NET_READY(n);
    DMF_REQUIRE((which == 0 || which == 1) && patches_dev && out_dev && N > 0, "net_debug_stem: bad argument");
    PatchSrc src{};
    src.patches = patches_dev;
    return launch_stems(n, src, false, N, (__nv_bfloat16*)out_dev, (__nv_bfloat16*)out_dev, which == 0 ? 1 : 2, (cudaStream_t)stream);
}

}  // extern "C"
