// Native training step of GMFNet on sm_100a — replaces the inner loop of Solver.train()
// (solver/mainsolver.py:49-55: zero_grad -> forward -> CrossEntropyLoss -> backward -> Adam.step;
// utils/utils.py:12 make_optimizer, :28-29 make_loss) for the network of model/gmfnet.py.
//
// Train-mode semantics follow torch: BatchNorm uses the statistics of the batch (biased variance for the
// normalisation, unbiased for running_var, momentum 0.1), max-pool routes the gradient to the first maximum,
// CrossEntropyLoss(mean), Adam without weight decay.  Arithmetic: bf16 operands, fp32 accumulation on the tensor
// cores (the reference runs fp32 cuDNN; tolerances are stated in tests/test_gpu_train.py).
//
// Forward, per conv layer:   Z = conv(A_in, W)                 conv_tc_kernel MODE 2 (raw bf16 Z + sum z, sum z^2)
//                            A_out = pool(relu(bn(Z)))          bn_apply_kernel (statistics -> scale/shift on the fly)
// Backward, per conv layer:  (sum dy, sum dy z)                bn_bwd_kernel<.., false>   dy from pooled / direct / GAP grad
//                            dZ = bn_backward(dy)              bn_bwd_kernel<.., true>
//                            dW += dZ^T (x) A_in                wgrad_tc_kernel (tcgen05, MN-major operands)
//                            dA_in = conv(dZ, flip(W)^T)       conv_tc_kernel MODE 1 with dgrad-packed weights
// The 1-channel PAN stem (K = 9) runs on CUDA cores in fp32 (pan1_fwd_kernel / pan1_wgrad_kernel); the MS stem uses
// the hi/lo split of the inference path, so both stems see fp32-grade inputs.
#include <math.h>
#include <string.h>

#include <map>
#include <mutex>
#include "common.cuh"
#include "conv_tc.cuh"
#include "net_geom.cuh"
#include "wgrad_tc.cuh"

namespace dmf {

constexpr int T_MS1 = 64, T_MS2 = 128, T_PAN1 = 32, T_PAN2 = 64, T_PAN3 = 128, T_CAT = 256, T_FUSE = 128, T_HID = 64;
constexpr float T_BN_EPS = 1e-5f, T_BN_MOM = 0.1f;
constexpr int kStatStride = 128;     // doubles per statistic row (max channels of a layer)
enum { L_MS1 = 0, L_MS2, L_PAN1, L_PAN2, L_PAN3, L_FUSE, L_COUNT };
static const char* kBlk[L_COUNT] = {"ms1", "ms2", "pan1", "pan2", "pan3", "fuse"};
static const int kCout[L_COUNT] = {T_MS1, T_MS2, T_PAN1, T_PAN2, T_PAN3, T_FUSE};

struct BnRefs {
    const float *gamma, *beta, *bias;
    float *dgamma, *dbeta;
    float *rmean, *rvar;
    int64_t* nbt;
    double* stats;      // [4][kStatStride]: sum z, sum z^2, sum dy, sum dy*z
};

// statistics -> normalisation coefficients of channel c; cnt = elements per channel
__device__ __forceinline__ void bn_coeffs(const double* st, int c, double cnt, float& mean, float& invstd) {
    const double m = st[c] / cnt;
    double var = st[kStatStride + c] / cnt - m * m;
    var = var < 0.0 ? 0.0 : var;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)T_BN_EPS));
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
        f[2 * k] = t.x; f[2 * k + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(tc::pack_bf16x2(f[0], f[1]), tc::pack_bf16x2(f[2], f[3]), tc::pack_bf16x2(f[4], f[5]), tc::pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------ input staging
// MS patches [N][4][p][p] fp32 -> hi/lo-split 16-channel C8-planar tensor (same format as net.cu's ms_prep_kernel)
__global__ void __launch_bounds__(256) ms_split_kernel(const float* __restrict__ patches, int p, int64_t N, __nv_bfloat16* __restrict__ out) {
    const int pp = p * p;
    const int64_t total = N * pp;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = t / pp;
        const int px = (int)(t - n * pp);
        const float* b = patches + n * 4 * pp + px;
        const float f[4] = {b[0], b[pp], b[2 * pp], b[3 * pp]};
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

// ------------------------------------------------------------------------------------ weight packing (every step)
struct PackJob { const float* w; __nv_bfloat16* out; int kind, cin, cout, taps; int64_t n; };
struct PackJobs { PackJob j[10]; int count; };

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackJobs J) {
    const PackJob& q = J.j[blockIdx.y];
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < q.n; o += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(o & 7);
        int64_t r = o >> 3;
        float v = 0.f;
        if (q.kind == 0) {             // forward: [tap][cin/8][cout][8]
            const int co = (int)(r % q.cout); r /= q.cout;
            const int kch = q.cin / 8;
            const int ch = (int)(r % kch);
            const int tap = (int)(r / kch);
            v = q.w[((int64_t)co * q.cin + ch * 8 + e) * q.taps + tap];
        } else if (q.kind == 1) {      // dgrad: [tap][cout/8][cin][8], taps flipped, channel roles exchanged
            const int ci = (int)(r % q.cin); r /= q.cin;
            const int och = q.cout / 8;
            const int ch = (int)(r % och);
            const int tap = (int)(r / och);
            v = q.w[((int64_t)(ch * 8 + e) * q.cin + ci) * q.taps + (q.taps - 1 - tap)];
        } else {                       // MS stem hi/lo: [tap][2][cout][8], k = [w_hi, w_hi | w_lo, 0]
            const int co = (int)(r % q.cout); r /= q.cout;
            const int ch = (int)(r % 2);
            const int tap = (int)(r / 2);
            const int k = ch * 8 + e;
            if (k < 12) {
                const float wv = q.w[((int64_t)co * 4 + (k & 3)) * 9 + tap];
                const float hi = __bfloat162float(__float2bfloat16_rn(wv));
                v = k < 8 ? hi : wv - hi;
            }
        }
        q.out[o] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------ wgrad scratch -> PyTorch layout
// wgrad_tc_kernel accumulates into fp32 scratch laid out [tap][ci][co] (co contiguous = the TMEM lane index, so every
// atomic instruction of a warp hits one 128-byte line).  This adds it into the gradient tensor [co][ci][kh][kw];
// mode 1 (MS stem, hi/lo-split input): dw[co][c] += scratch[c][co] + scratch[4 + c][co].
struct FinishJob { const float* scratch; float* grad; int co, ci, taps, cin_real, mode; };
struct FinishJobs { FinishJob j[6]; int count; };

__global__ void __launch_bounds__(256) wgrad_finish_kernel(const FinishJobs J) {
    const FinishJob& q = J.j[blockIdx.y];
    const int total = q.co * q.cin_real * q.taps;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
        // o enumerates (tap, ci, co) with co fastest so that the scratch reads are coalesced
        const int co = o % q.co;
        const int r = o / q.co;
        const int ci = r % q.cin_real, tap = r / q.cin_real;
        float v = q.scratch[((size_t)tap * q.ci + ci) * q.co + co];
        if (q.mode == 1) v += q.scratch[((size_t)tap * q.ci + ci + 4) * q.co + co];
        q.grad[((size_t)co * q.cin_real + ci) * q.taps + tap] += v;
    }
}

// ------------------------------------------------------------------------------------ PAN stem (CUDA cores, fp32)
// Both stem kernels give one THREAD a column of one 8-channel chunk and march it down `kStemRows` image rows with a
// sliding 3x3 window: per pixel 3 coalesced loads of the new window row, 72 FMAs on register-resident weights /
// accumulators, one 16-byte store (forward) or one 16-byte load (wgrad).  A warp = 32 adjacent columns, so every
// access is a contiguous 128..512-byte run.  Tasks = (patch, row segment, chunk, 32-column group), grid-strided.
constexpr int kStemRows = 16;

// forward: x [N][S][S] fp32 -> Z [N][4][S][S][8] bf16 (conv3x3 1->32, zero padding, no bias) + sum z, sum z^2
__global__ void __launch_bounds__(256) pan1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int S, int64_t N,
                                                       __nv_bfloat16* __restrict__ Z, double* __restrict__ stats) {
    __shared__ float st_s[2][T_PAN1];
    if (threadIdx.x < 2 * T_PAN1) (&st_s[0][0])[threadIdx.x] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int cgroups = S / 32, segs = S / kStemRows;
    const int64_t SS = (int64_t)S * S, tasks = N * segs * 4 * cgroups;
    for (int64_t task = (int64_t)blockIdx.x * wpb + warp; task < tasks; task += (int64_t)gridDim.x * wpb) {
        const int cg = (int)(task % cgroups);
        int64_t r = task / cgroups;
        const int ch = (int)(r & 3); r >>= 2;
        const int seg = (int)(r % segs);
        const int64_t n = r / segs;
        const int c = cg * 32 + lane, h0 = seg * kStemRows;
        float wr[8][9];                          // w is [32][1][3][3]
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 9; ++k) wr[j][k] = __ldg(w + (ch * 8 + j) * 9 + k);
        const float* xp = x + n * SS;
        auto ld = [&](int hh, int cc) { return (hh >= 0 && hh < S && cc >= 0 && cc < S) ? __ldg(xp + (int64_t)hh * S + cc) : 0.f; };
        float win[3][3];
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
            for (int e = 0; e < 3; ++e) win[d + 1][e] = ld(h0 - 1 + d, c - 1 + e);
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        uint4* zo = reinterpret_cast<uint4*>(Z) + ((n * 4 + ch) * SS + (int64_t)h0 * S + c);
#pragma unroll 4
        for (int i = 0; i < kStemRows; ++i) {
#pragma unroll
            for (int e = 0; e < 3; ++e) { win[0][e] = win[1][e]; win[1][e] = win[2][e]; win[2][e] = ld(h0 + i + 1, c - 1 + e); }
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) a = fmaf(win[k / 3][k % 3], wr[j][k], a);
                acc[j] = a;
                s1[j] += a;
                s2[j] = fmaf(a, a, s2[j]);
            }
            zo[(int64_t)i * S] = pack8(acc);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
        float m1 = s1[0], m2 = s2[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) { m1 = lane == j ? s1[j] : m1; m2 = lane == j ? s2[j] : m2; }
        if (lane < 8) {
            atomicAdd(&st_s[0][ch * 8 + lane], m1);
            atomicAdd(&st_s[1][ch * 8 + lane], m2);
        }
    }
    __syncthreads();
    if (threadIdx.x < T_PAN1) {
        atomicAdd(stats + threadIdx.x, (double)st_s[0][threadIdx.x]);
        atomicAdd(stats + kStatStride + threadIdx.x, (double)st_s[1][threadIdx.x]);
    }
}

// weight gradient: dW[co][tap] += sum_px dZ[co][px] * x[px + tap]
__global__ void __launch_bounds__(256) pan1_wgrad_kernel(const __nv_bfloat16* __restrict__ dZ, const float* __restrict__ x, int S,
                                                         int64_t N, float* __restrict__ dw) {
    __shared__ float acc_s[T_PAN1][9];
    for (int i = threadIdx.x; i < 9 * T_PAN1; i += blockDim.x) (&acc_s[0][0])[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int cgroups = S / 32, segs = S / kStemRows;
    const int64_t SS = (int64_t)S * S, tasks = N * segs * cgroups * 4;
    // chunk = task & 3 is fixed per warp when the stride is a multiple of 4: accumulate across tasks in registers
    const int64_t stride = (int64_t)gridDim.x * wpb;         // multiple of 8
    const int ch = warp & 3;
    float acc[8][9];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[j][k] = 0.f;
    for (int64_t task = (int64_t)blockIdx.x * wpb + warp; task < tasks; task += stride) {
        int64_t r = task >> 2;                                // (task & 3) == ch
        const int cg = (int)(r % cgroups); r /= cgroups;
        const int seg = (int)(r % segs);
        const int64_t n = r / segs;
        const int c = cg * 32 + lane, h0 = seg * kStemRows;
        const float* xp = x + n * SS;
        auto ld = [&](int hh, int cc) { return (hh >= 0 && hh < S && cc >= 0 && cc < S) ? __ldg(xp + (int64_t)hh * S + cc) : 0.f; };
        float win[3][3];
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
            for (int e = 0; e < 3; ++e) win[d + 1][e] = ld(h0 - 1 + d, c - 1 + e);
        const uint4* zi = reinterpret_cast<const uint4*>(dZ) + ((n * 4 + ch) * SS + (int64_t)h0 * S + c);
#pragma unroll 4
        for (int i = 0; i < kStemRows; ++i) {
#pragma unroll
            for (int e = 0; e < 3; ++e) { win[0][e] = win[1][e]; win[1][e] = win[2][e]; win[2][e] = ld(h0 + i + 1, c - 1 + e); }
            float g8[8];
            unpack8(__ldg(zi + (int64_t)i * S), g8);
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[j][k] = fmaf(g8[j], win[k / 3][k % 3], acc[j][k]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float v = acc[j][k];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(&acc_s[ch * 8 + j][k], v);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * T_PAN1; i += blockDim.x) atomicAdd(dw + i, (&acc_s[0][0])[i]);
}

// ------------------------------------------------------------------------------------ BatchNorm apply (+ReLU, +2x2 max-pool)
// Z [N][C/8][S][S][8] -> out [N][out_chunks][So][So][8] at chunk offset out_chunk0.  Block 0 also updates the running
// statistics exactly like torch (running_mean includes the conv bias, which Z leaves out because it cancels in BN).
template <bool POOL>
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ Z, BnRefs R, int C, int S, int64_t N,
                                                       __nv_bfloat16* __restrict__ out, int out_chunks, int out_chunk0) {
    __shared__ float sc_s[kStatStride], sh_s[kStatStride];
    const double cnt = (double)N * S * S;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, invstd;
        bn_coeffs(R.stats, c, cnt, mean, invstd);
        const float sc = R.gamma[c] * invstd;
        sc_s[c] = sc;
        sh_s[c] = R.beta[c] - mean * sc;
        if (blockIdx.x == 0) {
            const double m = R.stats[c] / cnt;
            double var = R.stats[kStatStride + c] / cnt - m * m;
            var = var < 0.0 ? 0.0 : var;
            R.rmean[c] = (1.f - T_BN_MOM) * R.rmean[c] + T_BN_MOM * ((float)m + R.bias[c]);
            R.rvar[c] = (1.f - T_BN_MOM) * R.rvar[c] + T_BN_MOM * (float)(var * cnt / (cnt > 1.0 ? cnt - 1.0 : 1.0));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && R.nbt) *R.nbt += 1;
    __syncthreads();
    const int kch = C / 8, So = POOL ? S / 2 : S;
    const int64_t total = N * kch * So * So;
    const uint4* z4 = reinterpret_cast<const uint4*>(Z);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int wo = (int)(t % So);
        int64_t r = t / So;
        const int ho = (int)(r % So); r /= So;
        const int ch = (int)(r % kch);
        const int64_t n = r / kch;
        const float* sc = sc_s + ch * 8;
        const float* sh = sh_s + ch * 8;
        float best[8];
        const int reps = POOL ? 2 : 1;
#pragma unroll
        for (int oy = 0; oy < reps; ++oy)
#pragma unroll
            for (int ox = 0; ox < reps; ++ox) {
                const int h = POOL ? 2 * ho + oy : ho, w = POOL ? 2 * wo + ox : wo;
                float z[8];
                unpack8(__ldg(z4 + ((n * kch + ch) * S + h) * S + w), z);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float y = fmaxf(fmaf(z[j], sc[j], sh[j]), 0.f);
                    best[j] = (oy | ox) ? fmaxf(best[j], y) : y;
                }
            }
        o4[((n * out_chunks + out_chunk0 + ch) * So + ho) * So + wo] = pack8(best);
    }
}

// fusion block tail: BN + ReLU + global average pool.  One warp per (patch, channel chunk).
__global__ void __launch_bounds__(256) bn_gap_kernel(const __nv_bfloat16* __restrict__ Z, BnRefs R, int C, int S, int64_t N, float* __restrict__ g) {
    __shared__ float sc_s[kStatStride], sh_s[kStatStride];
    const int npx = S * S;
    const double cnt = (double)N * npx;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, invstd;
        bn_coeffs(R.stats, c, cnt, mean, invstd);
        const float sc = R.gamma[c] * invstd;
        sc_s[c] = sc;
        sh_s[c] = R.beta[c] - mean * sc;
        if (blockIdx.x == 0) {
            const double m = R.stats[c] / cnt;
            double var = R.stats[kStatStride + c] / cnt - m * m;
            var = var < 0.0 ? 0.0 : var;
            R.rmean[c] = (1.f - T_BN_MOM) * R.rmean[c] + T_BN_MOM * ((float)m + R.bias[c]);
            R.rvar[c] = (1.f - T_BN_MOM) * R.rvar[c] + T_BN_MOM * (float)(var * cnt / (cnt > 1.0 ? cnt - 1.0 : 1.0));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && R.nbt) *R.nbt += 1;
    __syncthreads();
    const int kch = C / 8, lane = threadIdx.x & 31;
    const int64_t items = N * kch;
    const uint4* z4 = reinterpret_cast<const uint4*>(Z);
    const float inv = 1.f / (float)npx;
    for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < items; it += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int ch = (int)(it % kch);
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int px = lane; px < npx; px += 32) {
            float z[8];
            unpack8(__ldg(z4 + it * npx + px), z);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += fmaxf(fmaf(z[j], sc_s[ch * 8 + j], sh_s[ch * 8 + j]), 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int o = 16; o; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        float mine = s[0];                        // s[] is uniform across lanes after the butterfly: lane j keeps s[j]
#pragma unroll
        for (int j = 1; j < 8; ++j) mine = lane == j ? s[j] : mine;
        if (lane < 8) g[(it / kch) * C + ch * 8 + lane] = mine * inv;
    }
}

// ------------------------------------------------------------------------------------ BatchNorm backward
// SRC: where dL/d(activation after ReLU[/pool]) comes from.
//   0 POOLED: dout [N][dchunks][S/2][S/2][8] bf16 (chunk offset dchunk0): routed to the first maximum of each 2x2 window
//   1 DIRECT: dout [N][dchunks][S][S][8] bf16
//   2 GAP   : dg [N][C] fp32, spread evenly over the S*S pixels
// APPLY = false: accumulate sum dy and sum dy*z per channel (stats rows 2, 3).
// APPLY = true : dZ = gamma*invstd * (dy - mean(dy) - xhat * mean(dy*xhat)), written in Z's layout; block (0, ch) also
//                adds the affine gradients dgamma = sum dy*xhat, dbeta = sum dy.
template <int SRC, bool APPLY>
__global__ void __launch_bounds__(256) bn_bwd_kernel(const __nv_bfloat16* __restrict__ Z, const void* __restrict__ dsrc, int dchunks, int dchunk0,
                                                     BnRefs R, int C, int S, int64_t N, __nv_bfloat16* __restrict__ dZ) {
    const int ch = blockIdx.y, kch = C / 8;
    const double cnt = (double)N * S * S;
    __shared__ float sc_s[8], sh_s[8], a_s[8], b_s[8], c_s[8];
    __shared__ float red_s[8][16];
    if (threadIdx.x < 8) {
        const int c = ch * 8 + threadIdx.x;
        float mean, invstd;
        bn_coeffs(R.stats, c, cnt, mean, invstd);
        const float gam = R.gamma[c];
        sc_s[threadIdx.x] = gam * invstd;
        sh_s[threadIdx.x] = R.beta[c] - mean * gam * invstd;
        if (APPLY) {
            const double sdy = R.stats[2 * kStatStride + c], sdyz = R.stats[3 * kStatStride + c];
            const double dgam = (sdyz - (double)mean * sdy) * (double)invstd;       // sum dy * xhat
            // dz = a*dy + b*z + c0 with a = gamma*invstd, b = -a*invstd*dgam/cnt, c0 = -a*sdy/cnt - b*mean
            const double a = (double)gam * invstd;
            const double b = -a * (double)invstd * dgam / cnt;
            a_s[threadIdx.x] = (float)a;
            b_s[threadIdx.x] = (float)b;
            c_s[threadIdx.x] = (float)(-a * sdy / cnt - b * (double)mean);
            if (blockIdx.x == 0) {
                R.dgamma[c] += (float)dgam;
                R.dbeta[c] += (float)sdy;
            }
        }
    }
    __syncthreads();
    const int So = SRC == 0 ? S / 2 : S;
    const int64_t items = N * So * So;           // per channel chunk
    const uint4* z4 = reinterpret_cast<const uint4*>(Z);
    uint4* dz4 = reinterpret_cast<uint4*>(dZ);
    float s_dy[8], s_dyz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_dy[j] = 0.f; s_dyz[j] = 0.f; }
    const float gap_inv = 1.f / (float)(S * S);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < items; t += (int64_t)gridDim.x * blockDim.x) {
        const int wo = (int)(t % So);
        const int64_t r = t / So;
        const int ho = (int)(r % So);
        const int64_t n = r / So;
        float d[8];
        if (SRC == 2) {
            const float* dg = reinterpret_cast<const float*>(dsrc) + n * C + ch * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = dg[j] * gap_inv;
        } else {
            unpack8(__ldg(reinterpret_cast<const uint4*>(dsrc) + ((n * dchunks + dchunk0 + ch) * So + ho) * So + wo), d);
        }
        if (SRC == 0) {
            float z[4][8];
            const int64_t base = ((n * kch + ch) * S + 2 * ho) * S + 2 * wo;
            unpack8(__ldg(z4 + base), z[0]);
            unpack8(__ldg(z4 + base + 1), z[1]);
            unpack8(__ldg(z4 + base + S), z[2]);
            unpack8(__ldg(z4 + base + S + 1), z[3]);
            float o[4][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float best = fmaf(z[0][j], sc_s[j], sh_s[j]);
                int bi = 0;
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    const float y = fmaf(z[k][j], sc_s[j], sh_s[j]);
                    if (y > best) { best = y; bi = k; }
                }
                const float dy = best > 0.f ? d[j] : 0.f;
                if (!APPLY) {
                    const float zb = bi == 0 ? z[0][j] : bi == 1 ? z[1][j] : bi == 2 ? z[2][j] : z[3][j];
                    s_dy[j] += dy;
                    s_dyz[j] = fmaf(dy, zb, s_dyz[j]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k][j] = fmaf(a_s[j], k == bi ? dy : 0.f, fmaf(b_s[j], z[k][j], c_s[j]));
                }
            }
            if (APPLY) {
                dz4[base] = pack8(o[0]);
                dz4[base + 1] = pack8(o[1]);
                dz4[base + S] = pack8(o[2]);
                dz4[base + S + 1] = pack8(o[3]);
            }
        } else {
            float z[8], o[8];
            const int64_t base = ((n * kch + ch) * S + ho) * S + wo;
            unpack8(__ldg(z4 + base), z);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dy = fmaf(z[j], sc_s[j], sh_s[j]) > 0.f ? d[j] : 0.f;
                if (!APPLY) {
                    s_dy[j] += dy;
                    s_dyz[j] = fmaf(dy, z[j], s_dyz[j]);
                } else {
                    o[j] = fmaf(a_s[j], dy, fmaf(b_s[j], z[j], c_s[j]));
                }
            }
            if (APPLY) dz4[base] = pack8(o);
        }
    }
    if (!APPLY) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                s_dy[j] += __shfl_xor_sync(0xffffffffu, s_dy[j], o);
                s_dyz[j] += __shfl_xor_sync(0xffffffffu, s_dyz[j], o);
            }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { red_s[warp][j] = s_dy[j]; red_s[warp][8 + j] = s_dyz[j]; }
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            float tot = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red_s[w][threadIdx.x];
            const int j = threadIdx.x & 7;
            atomicAdd(R.stats + (threadIdx.x < 8 ? 2 : 3) * kStatStride + ch * 8 + j, (double)tot);
        }
    }
}

// ------------------------------------------------------------------------------------ head (fp32 CUDA cores)
// forward: g [N][128] -> hid = relu(fc1 g + b1) [N][64] -> logits = fc2 hid + b2 [N][C].  One warp per patch.
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ g, int64_t N, int C, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                       float* __restrict__ hid, float* __restrict__ logits) {
    extern __shared__ __align__(16) float hs[];
    float* w1t = hs;                          // [128][64]
    float* w2t = w1t + T_FUSE * T_HID;        // [64][C]
    float* gb = w2t + T_HID * C;              // per warp g[128] + hid[64]
    for (int i = threadIdx.x; i < T_FUSE * T_HID; i += blockDim.x) w1t[(i % T_FUSE) * T_HID + i / T_FUSE] = w1[i];
    for (int i = threadIdx.x; i < T_HID * C; i += blockDim.x) w2t[(i % T_HID) * C + i / T_HID] = w2[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float* gw = gb + warp * (T_FUSE + T_HID);
    float* hw = gw + T_FUSE;
    for (int64_t n = (int64_t)blockIdx.x * wpb + warp; n < N; n += (int64_t)gridDim.x * wpb) {
        *reinterpret_cast<float4*>(gw + 4 * lane) = __ldg(reinterpret_cast<const float4*>(g + n * T_FUSE) + lane);
        __syncwarp();
        float h0 = b1[lane], h1 = b1[lane + 32];
#pragma unroll 8
        for (int k = 0; k < T_FUSE; ++k) {
            h0 = fmaf(gw[k], w1t[k * T_HID + lane], h0);
            h1 = fmaf(gw[k], w1t[k * T_HID + lane + 32], h1);
        }
        h0 = fmaxf(h0, 0.f); h1 = fmaxf(h1, 0.f);
        hw[lane] = h0; hw[lane + 32] = h1;
        hid[n * T_HID + lane] = h0; hid[n * T_HID + lane + 32] = h1;
        __syncwarp();
        for (int c = lane; c < C; c += 32) {
            float a = b2[c];
#pragma unroll 8
            for (int k = 0; k < T_HID; ++k) a = fmaf(hw[k], w2t[k * C + c], a);
            logits[n * C + c] = a;
        }
        __syncwarp();
    }
}

// backward through the two linears for the activations: dhid = (W2^T dlogits) * (hid > 0), dg = W1^T dhid
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ hid, int64_t N, int C,
                                                       const float* __restrict__ w1, const float* __restrict__ w2,
                                                       float* __restrict__ dhid, float* __restrict__ dg) {
    extern __shared__ __align__(16) float hs[];
    float* w1s = hs;                          // [64][128] as stored
    float* w2s = w1s + T_HID * T_FUSE;        // [C][64] as stored
    float* buf = w2s + C * T_HID;             // per warp dlogits[64] + dhid[64]
    for (int i = threadIdx.x; i < T_HID * T_FUSE; i += blockDim.x) w1s[i] = w1[i];
    for (int i = threadIdx.x; i < C * T_HID; i += blockDim.x) w2s[i] = w2[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float* dl = buf + warp * 128;
    float* dh = dl + 64;
    for (int64_t n = (int64_t)blockIdx.x * wpb + warp; n < N; n += (int64_t)gridDim.x * wpb) {
        for (int c = lane; c < C; c += 32) dl[c] = dlogits[n * C + c];
        __syncwarp();
        float a0 = 0.f, a1 = 0.f;
        for (int c = 0; c < C; ++c) {
            a0 = fmaf(dl[c], w2s[c * T_HID + lane], a0);
            a1 = fmaf(dl[c], w2s[c * T_HID + lane + 32], a1);
        }
        a0 = hid[n * T_HID + lane] > 0.f ? a0 : 0.f;
        a1 = hid[n * T_HID + lane + 32] > 0.f ? a1 : 0.f;
        dh[lane] = a0; dh[lane + 32] = a1;
        dhid[n * T_HID + lane] = a0; dhid[n * T_HID + lane + 32] = a1;
        __syncwarp();
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int k = 0; k < T_HID; ++k) {
            const float v = dh[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fmaf(v, w1s[k * T_FUSE + lane + 32 * i], o[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) dg[n * T_FUSE + lane + 32 * i] = o[i];
        __syncwarp();
    }
}

// weight / bias gradients of the linears: block x < 64 -> row x of dW1 (= sum_n dhid[n][x] g[n][:]) and db1[x];
// block 64 + c -> row c of dW2 (= sum_n dlogits[n][c] hid[n][:]) and db2[c].  The batch is split over blockIdx.y
// (partials combined with atomics); 128 threads, 4 independent accumulators per thread.
__global__ void __launch_bounds__(128) head_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ hid, const float* __restrict__ dhid,
                                                         const float* __restrict__ dlogits, int64_t N, int C, float* __restrict__ dw1,
                                                         float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
    const bool first = blockIdx.x < T_HID;
    const int row = first ? blockIdx.x : blockIdx.x - T_HID;
    const int width = first ? T_FUSE : T_HID, ld = first ? T_HID : C;
    const float* coef = first ? dhid : dlogits;      // [N][ld], column `row`
    const float* act = first ? g : hid;              // [N][width]
    const int64_t per = (N + gridDim.y - 1) / gridDim.y, n0 = blockIdx.y * per, n1 = n0 + per < N ? n0 + per : N;
    const int j = threadIdx.x < width ? threadIdx.x : width - 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, bsum = 0.f;
    int64_t n = n0;
    for (; n + 4 <= n1; n += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float c = __ldg(coef + (n + u) * ld + row);
            acc[u] = fmaf(c, __ldg(act + (n + u) * width + j), acc[u]);
            bsum += c;
        }
    }
    for (; n < n1; ++n) {
        const float c = __ldg(coef + n * ld + row);
        acc[0] = fmaf(c, __ldg(act + n * width + j), acc[0]);
        bsum += c;
    }
    if ((int)threadIdx.x < width) atomicAdd((first ? dw1 : dw2) + row * width + threadIdx.x, (acc[0] + acc[1]) + (acc[2] + acc[3]));
    if (threadIdx.x == 0) atomicAdd((first ? db1 : db2) + row, bsum);
}

// CrossEntropyLoss(reduction='mean') forward + gradient (utils/utils.py:28-29; solver/mainsolver.py:53).
// One warp per sample; target as float32 (the loaders' labels) or int64 (after .long()).
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ logits, const void* __restrict__ target, int target_is_i64,
                                                         int64_t N, int C, float* __restrict__ loss, float* __restrict__ dlogits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float invN = 1.f / (float)N;
    float lsum = 0.f;
    for (int64_t n = (int64_t)blockIdx.x * wpb + warp; n < N; n += (int64_t)gridDim.x * wpb) {
        const float l0 = lane < C ? logits[n * C + lane] : -INFINITY;
        const float l1 = lane + 32 < C ? logits[n * C + lane + 32] : -INFINITY;
        float mx = fmaxf(l0, l1);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e0 = lane < C ? expf(l0 - mx) : 0.f, e1 = lane + 32 < C ? expf(l1 - mx) : 0.f;
        float se = e0 + e1;
#pragma unroll
        for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const int64_t tg = target_is_i64 ? reinterpret_cast<const int64_t*>(target)[n] : (int64_t)reinterpret_cast<const float*>(target)[n];
        const float lt = __shfl_sync(0xffffffffu, tg < 32 ? l0 : l1, (int)(tg & 31));
        if (lane == 0) lsum += (logf(se) + mx - lt);
        if (dlogits) {
            const float inv = 1.f / se;
            if (lane < C) dlogits[n * C + lane] = (e0 * inv - (tg == lane ? 1.f : 0.f)) * invN;
            if (lane + 32 < C) dlogits[n * C + lane + 32] = (e1 * inv - (tg == lane + 32 ? 1.f : 0.f)) * invN;
        }
    }
    __shared__ float ls[8];
    if (lane == 0) ls[warp] = lsum;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float t = 0.f;
        for (int w = 0; w < wpb; ++w) t += ls[w];
        atomicAdd(loss, t * invN);
    }
}

// torch.optim.Adam (no weight decay, no amsgrad), same operation order as torch's single-tensor path.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   int64_t n, float step_size, float b1, float b2, float eps, float bc2_sqrt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

}  // namespace dmf

using namespace dmf;

// ======================================================================================== host side
struct Bound { float* param = nullptr; float* grad = nullptr; int64_t numel = 0; };

struct TLayer {
    LayerGeom g{};             // forward geometry
    LayerGeom gd{};            // dgrad geometry (channel roles exchanged)
    __nv_bfloat16 *w = nullptr, *wd = nullptr;
    CUtensorMap map_in{};      // halo / 1x1 tile of the layer input (forward A operand, wgrad B operand)
    CUtensorMap map_dz_halo{}; // halo / 1x1 tile of dZ (dgrad A operand)
    CUtensorMap map_dz{};      // dense tile of dZ (wgrad A operand)
    int wg_stage = 0, wg_smem = 0;
    float* wg_scratch = nullptr;   // [taps][cin][cout] fp32 partial sums of wgrad_tc_kernel
};

struct dmf_train {
    int p = 0, C = 0, NB = 0;
    std::map<std::string, Bound> b;
    bool ready = false;
    int swap_lbo_sbo = 0;
    int64_t N = 0;             // batch of the last forward
    int64_t maps_N = -1;       // patch count the tensor maps were encoded for
    TLayer L[L_COUNT];
    BnRefs bn[L_COUNT];
    float *fc1w = nullptr, *fc1b = nullptr, *fc2w = nullptr, *fc2b = nullptr;
    float *dfc1w = nullptr, *dfc1b = nullptr, *dfc2w = nullptr, *dfc2b = nullptr;
    float *pan1w = nullptr, *dpan1w = nullptr;
    float* dconvw[L_COUNT] = {};
    double* stats = nullptr;   // [L_COUNT][4][kStatStride]
    float* wg_all = nullptr;   // all wgrad scratch tensors, one allocation (one memset per step)
    size_t wg_all_floats = 0;
    float *in_ms = nullptr, *in_pan = nullptr, *in_tgt = nullptr;   // staged patches when the batch comes from a scene
    const float* pan_patches = nullptr;                              // PAN input of the last forward (for the stem's wgrad)
    __nv_bfloat16 *X0 = nullptr, *A1 = nullptr, *B1 = nullptr, *B2 = nullptr, *CAT = nullptr;
    __nv_bfloat16* Z[L_COUNT] = {};
    __nv_bfloat16 *dZ = nullptr, *dA = nullptr, *dCAT = nullptr;
    float *g = nullptr, *hid = nullptr, *logits = nullptr, *dlogits = nullptr, *dhid = nullptr, *dg = nullptr, *loss = nullptr;
    std::map<std::string, std::pair<void*, size_t>> bufs;
};

namespace dmf {

static int S_of(const dmf_train* t, int layer) {
    const int p = t->p;
    switch (layer) {
        case L_MS1: case L_MS2: case L_PAN3: return p;
        case L_PAN1: return 4 * p;
        case L_PAN2: return 2 * p;
        default: return p / 2;
    }
}

template <typename T>
static int dalloc(dmf_train* t, const char* name, T** ptr, size_t count) {
    DMF_CUDA(cudaMalloc(ptr, count * sizeof(T)));
    t->bufs[name] = {(void*)*ptr, count * sizeof(T)};
    return DMF_OK;
}

template <int CI, int CO, int TAPS, int G, int NP, int MODE>
static int launch_raw(const LayerGeom& g, const CUtensorMap& map, const __nv_bfloat16* w, __nv_bfloat16* out, int out_chunks,
                      int64_t N, double* stats, cudaStream_t st) {
    tc::ConvParams P{};
    P.S = g.S; P.S_l2 = g.S_l2; P.NP = g.NP; P.NP_l2 = g.NP_l2; P.TH = g.TH; P.tiles_x_l2 = g.tiles_x_l2;
    P.PX = g.PX; P.PX_l2 = g.PX_l2; P.tpg_l2 = g.tpg_l2;
    P.N = (int)N;
    P.n_tiles = (int)((N + g.NP - 1) / g.NP) * g.tiles_per_group;
    P.a_plane = g.a_plane; P.a_stage = g.a_stage; P.n_stage = g.n_stage; P.sbo_a = g.sbo_a;
    P.out_chunks = out_chunks; P.out_chunk0 = 0;
    P.w = w; P.out = out; P.stats = stats; P.stat_stride = kStatStride;
    if (TAPS == 9 && (g.NP != NP || g.TH != 16 / NP)) { set_error("train conv geometry/template mismatch"); return DMF_ERR_STATE; }
    auto kern = tc::conv_tc_kernel<CI, CO, TAPS, false, G, NP, false, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        DMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    kern<<<std::min(P.n_tiles, num_sms()), 64 + 128 * G, g.smem, st>>>(map, P);
    DMF_LAUNCHED();
    return DMF_OK;
}

template <int CO, int CI, int TAPS, int ROLES, int NP>
static int launch_wgrad(const dmf_train* t, const TLayer& L, int64_t N, cudaStream_t st) {
    const LayerGeom& g = L.g;
    tc::WgradParams P{};
    P.n_tiles = (int)((N + g.NP - 1) / g.NP) * g.tiles_per_group;
    P.tpg_l2 = g.tpg_l2; P.tiles_x_l2 = g.tiles_x_l2; P.NP_l2 = g.NP_l2; P.PX_l2 = g.PX_l2;
    P.n_stage = L.wg_stage; P.swap_lbo_sbo = t->swap_lbo_sbo; P.dw = L.wg_scratch;
    if (TAPS == 9 && g.NP != NP) { set_error("train wgrad geometry/template mismatch"); return DMF_ERR_STATE; }
    auto kern = tc::wgrad_tc_kernel<CO, CI, TAPS, ROLES, NP>;
    static bool attr_set = false;
    if (!attr_set) {
        DMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    const int per_role = std::max(1, std::min(P.n_tiles, num_sms() / ROLES));
    kern<<<per_role * ROLES, 192, L.wg_smem, st>>>(L.map_dz, L.map_in, P);
    DMF_LAUNCHED();
    return DMF_OK;
}

static void wgrad_geom(TLayer& L) {
    const LayerGeom& g = L.g;
    const int dz_tile = g.cout / 8 * 2048, stage = dz_tile + g.a_stage;
    L.wg_stage = std::max(1, std::min(6, (kSmemLimit - dz_tile - 256) / stage));
    L.wg_smem = L.wg_stage * stage + dz_tile + 256;
}

// ---- per-layer dispatch (NP = 2 only for 8 x 8 maps, i.e. p = 8)
static int fwd_conv(dmf_train* t, int layer, int64_t N, cudaStream_t st) {
    TLayer& L = t->L[layer];
    double* s = t->bn[layer].stats;
    const bool np2 = L.g.NP == 2;
    switch (layer) {
        case L_MS1: return np2 ? launch_raw<16, T_MS1, 9, 3, 2, 2>(L.g, L.map_in, L.w, t->Z[layer], T_MS1 / 8, N, s, st)
                               : launch_raw<16, T_MS1, 9, 3, 1, 2>(L.g, L.map_in, L.w, t->Z[layer], T_MS1 / 8, N, s, st);
        case L_MS2: case L_PAN3:
            return np2 ? launch_raw<64, 128, 9, 3, 2, 2>(L.g, L.map_in, L.w, t->Z[layer], 16, N, s, st)
                       : launch_raw<64, 128, 9, 3, 1, 2>(L.g, L.map_in, L.w, t->Z[layer], 16, N, s, st);
        case L_PAN2: return launch_raw<T_PAN1, T_PAN2, 9, 3, 1, 2>(L.g, L.map_in, L.w, t->Z[layer], T_PAN2 / 8, N, s, st);
        case L_FUSE: return launch_raw<T_CAT, T_FUSE, 1, 2, 1, 2>(L.g, L.map_in, L.w, t->Z[layer], T_FUSE / 8, N, s, st);
    }
    return DMF_ERR_ARG;
}

static int dgrad_conv(dmf_train* t, int layer, __nv_bfloat16* out, int64_t N, cudaStream_t st) {
    TLayer& L = t->L[layer];
    const bool np2 = L.gd.NP == 2;
    switch (layer) {
        case L_MS2: case L_PAN3:
            return np2 ? launch_raw<128, 64, 9, 3, 2, 1>(L.gd, L.map_dz_halo, L.wd, out, 8, N, nullptr, st)
                       : launch_raw<128, 64, 9, 3, 1, 1>(L.gd, L.map_dz_halo, L.wd, out, 8, N, nullptr, st);
        case L_PAN2: return launch_raw<T_PAN2, T_PAN1, 9, 3, 1, 1>(L.gd, L.map_dz_halo, L.wd, out, T_PAN1 / 8, N, nullptr, st);
        case L_FUSE: return launch_raw<T_FUSE, T_CAT, 1, 2, 1, 1>(L.gd, L.map_dz_halo, L.wd, out, T_CAT / 8, N, nullptr, st);
    }
    return DMF_ERR_ARG;
}

static int wgrad_conv(dmf_train* t, int layer, int64_t N, cudaStream_t st) {
    TLayer& L = t->L[layer];
    const bool np2 = L.g.NP == 2;
    switch (layer) {
        case L_MS1: return np2 ? launch_wgrad<T_MS1, 16, 9, 1, 2>(t, L, N, st) : launch_wgrad<T_MS1, 16, 9, 1, 1>(t, L, N, st);
        case L_MS2: case L_PAN3:
            return np2 ? launch_wgrad<128, 64, 9, 3, 2>(t, L, N, st) : launch_wgrad<128, 64, 9, 3, 1>(t, L, N, st);
        case L_PAN2: return launch_wgrad<T_PAN2, T_PAN1, 9, 1, 1>(t, L, N, st);
        case L_FUSE: return launch_wgrad<T_FUSE, T_CAT, 1, 1, 1>(t, L, N, st);
    }
    return DMF_ERR_ARG;
}

// scratch of the listed layers -> gradient tensors (one launch)
static int wgrad_finish(dmf_train* t, const int* layers, int count, cudaStream_t st) {
    FinishJobs J{};
    for (int i = 0; i < count; ++i) {
        const int l = layers[i];
        const LayerGeom& g = t->L[l].g;
        J.j[i] = FinishJob{t->L[l].wg_scratch, t->dconvw[l], g.cout, g.cin, g.taps, l == L_MS1 ? 4 : g.cin, l == L_MS1 ? 1 : 0};
    }
    J.count = count;
    wgrad_finish_kernel<<<dim3(72, count), 256, 0, st>>>(J);
    DMF_LAUNCHED();
    return DMF_OK;
}

// Blocks of `threads` threads of a kernel that are resident per SM (registers / shared memory), cached per kernel.  The grid-stride
// kernels below are launched as exactly ONE wave: with 4 blocks per SM requested and 3 resident (79 registers) the statistics pass of
// the BatchNorm backward ran 1.33 waves and took as long as 2; the PAN stem forward (188 registers, 1 block resident) ran 4.
template <typename K>
static int resident_blocks(K kernel, int threads) {
    static std::map<const void*, int> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const void* key = reinterpret_cast<const void*>(kernel);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) n = 1;
    cache[key] = n;
    return n;
}

// PAN stem kernels: one warp per (patch, 16-row segment, chunk, 32-column group)
template <typename K>
static int stem_grid(K kernel, int64_t N, int S) {
    const int64_t tasks = N * (S / kStemRows) * 4 * (S / 32);
    return (int)std::max<int64_t>(1, std::min<int64_t>((tasks + 7) / 8, (int64_t)num_sms() * resident_blocks(kernel, 256)));
}

static int ew_grid(int64_t items) { return (int)std::min<int64_t>((items + 255) / 256, (int64_t)num_sms() * 8); }

// A_out = pool(relu(bn(Z)))
static int bn_forward(dmf_train* t, int layer, bool pool, __nv_bfloat16* out, int out_chunks, int out_chunk0, int64_t N, cudaStream_t st) {
    const int C = kCout[layer], S = S_of(t, layer), So = pool ? S / 2 : S;
    const int grid = ew_grid(N * (C / 8) * So * So);
    if (pool) bn_apply_kernel<true><<<grid, 256, 0, st>>>(t->Z[layer], t->bn[layer], C, S, N, out, out_chunks, out_chunk0);
    else bn_apply_kernel<false><<<grid, 256, 0, st>>>(t->Z[layer], t->bn[layer], C, S, N, out, out_chunks, out_chunk0);
    DMF_LAUNCHED();
    return DMF_OK;
}

// dZ(layer) from the gradient of its activated (and pooled) output
template <int SRC>
static int bn_backward(dmf_train* t, int layer, const void* dsrc, int dchunks, int dchunk0, int64_t N, cudaStream_t st) {
    const int C = kCout[layer], S = S_of(t, layer), So = SRC == 0 ? S / 2 : S;
    const int64_t items = N * So * So;
    auto grid_of = [&](int resident) {
        return dim3((unsigned)std::max<int64_t>(1, std::min<int64_t>((items + 255) / 256, std::max(1, num_sms() * resident / (C / 8)))), C / 8);
    };
    bn_bwd_kernel<SRC, false><<<grid_of(resident_blocks(bn_bwd_kernel<SRC, false>, 256)), 256, 0, st>>>(t->Z[layer], dsrc, dchunks, dchunk0, t->bn[layer], C, S, N, t->dZ);
    DMF_LAUNCHED();
    bn_bwd_kernel<SRC, true><<<grid_of(resident_blocks(bn_bwd_kernel<SRC, true>, 256)), 256, 0, st>>>(t->Z[layer], dsrc, dchunks, dchunk0, t->bn[layer], C, S, N, t->dZ);
    DMF_LAUNCHED();
    return DMF_OK;
}

static size_t head_fwd_smem(int C) { return sizeof(float) * (T_FUSE * T_HID + T_HID * C + 8 * (T_FUSE + T_HID)); }
static size_t head_bwd_smem(int C) { return sizeof(float) * (T_FUSE * T_HID + T_HID * C + 8 * 128); }

static int pack_all(dmf_train* t, cudaStream_t st) {
    PackJobs J{};
    int k = 0;
    auto add = [&](const float* w, __nv_bfloat16* out, int kind, int cin, int cout, int taps, int64_t n) {
        J.j[k].w = w; J.j[k].out = out; J.j[k].kind = kind; J.j[k].cin = cin; J.j[k].cout = cout; J.j[k].taps = taps; J.j[k].n = n; ++k;
    };
    auto W = [&](int layer) { return (const float*)t->b[std::string(kBlk[layer]) + ".0.weight"].param; };
    add(W(L_MS1), t->L[L_MS1].w, 2, 4, T_MS1, 9, 9 * 16 * T_MS1);
    const int ls[4] = {L_MS2, L_PAN2, L_PAN3, L_FUSE};
    for (int l : ls) {
        const LayerGeom& g = t->L[l].g;
        const int64_t n = (int64_t)g.taps * g.cin * g.cout;
        add(W(l), t->L[l].w, 0, g.cin, g.cout, g.taps, n);
        add(W(l), t->L[l].wd, 1, g.cin, g.cout, g.taps, n);
    }
    J.count = k;
    pack_weights_kernel<<<dim3(64, k), 256, 0, st>>>(J);
    DMF_LAUNCHED();
    return DMF_OK;
}

static int train_set_maps(dmf_train* t, int64_t N);

static int train_forward(dmf_train* t, const float* ms, const float* pan, int64_t N, cudaStream_t st) {
    const int p = t->p;
    t->N = N;
    DMF_TRY(train_set_maps(t, N));
    t->pan_patches = pan;
    DMF_CUDA(cudaMemsetAsync(t->stats, 0, sizeof(double) * L_COUNT * 4 * kStatStride, st));
    DMF_TRY(pack_all(t, st));
    // MS branch
    ms_split_kernel<<<ew_grid(N * p * p), 256, 0, st>>>(ms, p, N, t->X0);
    DMF_LAUNCHED();
    DMF_TRY(fwd_conv(t, L_MS1, N, st));
    DMF_TRY(bn_forward(t, L_MS1, false, t->A1, T_MS1 / 8, 0, N, st));
    DMF_TRY(fwd_conv(t, L_MS2, N, st));
    DMF_TRY(bn_forward(t, L_MS2, true, t->CAT, T_CAT / 8, 0, N, st));
    // PAN branch
    {
        const int S = 4 * p;
        pan1_fwd_kernel<<<stem_grid(pan1_fwd_kernel, N, S), 256, 0, st>>>(pan, t->pan1w, S, N, t->Z[L_PAN1], t->bn[L_PAN1].stats);
        DMF_LAUNCHED();
    }
    DMF_TRY(bn_forward(t, L_PAN1, true, t->B1, T_PAN1 / 8, 0, N, st));
    DMF_TRY(fwd_conv(t, L_PAN2, N, st));
    DMF_TRY(bn_forward(t, L_PAN2, true, t->B2, T_PAN2 / 8, 0, N, st));
    DMF_TRY(fwd_conv(t, L_PAN3, N, st));
    DMF_TRY(bn_forward(t, L_PAN3, true, t->CAT, T_CAT / 8, T_MS2 / 8, N, st));
    // fusion + head
    DMF_TRY(fwd_conv(t, L_FUSE, N, st));
    {
        const int grid = (int)std::min<int64_t>((N * (T_FUSE / 8) + 7) / 8, (int64_t)num_sms() * 8);
        bn_gap_kernel<<<grid, 256, 0, st>>>(t->Z[L_FUSE], t->bn[L_FUSE], T_FUSE, p / 2, N, t->g);
        DMF_LAUNCHED();
    }
    head_fwd_kernel<<<(int)std::min<int64_t>((N + 7) / 8, num_sms()), 256, head_fwd_smem(t->C), st>>>(
        t->g, N, t->C, t->fc1w, t->fc1b, t->fc2w, t->fc2b, t->hid, t->logits);
    DMF_LAUNCHED();
    return DMF_OK;
}

static int train_backward(dmf_train* t, const float* dlogits, cudaStream_t st) {
    const int64_t N = t->N;
    const int p = t->p;
    DMF_REQUIRE(N > 0, "train_backward: no forward to differentiate");
    DMF_CUDA(cudaMemsetAsync(t->wg_all, 0, sizeof(float) * t->wg_all_floats, st));
    head_bwd_kernel<<<(int)std::min<int64_t>((N + 7) / 8, num_sms()), 256, head_bwd_smem(t->C), st>>>(dlogits, t->hid, N, t->C, t->fc1w, t->fc2w,
                                                                                                       t->dhid, t->dg);
    DMF_LAUNCHED();
    head_wgrad_kernel<<<dim3(T_HID + t->C, (unsigned)std::max<int64_t>(1, std::min<int64_t>(16, N / 32))), 128, 0, st>>>(t->g, t->hid, t->dhid, dlogits, N, t->C, t->dfc1w, t->dfc1b, t->dfc2w, t->dfc2b);
    DMF_LAUNCHED();
    // fusion block
    DMF_TRY(bn_backward<2>(t, L_FUSE, t->dg, 0, 0, N, st));
    DMF_TRY(wgrad_conv(t, L_FUSE, N, st));
    DMF_TRY(dgrad_conv(t, L_FUSE, t->dCAT, N, st));
    // MS branch
    DMF_TRY(bn_backward<0>(t, L_MS2, t->dCAT, T_CAT / 8, 0, N, st));
    DMF_TRY(wgrad_conv(t, L_MS2, N, st));
    DMF_TRY(dgrad_conv(t, L_MS2, t->dA, N, st));
    DMF_TRY(bn_backward<1>(t, L_MS1, t->dA, T_MS1 / 8, 0, N, st));
    DMF_TRY(wgrad_conv(t, L_MS1, N, st));
    // PAN branch
    DMF_TRY(bn_backward<0>(t, L_PAN3, t->dCAT, T_CAT / 8, T_MS2 / 8, N, st));
    DMF_TRY(wgrad_conv(t, L_PAN3, N, st));
    DMF_TRY(dgrad_conv(t, L_PAN3, t->dA, N, st));
    DMF_TRY(bn_backward<0>(t, L_PAN2, t->dA, T_PAN2 / 8, 0, N, st));
    DMF_TRY(wgrad_conv(t, L_PAN2, N, st));
    DMF_TRY(dgrad_conv(t, L_PAN2, t->dA, N, st));
    DMF_TRY(bn_backward<0>(t, L_PAN1, t->dA, T_PAN1 / 8, 0, N, st));
    {
        const int S = 4 * p;
        pan1_wgrad_kernel<<<stem_grid(pan1_wgrad_kernel, N, S), 256, 0, st>>>(t->dZ, t->pan_patches, S, N, t->dpan1w);
        DMF_LAUNCHED();
    }
    const int all[5] = {L_MS1, L_MS2, L_PAN2, L_PAN3, L_FUSE};
    return wgrad_finish(t, all, 5, st);
}

}  // namespace dmf

namespace dmf {
// Tensor maps over the activation / gradient buffers for a batch of exactly N patches: layer inputs (forward + wgrad B operand)
// and dZ views (dgrad A operand, wgrad A operand).  The patch dimension of every map is N, not the capacity NB: tiles that hold
// several patches (8x8 maps: 2 per tile, the 1x1 layer up to 8) read ZEROS for the patch slots beyond the batch, so that the
// batch statistics and the weight gradients (sums over every pixel of a tile) never see stale workspace contents.
static int train_set_maps(dmf_train* t, int64_t N) {
    if (t->maps_N == N) return DMF_OK;
    const int ls[4] = {L_MS2, L_PAN2, L_PAN3, L_FUSE};
    DMF_TRY(make_map(&t->L[L_MS1].map_in, t->L[L_MS1].g, t->X0, N));
    DMF_TRY(make_map(&t->L[L_MS2].map_in, t->L[L_MS2].g, t->A1, N));
    DMF_TRY(make_map(&t->L[L_PAN2].map_in, t->L[L_PAN2].g, t->B1, N));
    DMF_TRY(make_map(&t->L[L_PAN3].map_in, t->L[L_PAN3].g, t->B2, N));
    DMF_TRY(make_map(&t->L[L_FUSE].map_in, t->L[L_FUSE].g, t->CAT, N));
    for (int l : ls) DMF_TRY(make_map(&t->L[l].map_dz_halo, t->L[l].gd, t->dZ, N));
    const int lw[5] = {L_MS1, L_MS2, L_PAN2, L_PAN3, L_FUSE};
    for (int l : lw) {
        LayerGeom gz = t->L[l].g;         // same tiling, channel count of the layer OUTPUT
        gz.cin = gz.cout;
        DMF_TRY(make_map(&t->L[l].map_dz, gz, t->dZ, N, false));
    }
    t->maps_N = N;
    return DMF_OK;
}
}  // namespace dmf

extern "C" {

int dmf_train_create(dmf_train** out, int p, int num_classes, int max_batch) {
    DMF_REQUIRE(out, "train_create: null");
    DMF_REQUIRE(p == 8 || p == 16 || p == 32, "train_create: patch_size must be 8, 16 or 32 (got %d)", p);
    DMF_REQUIRE(num_classes >= 2 && num_classes <= 64, "train_create: 2 <= Categories_Number <= 64");
    DMF_REQUIRE(max_batch >= 1 && max_batch <= (1 << 16), "train_create: bad max_batch");
    dmf_train* t = new dmf_train();
    t->p = p; t->C = num_classes; t->NB = max_batch;
    int rc = make_geom(t->L[L_MS1].g, p, 9, 16, T_MS1, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_MS2].g, p, 9, T_MS1, T_MS2, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_PAN2].g, 2 * p, 9, T_PAN1, T_PAN2, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_PAN3].g, p, 9, T_PAN2, T_PAN3, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_FUSE].g, p / 2, 1, T_CAT, T_FUSE, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_MS2].gd, p, 9, T_MS2, T_MS1, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_PAN2].gd, 2 * p, 9, T_PAN2, T_PAN1, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_PAN3].gd, p, 9, T_PAN3, T_PAN2, 0);
    if (rc == DMF_OK) rc = make_geom(t->L[L_FUSE].gd, p / 2, 1, T_FUSE, T_CAT, 0);
    if (rc != DMF_OK) { delete t; return rc; }
    const int ls[5] = {L_MS1, L_MS2, L_PAN2, L_PAN3, L_FUSE};
    for (int l : ls) wgrad_geom(t->L[l]);
    *out = t;
    return DMF_OK;
}

int dmf_train_destroy(dmf_train* t) {
    if (!t) return DMF_OK;
    for (auto& kv : t->bufs) cudaFree(kv.second.first);
    delete t;
    return DMF_OK;
}

int dmf_train_bind(dmf_train* t, const char* name, void* param_dev, float* grad_dev, int64_t numel) {
    DMF_REQUIRE(t && name && param_dev && numel > 0, "train_bind: bad argument");
    Bound b; b.param = (float*)param_dev; b.grad = grad_dev; b.numel = numel;
    t->b[name] = b;
    t->ready = false;
    return DMF_OK;
}

int dmf_train_set_debug(dmf_train* t, int swap_lbo_sbo) {
    DMF_REQUIRE(t, "train_set_debug: null");
    t->swap_lbo_sbo = swap_lbo_sbo;
    return DMF_OK;
}

static int need(dmf_train* t, const std::string& k, int64_t numel, bool grad, Bound** out) {
    auto it = t->b.find(k);
    if (it == t->b.end()) { set_error("train: tensor '%s' was not bound", k.c_str()); return DMF_ERR_STATE; }
    if (it->second.numel != numel) { set_error("train: tensor '%s' has %lld elements, expected %lld", k.c_str(), (long long)it->second.numel, (long long)numel); return DMF_ERR_STATE; }
    if (grad && !it->second.grad) { set_error("train: tensor '%s' needs a gradient buffer", k.c_str()); return DMF_ERR_STATE; }
    *out = &it->second;
    return DMF_OK;
}

int dmf_train_finalize(dmf_train* t) {
    DMF_REQUIRE(t, "train_finalize: null");
    const int p = t->p, C = t->C;
    const int cin[L_COUNT] = {4, T_MS1, 1, T_PAN1, T_PAN2, T_CAT};
    const bool first = t->stats == nullptr;
    if (first) DMF_TRY(dalloc(t, "stats", &t->stats, (size_t)L_COUNT * 4 * kStatStride));
    for (int l = 0; l < L_COUNT; ++l) {
        const std::string blk = kBlk[l];
        const int co = kCout[l], taps = l == L_FUSE ? 1 : 9;
        Bound *w, *cb, *gw, *gb, *rm, *rv;
        DMF_TRY(need(t, blk + ".0.weight", (int64_t)co * cin[l] * taps, true, &w));
        DMF_TRY(need(t, blk + ".0.bias", co, true, &cb));
        DMF_TRY(need(t, blk + ".1.weight", co, true, &gw));
        DMF_TRY(need(t, blk + ".1.bias", co, true, &gb));
        DMF_TRY(need(t, blk + ".1.running_mean", co, false, &rm));
        DMF_TRY(need(t, blk + ".1.running_var", co, false, &rv));
        BnRefs& R = t->bn[l];
        R.gamma = gw->param; R.beta = gb->param; R.bias = cb->param; R.dgamma = gw->grad; R.dbeta = gb->grad;
        R.rmean = rm->param; R.rvar = rv->param;
        auto it = t->b.find(blk + ".1.num_batches_tracked");
        R.nbt = it == t->b.end() ? nullptr : (int64_t*)it->second.param;
        R.stats = t->stats + (size_t)l * 4 * kStatStride;
        t->dconvw[l] = w->grad;
        if (l == L_PAN1) { t->pan1w = w->param; t->dpan1w = w->grad; }
    }
    Bound *a, *b1, *c, *d;
    DMF_TRY(need(t, "fc1.weight", (int64_t)T_HID * T_FUSE, true, &a));
    DMF_TRY(need(t, "fc1.bias", T_HID, true, &b1));
    DMF_TRY(need(t, "fc2.weight", (int64_t)C * T_HID, true, &c));
    DMF_TRY(need(t, "fc2.bias", C, true, &d));
    t->fc1w = a->param; t->dfc1w = a->grad; t->fc1b = b1->param; t->dfc1b = b1->grad;
    t->fc2w = c->param; t->dfc2w = c->grad; t->fc2b = d->param; t->dfc2b = d->grad;
    if (first) {
        const size_t NB = t->NB, pp = (size_t)p * p;
        DMF_TRY(dalloc(t, "in_ms", &t->in_ms, NB * 4 * pp));
        DMF_TRY(dalloc(t, "in_pan", &t->in_pan, NB * 16 * pp));
        DMF_TRY(dalloc(t, "in_tgt", &t->in_tgt, NB));
        DMF_TRY(dalloc(t, "X0", &t->X0, NB * 16 * pp));
        DMF_TRY(dalloc(t, "A1", &t->A1, NB * T_MS1 * pp));
        DMF_TRY(dalloc(t, "B1", &t->B1, NB * T_PAN1 * 4 * pp));
        DMF_TRY(dalloc(t, "B2", &t->B2, NB * T_PAN2 * pp));
        DMF_TRY(dalloc(t, "CAT", &t->CAT, NB * T_CAT * pp / 4));
        DMF_TRY(dalloc(t, "dCAT", &t->dCAT, NB * T_CAT * pp / 4));
        DMF_TRY(dalloc(t, "Z_ms1", &t->Z[L_MS1], NB * T_MS1 * pp));
        DMF_TRY(dalloc(t, "Z_ms2", &t->Z[L_MS2], NB * T_MS2 * pp));
        DMF_TRY(dalloc(t, "Z_pan1", &t->Z[L_PAN1], NB * T_PAN1 * 16 * pp));
        DMF_TRY(dalloc(t, "Z_pan2", &t->Z[L_PAN2], NB * T_PAN2 * 4 * pp));
        DMF_TRY(dalloc(t, "Z_pan3", &t->Z[L_PAN3], NB * T_PAN3 * pp));
        DMF_TRY(dalloc(t, "Z_fuse", &t->Z[L_FUSE], NB * T_FUSE * pp / 4));
        DMF_TRY(dalloc(t, "dZ", &t->dZ, NB * T_PAN1 * 16 * pp));
        DMF_TRY(dalloc(t, "dA", &t->dA, NB * T_PAN1 * 4 * pp));
        DMF_TRY(dalloc(t, "g", &t->g, NB * T_FUSE));
        DMF_TRY(dalloc(t, "hid", &t->hid, NB * T_HID));
        DMF_TRY(dalloc(t, "logits", &t->logits, NB * C));
        DMF_TRY(dalloc(t, "dlogits", &t->dlogits, NB * C));
        DMF_TRY(dalloc(t, "dhid", &t->dhid, NB * T_HID));
        DMF_TRY(dalloc(t, "dg", &t->dg, NB * T_FUSE));
        DMF_TRY(dalloc(t, "loss", &t->loss, 1));
        DMF_TRY(dalloc(t, "w_ms1", &t->L[L_MS1].w, (size_t)9 * 16 * T_MS1));
        const int ls[4] = {L_MS2, L_PAN2, L_PAN3, L_FUSE};
        for (int l : ls) {
            const LayerGeom& g = t->L[l].g;
            const size_t n = (size_t)g.taps * g.cin * g.cout;
            DMF_TRY(dalloc(t, (std::string("w_") + kBlk[l]).c_str(), &t->L[l].w, n));
            DMF_TRY(dalloc(t, (std::string("wd_") + kBlk[l]).c_str(), &t->L[l].wd, n));
        }
        {
            const int lw[5] = {L_MS1, L_MS2, L_PAN2, L_PAN3, L_FUSE};
            size_t tot = 0;
            for (int l : lw) tot += (size_t)t->L[l].g.taps * t->L[l].g.cin * t->L[l].g.cout;
            DMF_TRY(dalloc(t, "wg_scratch", &t->wg_all, tot));
            t->wg_all_floats = tot;
            size_t off = 0;
            for (int l : lw) { t->L[l].wg_scratch = t->wg_all + off; off += (size_t)t->L[l].g.taps * t->L[l].g.cin * t->L[l].g.cout; }
        }
        DMF_TRY(train_set_maps(t, t->NB));
        DMF_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        DMF_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    DMF_CUDA(cudaDeviceSynchronize());
    t->ready = true;
    return DMF_OK;
}

#define DMF_TRAIN_READY(t)                                                                            \
    do {                                                                                              \
        if (!(t) || !(t)->ready) { dmf::set_error("train: call dmf_train_finalize after binding all tensors"); return DMF_ERR_STATE; } \
    } while (0)

int dmf_train_forward(dmf_train* t, const float* ms_dev, const float* pan_dev, int64_t N, float* logits_out_dev, void* stream) {
    DMF_TRAIN_READY(t);
    DMF_REQUIRE(ms_dev && pan_dev && N >= 1 && N <= t->NB, "train_forward: batch must be 1..%d patches", t->NB);
    cudaStream_t st = (cudaStream_t)stream;
    DMF_TRY(train_forward(t, ms_dev, pan_dev, N, st));
    if (logits_out_dev) DMF_CUDA(cudaMemcpyAsync(logits_out_dev, t->logits, sizeof(float) * N * t->C, cudaMemcpyDeviceToDevice, st));
    return DMF_OK;
}

int dmf_train_backward(dmf_train* t, const float* dlogits_dev, void* stream) {
    DMF_TRAIN_READY(t);
    DMF_REQUIRE(dlogits_dev, "train_backward: null gradient");
    return train_backward(t, dlogits_dev, (cudaStream_t)stream);
}

int dmf_softmax_ce(const float* logits_dev, const void* target_dev, int target_is_i64, int64_t N, int C, float* loss_out_dev,
                   float* dlogits_out_dev, void* stream) {
    DMF_REQUIRE(logits_dev && target_dev && N >= 1 && C >= 2 && C <= 64, "softmax_ce: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (loss_out_dev) DMF_CUDA(cudaMemsetAsync(loss_out_dev, 0, sizeof(float), st));
    softmax_ce_kernel<<<(int)std::min<int64_t>((N + 7) / 8, 4 * num_sms()), 256, 0, st>>>(logits_dev, target_dev, target_is_i64, N, C, loss_out_dev,
                                                                                         dlogits_out_dev);
    DMF_LAUNCHED();
    return DMF_OK;
}

int dmf_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t numel, float lr,
                  float beta1, float beta2, float eps, int64_t step, void* stream) {
    DMF_REQUIRE(param_dev && grad_dev && exp_avg_dev && exp_avg_sq_dev && numel >= 0 && step >= 1, "adam_step: bad argument");
    if (numel == 0) return DMF_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<ew_grid(numel), 256, 0, (cudaStream_t)stream>>>(param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, numel, (float)((double)lr / bc1), beta1, beta2,
                                                                  eps, (float)sqrt(bc2));
    DMF_LAUNCHED();
    return DMF_OK;
}

/* one whole training step from patches: forward, CrossEntropyLoss(mean), backward into the bound gradient buffers
 * (accumulating — zero them first, like optimizer.zero_grad()).  loss_out_dev receives the batch loss. */
int dmf_train_step_patches(dmf_train* t, const float* ms_dev, const float* pan_dev, const void* target_dev, int target_is_i64, int64_t N,
                           float* loss_out_dev, void* stream) {
    DMF_TRAIN_READY(t);
    DMF_REQUIRE(ms_dev && pan_dev && target_dev && N >= 1 && N <= t->NB, "train_step: batch must be 1..%d patches", t->NB);
    cudaStream_t st = (cudaStream_t)stream;
    DMF_TRY(train_forward(t, ms_dev, pan_dev, N, st));
    DMF_TRY(dmf_softmax_ce(t->logits, target_dev, target_is_i64, N, t->C, loss_out_dev ? loss_out_dev : t->loss, t->dlogits, stream));
    return train_backward(t, t->dlogits, st);
}

/* the same, with the batch cropped from a device scene by flat pixel index (K1 fused in front): the PAN input is the
 * PAN window, or the IHS product's window when use_mspan != 0 (dataset_tri's third tensor, train/dataset.py:259-279). */
int dmf_train_step_scene(dmf_train* t, const dmf_scene* s, const int64_t* flat_idx_dev, int64_t N, int use_mspan, float* loss_out_dev,
                         void* stream) {
    DMF_TRAIN_READY(t);
    DMF_REQUIRE(s && flat_idx_dev && N >= 1 && N <= t->NB, "train_step_scene: batch must be 1..%d pixels", t->NB);
    DMF_REQUIRE(s->p == t->p, "train_step_scene: scene patch size %d != net patch size %d", s->p, t->p);
    DMF_REQUIRE(s->label, "train_step_scene: targets need dmf_scene_set_labels");
    DMF_REQUIRE(!use_mspan || s->mspan, "train_step_scene: use_mspan needs dmf_scene_set_mspan");
    if (use_mspan) {
        // tri gather: in_pan receives the MSPAN windows; the plain PAN windows go to dZ (free scratch at this point)
        DMF_TRY(dmf_gather(s, flat_idx_dev, N, t->in_ms, (float*)t->dZ, t->in_pan, t->in_tgt, stream));
    } else {
        DMF_TRY(dmf_gather(s, flat_idx_dev, N, t->in_ms, t->in_pan, nullptr, t->in_tgt, stream));
    }
    return dmf_train_step_patches(t, t->in_ms, t->in_pan, t->in_tgt, 0, N, loss_out_dev, stream);
}

/* test hook: run ONE stage on whatever the internal buffers hold (tests fill them through dmf_train_buffer):
 * op 0 = pack the bound weights, 1 = forward conv of `layer` (input buffer -> Z_<layer>, statistics), 2 = wgrad of `layer`
 * (dZ x input buffer -> bound gradient, accumulated), 3 = dgrad of `layer` (dZ -> dA, or dCAT for the fusion layer).
 * layer: 0 ms1, 1 ms2, 2 pan1 (ops 1, 2 only: CUDA-core stem on in_pan), 3 pan2, 4 pan3, 5 fuse. */
int dmf_train_debug_op(dmf_train* t, int op, int layer, int64_t N, void* stream) {
    DMF_TRAIN_READY(t);
    DMF_REQUIRE(op >= 0 && op <= 3 && layer >= 0 && layer < L_COUNT && N >= 1 && N <= t->NB, "train_debug_op: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (op == 0) return pack_all(t, st);
    DMF_TRY(train_set_maps(t, N));
    const int S = 4 * t->p;
    if (op == 1) {
        DMF_CUDA(cudaMemsetAsync(t->bn[layer].stats, 0, sizeof(double) * 4 * kStatStride, st));
        if (layer != L_PAN1) return fwd_conv(t, layer, N, st);
        pan1_fwd_kernel<<<stem_grid(pan1_fwd_kernel, N, S), 256, 0, st>>>(t->in_pan, t->pan1w, S, N, t->Z[L_PAN1], t->bn[L_PAN1].stats);
        DMF_LAUNCHED();
        return DMF_OK;
    }
    if (op == 2) {
        if (layer != L_PAN1) {
            DMF_CUDA(cudaMemsetAsync(t->wg_all, 0, sizeof(float) * t->wg_all_floats, st));
            DMF_TRY(wgrad_conv(t, layer, N, st));
            return wgrad_finish(t, &layer, 1, st);
        }
        pan1_wgrad_kernel<<<stem_grid(pan1_wgrad_kernel, N, S), 256, 0, st>>>(t->dZ, t->in_pan, S, N, t->dpan1w);
        DMF_LAUNCHED();
        return DMF_OK;
    }
    DMF_REQUIRE(layer == L_MS2 || layer == L_PAN2 || layer == L_PAN3 || layer == L_FUSE, "train_debug_op: layer %d has no dgrad", layer);
    return dgrad_conv(t, layer, layer == L_FUSE ? t->dCAT : t->dA, N, st);
}

/* test hook: device pointer + size of an internal buffer ("Z_ms2", "dZ", "A1", "logits", ...) */
int dmf_train_buffer(dmf_train* t, const char* name, void** ptr_out, int64_t* bytes_out) {
    DMF_REQUIRE(t && name && ptr_out && bytes_out, "train_buffer: bad argument");
    auto it = t->bufs.find(name);
    DMF_REQUIRE(it != t->bufs.end(), "train_buffer: no buffer named '%s'", name);
    *ptr_out = it->second.first;
    *bytes_out = (int64_t)it->second.second;
    return DMF_OK;
}

}  // extern "C"
