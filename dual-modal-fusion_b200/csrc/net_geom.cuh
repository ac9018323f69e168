// Layer geometry + TMA tensor maps shared by the inference network (net.cu) and the training path
// (train.cu): tile shapes of the implicit-GEMM kernels of conv_tc.cuh / wgrad_tc.cuh over the
// C8-planar activation layout [N][C/8][S][S][8] bf16.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "conv_tc.cuh"

namespace dmf {

constexpr int kSmemLimit = 227 * 1024;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct LayerGeom {
    int S, taps, cin, cout, pool;
    int NP, TH, tiles_x, tiles_y, PX, tiles_per_group, a_plane, a_stage, sbo_a, n_stage, smem;
    int S_l2, NP_l2, tiles_x_l2, PX_l2, tpg_l2;
    int rowpair;   // conv_rowpair_kernel: 32-row x 8-column tiles, N = 2 x cout
};

static inline int make_geom(LayerGeom& g, int S, int taps, int cin, int cout, int pool, int rowpair = 0) {
    g.S = S; g.taps = taps; g.cin = cin; g.cout = cout; g.pool = pool; g.rowpair = rowpair;
    const int kch = cin / 8;
    if (rowpair) {
        DMF_REQUIRE(taps == 9 && pool && cout == 64 && S % 32 == 0, "row-pair conv needs a pooled 3x3 layer, cout 64, map multiple of 32");
        g.TH = 32; g.NP = 1; g.tiles_x = S / 8; g.tiles_y = S / 32; g.PX = 0;
        g.tiles_per_group = g.tiles_x * g.tiles_y;
        g.sbo_a = 2 * tc::kPitch * 16;
        g.a_plane = (g.TH + 2) * tc::kPitch * 16;
    } else if (taps == 9) {
        DMF_REQUIRE(S % 8 == 0, "conv3x3 map size %d is not a multiple of 8", S);
        g.TH = (S % 16 == 0) ? 16 : 8;
        g.NP = 16 / g.TH;
        g.tiles_x = S / 8; g.tiles_y = S / g.TH; g.PX = 0;
        g.tiles_per_group = g.tiles_x * g.tiles_y;
        g.sbo_a = tc::kPitch * 16;
        g.a_plane = (g.TH + 2) * g.NP * tc::kPitch * 16;
    } else {
        const int px = S * S;
        DMF_REQUIRE(px % 128 == 0 || 128 % px == 0, "conv1x1 map %dx%d does not tile into 128 pixels", S, S);
        g.PX = std::min(px, 128); g.NP = 128 / g.PX; g.TH = 0; g.tiles_x = g.tiles_y = 0;
        g.tiles_per_group = px / g.PX;
        g.sbo_a = 128;
        g.a_plane = 128 * 16;
    }
    g.a_stage = kch * g.a_plane;
    DMF_REQUIRE(g.a_stage % 128 == 0, "A stage not 128-byte aligned");
    auto l2 = [](int v) { int e = 0; while ((1 << e) < v) ++e; return e; };
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    DMF_REQUIRE(pow2(S) && pow2(g.NP) && pow2(g.tiles_per_group) && (taps == 1 || pow2(g.tiles_x)) && (taps == 9 || pow2(g.PX)),
                "layer geometry must be powers of two");
    g.S_l2 = l2(S); g.NP_l2 = l2(g.NP); g.tiles_x_l2 = taps == 9 ? l2(g.tiles_x) : 0; g.PX_l2 = taps == 1 ? l2(g.PX) : 0;
    g.tpg_l2 = l2(g.tiles_per_group);
    const int fixed = (rowpair ? 12 * cin * 2 * cout * 2 : taps * cin * cout * 2) + 4 * cout * 4 + 256;   // weights, scale/shift, batch-statistics scratch, barriers
    g.n_stage = std::min(6, (kSmemLimit - fixed) / g.a_stage);
    DMF_REQUIRE(g.n_stage >= 1, "layer does not fit in shared memory");
    g.smem = fixed + g.n_stage * g.a_stage;
    return DMF_OK;
}

// tensor map over an activation tensor [N][C/8][S][S][8] bf16 for a layer's A-tile box
// halo = false: the dense TH x 8 tile without the 1-pixel border (the dZ operand of wgrad_tc.cuh)
static inline int make_map(CUtensorMap* m, const LayerGeom& g, const void* base, int64_t N, bool halo = true) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return DMF_ERR_CUDA; }
    const int kch = g.cin / 8, S = g.S;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], es[4] = {1, 1, 1, 1};
    if (g.taps == 9) {
        dims[0] = 8ull * S; dims[1] = (cuuint64_t)N; dims[2] = S; dims[3] = kch;
        strides[0] = (cuuint64_t)kch * S * S * 16; strides[1] = (cuuint64_t)S * 16; strides[2] = (cuuint64_t)S * S * 16;
        box[0] = halo ? 8 * tc::kPitch : 64; box[1] = g.NP; box[2] = halo ? g.TH + 2 : g.TH; box[3] = kch;
    } else {
        // inner dimension = 8 channels x IB pixels merged (contiguous in memory): 512-byte TMA rows
        const int IB = std::min(g.PX, 32);
        dims[0] = 8ull * IB; dims[1] = (cuuint64_t)S * S / IB; dims[2] = (cuuint64_t)N; dims[3] = kch;
        strides[0] = (cuuint64_t)IB * 16; strides[1] = (cuuint64_t)kch * S * S * 16; strides[2] = (cuuint64_t)S * S * 16;
        box[0] = 8 * IB; box[1] = g.PX / IB; box[2] = g.NP; box[3] = kch;
    }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return DMF_ERR_CUDA; }
    return DMF_OK;
}

}  // namespace dmf
