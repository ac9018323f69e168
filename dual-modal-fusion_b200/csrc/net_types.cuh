// Types shared by the per-patch network (net.cu) and the scene-dense path (dense.cu).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "net_geom.cuh"

namespace dmf {

constexpr int C_MS1 = 64, C_MS2 = 128, C_PAN1 = 32, C_PAN2 = 64, C_PAN3 = 128, C_CAT = 256, C_FUSE = 128, C_HID = 64;
constexpr float BN_EPS = 1e-5f;
constexpr int kHeadWarps = 8;

struct ConvLayer {
    LayerGeom g;
    __nv_bfloat16* w = nullptr;   // packed [tap][cin/8][cout][8]; fp16 bit patterns when f16 (tc::w16), in the 16-bit container type
    bool f16 = false;             // fp16 operands (every inference layer but the hi/lo-split MS stem)
    float* scale = nullptr;
    float* shift = nullptr;
    CUtensorMap map;              // over the workspace input buffer
};

struct DenseWs;                   // dense.cu: workspace + weights of the scene-dense path

}  // namespace dmf

struct dmf_net {
    int p = 0, C = 0, NB = 0;
    std::map<std::string, std::vector<float>> params;
    bool ready = false;
    bool timing = false;
    // stems / head weights
    __nv_bfloat16* w_pan1 = nullptr;                                    // hi/lo-split PAN stem weights [4][32][8]
    float *sc_pan1 = nullptr, *sh_pan1 = nullptr;
    float *fc1t = nullptr, *fc1b = nullptr, *fc2t = nullptr, *fc2b = nullptr;
    dmf::ConvLayer L[5];   // ms2, pan2, pan3, fuse, ms1 (hi/lo-split stem, 16 -> 64)
    __nv_bfloat16 *X0 = nullptr, *A1 = nullptr, *B1 = nullptr, *B2 = nullptr, *CAT = nullptr, *F = nullptr;
    float* gap = nullptr;     // [NB][128] per-patch channel sums when the pooling is fused into the fusion conv (p <= 16)
    cudaEvent_t ev[8] = {};
    float stage_ms[8] = {};
    dmf::DenseWs* dense = nullptr;
    int dense_band = 512;     // anchor rows per band of the dense path
    int dense_mode = 1;       // dmf_infer_scene: 1 = scene-dense maps, 0 = per-patch kernels
    int use_mspan = 0;        // scene inference reads the IHS product (dataset_tri's third raster) as the PAN input
};

namespace dmf {

const std::vector<float>* param(const dmf_net* n, const std::string& k, size_t numel);
// eval-mode BatchNorm folded with the conv bias: y = acc*scale + shift
int fold_bn(const dmf_net* n, const std::string& blk, int cout, std::vector<float>& scale, std::vector<float>& shift);

template <typename T>
static int to_device(T** dst, const std::vector<T>& v) {
    if (*dst) cudaFree(*dst);
    *dst = nullptr;
    DMF_CUDA(cudaMalloc(dst, sizeof(T) * v.size()));
    DMF_CUDA(cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return DMF_OK;
}

size_t head_smem(int C);

// dense.cu
int dense_pack(dmf_net* n);        // dense-only weights, after the per-patch ones are packed
void dense_release(dmf_net* n);
int dense_infer(dmf_net* n, const dmf_scene* s, int row0, int row1, float* logits_dev, uint8_t* pred_map_dev, int64_t* cm_dev, cudaStream_t st);

}  // namespace dmf
