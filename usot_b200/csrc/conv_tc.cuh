// Interface of the tcgen05 implicit-GEMM convolution (conv_tc.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <vector>

namespace usot {

struct TcParams {
    CUtensorMap a[2][4];  // activation maps [plane hi/lo][stride-2 parity]
    CUtensorMap b[2];     // weight maps [plane]
    CUtensorMap o[2];     // output maps [plane] for the TMA-store epilogue (box {32 ch, bw, bh, 1}, 64B swizzle)
    CUtensorMap r[2];     // residual maps [plane], same box/swizzle as o[]: the residual chunk is TMA-loaded INTO the staging buffer
    int tma_res;          // 1: residual arrives by TMA (requires tma_store)
    int stages;           // operand ring depth used by this launch (<= the config's maximum)
    int nbuf;             // staging buffers per epilogue group: 1, or 3-4 when the residual is prefetched by TMA
    int res_ahead;        // residual chunks requested ahead of the one being processed (1 .. nbuf-1)
    int pdl;              // 1: launched with programmatic stream serialization; the kernel runs griddepcontrol.launch_dependents / .wait
    int l2_prefetch;      // 1: cp.async.bulk.prefetch.tensor hints for the next tile (residual chunks; A boxes of 1x1 layers)
    int tma_f32;          // 1: fp32-only output through the staging buffers + TMA store (o[0] is then an fp32 map); implies tma_store
    int fuse_cross;       // 1: split mode issues hi*[hi|lo] as one N = 2*BN MMA (main and cross accumulators are adjacent)
    int tma_store;        // 1: split-fp16 outputs leave through smem staging + cp.async.bulk.tensor stores
    int n_img, ho, wo, cout;
    int bw, bh, tiles_w, tiles_h;  // spatial patch of one M tile
    int hin, skip_pad_rows;        // input height; 1: one-row tiles skip the K-steps whose filter row lies in the zero padding
    int bimg;                      // images per M tile (bw*bh*bimg <= 128 rows; one TMA box {64 ch, bw, bh, bimg})
    int n_tiles_n, num_tiles;
    int num_pair_tiles;   // > 0: CTA-pair launch (cta_group::2, clusters of two CTAs): ceil(image groups / 2) * patches * N blocks
    int group;            // maps walked back to back per (patch, N block): 1, or N_q in the fused Conf_Fusion launch (image = sample * group + q)
    int fuse_cout;        // fused Conf_Fusion: channels of the fused output map (= cout / 2)
    int a_rank5;          // fused stem: the activation maps are 5-D {16 ch, 4 kx, ox, row, image} (fallback when the driver rejects the overlapping 4-D view)
    int pool_bands, pool_band_rows, pool_po;  // fused stem + max-pool: bands per image, pooled rows per band, pooled map size (0: off)
    int taps, kw, cin_chunks;
    int stride, ph, pw, dh, dw;
    const float* scale;
    const float* shift;
    const __half* res_hi;
    const __half* res_lo;
    __half* out_hi;
    __half* out_lo;
    float* out_f32;
    int relu;
};

struct TcTensor {  // NHWC fp16 planes of one activation tensor (lo may be null in single-fp16 mode)
    const __half* hi;
    const __half* lo;
};

struct TcWeights {  // [cout][K] K-major fp16 planes of w * 2^e[co]; scale already multiplied by 2^-e[co]
    const __half* hi;
    const __half* lo;
    const float* scale;
    int K;
};

struct TcEpilogue {
    const float* shift;
    const __half* res_hi;
    const __half* res_lo;
    __half* out_hi;
    __half* out_lo;
    float* out_f32;
    int relu;
};

// generic cuTensorMapEncodeTiled wrapper (dtype / swizzle are CUtensorMapDataType / CUtensorMapSwizzle values)
int encode_tmap(CUtensorMap* m, int dtype, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                const cuuint32_t* box, int swizzle);
extern Tunable g_tc_res_ahead, g_tc_skip_pad_rows, g_tc_multi_image_tiles, g_tc_bn_max, g_tc_split_bn_max, g_tc_tma_store, g_tc_tma_res, g_tc_fuse_cross, g_tc_tma_f32, g_tc_l2_prefetch, g_tc_latency_split, g_tc_pdl, g_tc_cta_pair;
// fuse_group > 0: fused Conf_Fusion launch (connect.py:123-144).  `w` holds conf_gen / value_gen interleaved in blocks of 64 output channels
// (rows [128 b, 128 b + 64) = conf channels [64 b, 64 b + 64), rows [128 b + 64, 128 b + 128) = the same value channels; scale / shift alike),
// g.n = samples * fuse_group maps, and ep.out_* receive the (samples, ho, wo, cout / 2) fused map.
int launch_conv_tc(const TcTensor& in, const ConvGeom& g, const TcWeights& w, const TcEpilogue& ep, bool split, cudaStream_t st, int fuse_group = 0);
int launch_f32_to_split(const float* in, size_t n, __half* hi, __half* lo, cudaStream_t st, const float* mul = nullptr);  // mul: device scalar
int launch_split_to_f32(const __half* hi, const __half* lo /*or null*/, size_t n, float* out, cudaStream_t st);  // out = hi + lo
int launch_pack_tc_weights(const float* w_kn, int K, int cout, const float* scale_in, __half* hi, __half* lo /*or null*/, float* scale_out,
                           cudaStream_t st);  // device-side twin of pack_tc_weights_host
void pack_tc_weights_host(const float* w_kn, int K, int cout, const float* scale_in, std::vector<__half>& hi, std::vector<__half>& lo,
                          std::vector<float>& scale_out);

// wgrad_tc.cu: weight gradient on tcgen05 (MN-major operands straight from the NHWC maps) + the power-of-two gradient scaling it needs
bool wgrad_tc_supported(const ConvGeom& g);
int launch_conv_wgrad_tc(const float* x, const float* dy, const ConvGeom& g, float* dw_kn, bool split, cudaStream_t st);
int launch_pow2_scale(const float* x, size_t n, int target_log2, float* y /*or null*/, float* out2 /*{s, 1/s}*/, cudaStream_t st);
// Stem + max-pool as a TMA-fed implicit GEMM over the space-to-depth image (conv_tc.cu).  stem_pool_bands() == 0: shape not supported.
int stem_pool_bands(int n, int s);
size_t stem_s2d_plane_elems(int n, int s);   // halves per s2d plane: n * HP * HP * 16 with HP = (s - 7) / 2 + 4
int launch_stem_s2d_weights(const float* stem_w /*[147][64]*/, const float* stem_scale, float* w_kn2_scratch /*[256][64]*/, __half* w_hi,
                            __half* w_lo /*[64][256]*/, float* scale_out /*[64]*/, cudaStream_t st);
int launch_stem_s2d_pool(const float* x_nchw, int n, int s, const __half* w_hi, const __half* w_lo, const float* scale_tc, const float* shift,
                         __half* s2d_hi, __half* s2d_lo /*scratch planes*/, __half* pool_hi, __half* pool_lo /*(n,PO,PO,64)*/, bool split,
                         cudaStream_t st);
size_t stem_tc_image_bytes();  // size of the packed stem weight tile (validated when a packed-weight image is imported)
void pack_stem_tc_host(const float* w_oihw, const float* scale_in, std::vector<uint8_t>& img, std::vector<float>& scale_out);

}  // namespace usot
