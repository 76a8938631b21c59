// C ABI of the small stand-alone operators (declared in include/usot_b200.h): each is one reference op of the forward path exposed on
// its own so that tests can check it against the oracle function it replaces, not only end to end.
#include "common.cuh"
#include "conv_tc.cuh"
#include "../../include/usot_b200.h"

#include <vector>

using namespace usot;

extern "C" {

int usot_maxpool3x3s2p1_nhwc(const float* in, int n, int h, int w, int channels, float* out, float* out_split_sum, void* stream) {
    USOT_REQUIRE(n == 0 || (in && (out || out_split_sum)), "null pointer");
    USOT_REQUIRE(n >= 0 && h > 0 && w > 0 && channels > 0 && channels % 4 == 0, "bad shape");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    count_op_launch(OPFAM_POOL, 1);
    if (out)
        if (int rc = launch_maxpool3x3s2p1(in, n, h, w, channels, out, nullptr, nullptr, st)) return rc;
    if (out_split_sum) {
        // the engine's variant: the pooled map leaves as split-fp16 planes; reported here as hi + lo in fp32
        const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
        const size_t numel = (size_t)n * ho * wo * channels;
        __half *hi = nullptr, *lo = nullptr;
        USOT_CUDA_OK(cudaMalloc(&hi, numel * 2));
        if (cudaMalloc(&lo, numel * 2) != cudaSuccess) { cudaFree(hi); set_error("usot_maxpool3x3s2p1_nhwc: cudaMalloc failed"); return 1; }
        int rc = launch_maxpool3x3s2p1(in, n, h, w, channels, nullptr, hi, lo, st);
        if (!rc) rc = launch_split_to_f32(hi, lo, numel, out_split_sum, st);
        cudaStreamSynchronize(st);
        cudaFree(hi); cudaFree(lo);
        return rc;
    }
    return 0;
}

int usot_stem_conv(const float* x, int n, int size, const float* host_weight_oihw, const float* host_scale, const float* host_shift,
                   float* out, int precision, void* stream) {
    USOT_REQUIRE(x && host_weight_oihw && host_scale && host_shift && out, "null pointer");
    USOT_REQUIRE(n > 0 && size >= 7, "bad shape");
    USOT_REQUIRE(precision >= USOT_PREC_FP32_SIMT && precision <= USOT_PREC_FP16_TC, "unknown precision mode");
    cudaStream_t st = (cudaStream_t)stream;
    float *d_w = nullptr, *d_scale = nullptr, *d_shift = nullptr;
    int rc = 0;
    do {
        if (cudaMalloc(&d_scale, 64 * 4) || cudaMalloc(&d_shift, 64 * 4)) { set_error("usot_stem_conv: cudaMalloc failed"); rc = 1; break; }
        cudaMemcpyAsync(d_shift, host_shift, 64 * 4, cudaMemcpyHostToDevice, st);
        if (precision == USOT_PREC_FP32_SIMT) {
            std::vector<float> packed(147 * 64);
            for (int co = 0; co < 64; ++co)
                for (int k = 0; k < 147; ++k) packed[(size_t)k * 64 + co] = host_weight_oihw[(size_t)co * 147 + k];
            if (cudaMalloc(&d_w, packed.size() * 4)) { set_error("usot_stem_conv: cudaMalloc failed"); rc = 1; break; }
            cudaMemcpyAsync(d_w, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_scale, host_scale, 64 * 4, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);  // `packed` dies at the end of this scope
            rc = launch_stem(x, n, size, d_w, d_scale, d_shift, out, st);
        } else {
            std::vector<uint8_t> img;
            std::vector<float> scale_tc;
            pack_stem_tc_host(host_weight_oihw, host_scale, img, scale_tc);
            if (cudaMalloc(&d_w, img.size())) { set_error("usot_stem_conv: cudaMalloc failed"); rc = 1; break; }
            cudaMemcpyAsync(d_w, img.data(), img.size(), cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_scale, scale_tc.data(), 64 * 4, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            rc = launch_stem_tc(x, n, size, d_w, d_scale, d_shift, out, precision == USOT_PREC_FP16X3_TC, st);
        }
        if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("usot_stem_conv: ") + cudaGetErrorString(cudaGetLastError())); rc = 1; }
    } while (0);
    cudaFree(d_w); cudaFree(d_scale); cudaFree(d_shift);
    return rc;
}

int usot_stem_maxpool(const float* x, int n, int size, const float* host_weight_oihw, const float* host_scale, const float* host_shift,
                      float* out, int precision, void* stream) {
    USOT_REQUIRE(x && host_weight_oihw && host_scale && host_shift && out, "null pointer");
    USOT_REQUIRE(n > 0 && size >= 13, "bad shape");
    USOT_REQUIRE(precision == USOT_PREC_FP16X3_TC || precision == USOT_PREC_FP16_TC, "usot_stem_maxpool runs the tensor-core path (fp16x3 / fp16)");
    USOT_REQUIRE(stem_pool_bands(n, size) > 0, "usot_stem_maxpool: the conv map must fit one 128-pixel tile (size <= 261)");
    cudaStream_t st = (cudaStream_t)stream;
    const bool split = precision == USOT_PREC_FP16X3_TC;
    const int HO = (size - 7) / 2 + 1, PO = (HO - 1) / 2 + 1;
    const size_t n_pool = (size_t)n * PO * PO * 64, n_s2d = stem_s2d_plane_elems(n, size);
    std::vector<float> packed(147 * 64);
    for (int co = 0; co < 64; ++co)
        for (int k = 0; k < 147; ++k) packed[(size_t)k * 64 + co] = host_weight_oihw[(size_t)co * 147 + k];
    float *d_w = nullptr, *d_scale = nullptr, *d_shift = nullptr, *d_scratch = nullptr, *d_scale2 = nullptr;
    __half *w_hi = nullptr, *w_lo = nullptr, *s_hi = nullptr, *s_lo = nullptr, *p_hi = nullptr, *p_lo = nullptr;
    int rc = 0;
    do {
        if (cudaMalloc(&d_w, packed.size() * 4) || cudaMalloc(&d_scale, 256) || cudaMalloc(&d_shift, 256) || cudaMalloc(&d_scale2, 256) ||
            cudaMalloc(&d_scratch, 256 * 64 * 4) || cudaMalloc(&w_hi, 64 * 256 * 2) || cudaMalloc(&w_lo, 64 * 256 * 2) ||
            cudaMalloc(&s_hi, n_s2d * 2) || cudaMalloc(&s_lo, n_s2d * 2) || cudaMalloc(&p_hi, n_pool * 2) || cudaMalloc(&p_lo, n_pool * 2)) {
            set_error("usot_stem_maxpool: cudaMalloc failed"); rc = 1; break;
        }
        cudaMemcpyAsync(d_w, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_scale, host_scale, 256, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_shift, host_shift, 256, cudaMemcpyHostToDevice, st);
        if ((rc = launch_stem_s2d_weights(d_w, d_scale, d_scratch, w_hi, w_lo, d_scale2, st))) break;
        count_op_launch(OPFAM_STEM, 2);
        if ((rc = launch_stem_s2d_pool(x, n, size, w_hi, w_lo, d_scale2, d_shift, s_hi, split ? s_lo : nullptr, p_hi, split ? p_lo : nullptr, split, st))) break;
        if ((rc = launch_split_to_f32(p_hi, split ? p_lo : nullptr, n_pool, out, st))) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("usot_stem_maxpool: ") + cudaGetErrorString(cudaGetLastError())); rc = 1; }
    } while (0);
    cudaStreamSynchronize(st);
    cudaFree(d_w); cudaFree(d_scale); cudaFree(d_shift); cudaFree(d_scale2); cudaFree(d_scratch); cudaFree(w_hi); cudaFree(w_lo);
    cudaFree(s_hi); cudaFree(s_lo); cudaFree(p_hi); cudaFree(p_lo);
    return rc;
}

int usot_stem_conv_raw(const float* x, int n, int size, const float* weight_kn, float* out, void* stream) {
    USOT_REQUIRE(n == 0 || (x && weight_kn && out), "null pointer");
    USOT_REQUIRE(n >= 0 && size >= 7, "bad shape");
    if (n == 0) return 0;
    count_op_launch(OPFAM_STEM, 1);
    return launch_stem(x, n, size, weight_kn, nullptr, nullptr, out, (cudaStream_t)stream, /*relu=*/0);
}

int usot_conf_fusion(const float* conf, const float* value, int batch, int nq, int64_t per_map, float* out, void* stream) {
    USOT_REQUIRE(batch == 0 || (conf && value && out), "null pointer");
    USOT_REQUIRE(batch >= 0 && nq > 0 && per_map > 0 && per_map % 4 == 0, "bad shape");
    count_op_launch(OPFAM_FUSION, 1);
    return launch_conf_fusion(conf, value, batch, nq, (size_t)per_map, out, (cudaStream_t)stream);
}

int usot_cycle_glue(const float* off_cls, const float* mem_cls, const float* off_bbox, int n, int score_size, int search_size,
                    int search_feature_size, float cls_ratio, float* pool_box, float* best_score, int32_t* best_idx, void* stream) {
    USOT_REQUIRE(n == 0 || (off_cls && mem_cls && off_bbox && pool_box), "null pointer");
    USOT_REQUIRE(n >= 0 && score_size > 0 && search_size > 0 && search_feature_size > 1, "bad shape");
    return launch_cycle_glue(off_cls, mem_cls, off_bbox, n, score_size, search_size, search_feature_size, cls_ratio, pool_box, best_score,
                             best_idx, (cudaStream_t)stream);
}

int usot_weighted_bce(const float* pred, const float* label, int count, float* loss, void* stream) {
    USOT_REQUIRE(pred && label && loss && count > 0, "bad argument");
    return launch_bce(pred, label, count, loss, (cudaStream_t)stream);
}

int usot_iou_loss(const float* bbox, const float* reg_target, const float* reg_weight, int n, int cells, float* loss, void* stream) {
    USOT_REQUIRE(bbox && reg_target && reg_weight && loss && n > 0 && cells > 0, "bad argument");
    return launch_iou(bbox, reg_target, reg_weight, n, cells, loss, (cudaStream_t)stream);
}

}  // extern "C"
