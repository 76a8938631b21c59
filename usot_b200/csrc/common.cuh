// Shared declarations for the usot_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <string>

namespace usot {

// Process-wide performance knobs (usot_set_tunable) are read by launches on any host thread: atomics, never plain ints.
typedef std::atomic<int> Tunable;

// ---- error plumbing: kernels never exit(); launchers return cudaError_t-like ints ----------
void set_error(const std::string& msg);
// per-family launch accounting (engine.cu; usot_profile_read): families 8 = conv_wgrad, 9 = train_other, 7 = other, 2 = maxpool, 1 = stem
void count_op_launch(int family, int n);
enum { OPFAM_STEM = 1, OPFAM_POOL = 2, OPFAM_FUSION = 5, OPFAM_OTHER = 7, OPFAM_WGRAD = 8, OPFAM_TRAIN = 9 };
#define USOT_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::usot::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)
#define USOT_REQUIRE(cond, msg)                                                                 \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            ::usot::set_error(std::string("usot_b200: ") + (msg) + " [" #cond "]");             \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

// Opt-in dynamic shared memory is a per-device function attribute: remember, per device, the largest size already granted
// (engines of several devices may live in one process, e.g. DataParallel-style replicas).
struct SmemAttrCache {
    int granted[64] = {0};
    template <typename Kernel>
    int ensure(Kernel kernel, int bytes) {
        int dev = 0;
        USOT_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || bytes > granted[dev]) {
            USOT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            if (dev >= 0 && dev < 64) granted[dev] = bytes;
        }
        return 0;
    }
};

// SM count of the current device (cached per device; engines of several devices may live in one process).
inline int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Scratch buffers of the stand-alone / training ops come from the stream-ordered allocator (cudaMallocAsync: no host sync).  Keep freed
// blocks cached in the device's default pool instead of returning them to the driver at every synchronisation point.
inline void ensure_async_pool() {
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

// ---- activation storage ------------------------------------------------------------------
// All internal activations are NHWC.  Two storage formats:
//   F32   : one fp32 plane (SIMT path, bandwidth kernels)
//   SPLIT : two fp16 planes, value = hi + lo (tcgen05 path; hi = rn_fp16(v), lo = rn_fp16(v - hi))
struct Act {
    float* f32 = nullptr;
    __half* hi = nullptr;
    __half* lo = nullptr;
    int n = 0, h = 0, w = 0, c = 0;
    size_t numel() const { return (size_t)n * h * w * c; }
};

struct ConvGeom {
    int n, h, w, cin;        // input NHWC
    int cout, kh, kw;        // filter
    int stride, ph, pw, dh, dw;
    int ho, wo;              // output spatial size
};

inline int conv_out(int in, int k, int stride, int pad, int dil) { return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1; }

// Epilogue of every dense conv: y = acc*scale[c] + shift[c] (+ residual) ; optional ReLU.
struct Epilogue {
    const float* scale;     // [cout] folded BN scale (1 if none)
    const float* shift;     // [cout] folded BN shift + conv bias
    const float* residual;  // NHWC fp32 [m][cout] or nullptr
    int relu;
};

// ---- launchers (kernels_simt.cu) -----------------------------------------------------------
int launch_stem(const float* x_nchw, int n, int s, const float* w_packed /*[147][64]*/, const float* scale,
                const float* shift, float* out_nhwc /*[n][ho][ho][64]*/, cudaStream_t st, int relu = 1);  // scale / shift may be null (1 / 0)
// tensor-core stem (stem_tc.cu): w_img = packed shared-memory image of the weight tile, see pack_stem_tc_host
int launch_stem_tc(const float* x_nchw, int n, int s, const void* w_img, const float* scale_tc, const float* shift, float* out_nhwc,
                   bool split, cudaStream_t st);
int launch_maxpool3x3s2p1(const float* in, int n, int h, int w, int c, float* out /*fp32 or null*/, __half* out_hi /*split planes or null*/,
                          __half* out_lo, cudaStream_t st);
int launch_conv_simt(const float* in, const ConvGeom& g, const float* w_kn /*[kh*kw*cin][cout]*/, const Epilogue& ep,
                     float* out, cudaStream_t st);
// Fused GroupDW: out[n] = sum_i sw[i] * xcorr(x_i[n / (n_out/nx)], z_i[n / (n_out/nz)])
struct GroupDWArgs {
    const float* x11; const float* x12; const float* x21;  // [nx][F-2][F-2][C], [nx][F-4][F-2][C], [nx][F-2][F-4][C]
    const float* z11; const float* z12; const float* z21;  // [nz][5][5][C], [nz][3][5][C], [nz][5][3][C]
    const float* dw_weight;                                // [3] raw (softmax applied inside)
    float* out;                                            // [n_out][R][R][C]
    int nx, nz, n_out, C, F;                               // R = F - 6
    __half* out_hi = nullptr;                              // optional: write the split-fp16 planes the tower convs read instead of
    __half* out_lo = nullptr;                              // fp32 `out` (FFMA2 kernel only; out_lo may be null in single-fp16 mode)
};
extern Tunable g_groupdw_strips, g_groupdw_tma, g_groupdw_row_split, g_groupdw_warps4;
bool groupdw_split_output_supported(int F);  // true when launch_groupdw_w will run the FFMA2 kernel, which can write split-fp16 planes
int launch_groupdw_tma(const GroupDWArgs& a, float w0, float w1, float w2, cudaStream_t st);  // xcorr_tma.cu
int launch_groupdw(const GroupDWArgs& a, cudaStream_t st);  // reads dw_weight back (one stream sync)
int launch_groupdw_w(const GroupDWArgs& a, float w0, float w1, float w2, cudaStream_t st);  // softmaxed weights given
// Single depth-wise xcorr, NCHW (the reference op, connect.py:147-157)
int launch_xcorr_nchw(const float* x, const float* k, float* out, int nx, int nk, int C, int hx, int wx, int hk, int wk,
                      cudaStream_t st);
// Skinny prediction conv 3x3 p1, Cin = C, Cout in {1,4}, NHWC in -> NCHW out (pred_tma.cu).  mode 0: out = mul*(y+b) ; mode 1:
// exp(adjust*(y+b)+bias4[co]).  w4 = launch_pred_repack(w) or nullptr (repacked on the fly).
int launch_pred_conv(const float* in, int n, int r, int C, const float* w /*[9][cout][C]*/, const float* w4 /*[C/4][9*cout][4]*/,
                     const float* b, int cout, int mode, float mul, const float* adjust, const float* bias4, float* out_nchw, cudaStream_t st);
int launch_pred_repack(const float* w, int cout, int C, float* w4, cudaStream_t st);
// TMA-streamed per-image variant for large batches (pred_tma.cu); launch_pred_conv dispatches to it when supported
extern Tunable g_pred_tma_min_batch;
bool pred_tma_supported(int n, int r, int C, int cout);
int launch_pred_tma(const float* in, int n, int r, int C, const float* w, const float* b, int cout, int mode, float mul,
                    const float* adjust, const float* bias4, float* out_nchw, cudaStream_t st);
int launch_conf_fusion(const float* conf, const float* value, int b, int nq, size_t per_map /*R*R*C*/, float* out,
                       cudaStream_t st);
int launch_prroi_nhwc(const float* feat, int n_feat, int h, int w, int c, const float* boxes4, int n_rois, float* out_nhwc,
                      cudaStream_t st);
int launch_prroi_nchw(const float* feat, int c, int h, int w, const float* rois5, int n_rois, int ph, int pw, float scale,
                      float* out, cudaStream_t st);
int launch_nchw_to_nhwc(const float* in, int n, int c, int h, int w, float* out, cudaStream_t st);
int launch_nhwc_to_nchw(const float* in, int n, int h, int w, int c, float* out, cudaStream_t st);
int launch_tracker_post(const float* cls, const float* cls_mem, const float* bbox, const double* window, int R, int instance_size,
                        double tw, double th, float ratio, double penalty_k, double window_influence, double* result, cudaStream_t st);
int launch_cycle_glue(const float* off_cls, const float* mem_cls, const float* off_bbox, int n, int R, int search_size, int sf_size,
                      float ratio, float* pool_box, float* best_score, int* best_idx, cudaStream_t st);
int launch_prroi_backward(const float* rois, const float* top_diff, float* bottom_diff, int n_features, int n_rois, int C, int H, int W,
                          int PH, int PW, float scale, cudaStream_t st);
int launch_prroi_coor_backward(const float* feat, const float* rois, const float* top, const float* top_diff, float* rois_diff, int n_rois,
                               int C, int H, int W, int PH, int PW, float scale, cudaStream_t st);
int launch_bce(const float* pred, const float* label, int count, float* out, cudaStream_t st);
int launch_iou(const float* bbox, const float* target, const float* weight, int n, int cells, float* out, cudaStream_t st);
// crop.cu: batched context-window crop + average-colour padding + fixed-point bilinear resize (bit-exact with cv2.resize on uint8)
int launch_crop_resize(const uint8_t* frames, int n_frames, int H, int W, const int* crops /*[n][4] = frame, xmin, ymin, original_sz*/,
                       const uint8_t* fills /*[n][3]*/, int n, int model_sz, float* out_nchw, cudaStream_t st);
int launch_crop_resize_one(const uint8_t* frame, int H, int W, int xmin, int ymin, int osz, const uint8_t* fill3, int model_sz, float* out_nchw,
                           cudaStream_t st);
// kernels_glue.cu: pieces of the one-call tracker frame (usot_engine_track_frame)
int launch_xcorr_backward(const float* x, const float* k, const float* gout, float* gx /*or null*/, float* gk /*or null*/, int nx, int nk, int C,
                          int hx, int wx, int hk, int wk, cudaStream_t st);
int launch_gather_rows(const float* buf, const int* rows_host, int n_rows, size_t row_floats, float* out, cudaStream_t st);
int launch_pool_box_from_result(const double* result, int score_size, int instance_size, int total_stride, float* box4, cudaStream_t st);
int launch_center_crop_nhwc(const float* in, int n, int h, int w, int c, int l, float* out, cudaStream_t st);

}  // namespace usot
