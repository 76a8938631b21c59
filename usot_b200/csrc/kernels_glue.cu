// Small latency-bound kernels around the heads:
//   tracker_post_kernel   tensor path of USOTTracker.update          lib/tracker/usot_tracker.py:137-163
//   cycle_glue_kernel     forward-track argmax + box gather + PrPool box map   lib/models/models.py:131-162,265-274
//   bce_kernel            _weighted_BCE / _cls_loss                   lib/models/models.py:42-58
//   iou_kernel            add_iouloss / _IOULoss                      lib/models/models.py:60-100
#include "common.cuh"
#include <cfloat>
#include <cmath>

namespace usot {

// ---- block-wide argmax (first index wins ties, like numpy.argmax / torch.max on a flat map) -------------------
template <typename TV>
static __device__ void block_argmax(TV& v, int& idx) {
    __shared__ double sv[32];
    __shared__ int si[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double dv = (double)v;
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, dv, off);
        int oi = __shfl_down_sync(0xffffffffu, idx, off);
        if (ov > dv || (ov == dv && oi < idx)) { dv = ov; idx = oi; }
    }
    if (lane == 0) { sv[warp] = dv; si[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        dv = lane < nw ? sv[lane] : -DBL_MAX;
        idx = lane < nw ? si[lane] : 0x7fffffff;
        for (int off = 16; off > 0; off >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, dv, off);
            int oi = __shfl_down_sync(0xffffffffu, idx, off);
            if (ov > dv || (ov == dv && oi < idx)) { dv = ov; idx = oi; }
        }
        if (lane == 0) { sv[0] = dv; si[0] = idx; }
    }
    __syncthreads();
    v = (TV)sv[0];
    idx = si[0];
    __syncthreads();
}

// ---- tracker post-processing: one block, one frame ------------------------------------------------------------
// Arithmetic types follow the reference: sigmoid / ratio mix in float32 (torch + numpy float32), everything that touches the
// float64 grids (box decode, penalties, window) in double.
__global__ void __launch_bounds__(256) tracker_post_kernel(const float* __restrict__ cls, const float* __restrict__ cls_mem,
                                                           const float* __restrict__ bbox, const double* __restrict__ window, int R,
                                                           int instance_size, double tw, double th, float ratio, double penalty_k,
                                                           double window_influence, double* __restrict__ result) {
    const int cells = R * R;
    double best = -DBL_MAX;
    int best_i = 0x7fffffff;
    const double half = (double)(instance_size / 2);
    auto change = [](double r) { return fmax(r, 1.0 / r); };
    auto szf = [](double w, double h) { double pad = (w + h) * 0.5; return sqrt((w + pad) * (h + pad)); };
    const double sz_t = szf(tw, th);
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const int r = i / R, c = i % R;
        const float s0 = 1.f / (1.f + expf(-cls[i])), s1 = 1.f / (1.f + expf(-cls_mem[i]));
        const float mix = ratio * s0 + (1.f - ratio) * s1;
        const double gx = (double)(c - R / 2) * 8.0 + half, gy = (double)(r - R / 2) * 8.0 + half;
        const double x1 = gx - (double)bbox[i], y1 = gy - (double)bbox[cells + i];
        const double x2 = gx + (double)bbox[2 * cells + i], y2 = gy + (double)bbox[3 * cells + i];
        const double s_c = change(szf(x2 - x1, y2 - y1) / sz_t);
        const double r_c = change((tw / th) / ((x2 - x1) / (y2 - y1)));
        const double pen = exp(-(r_c * s_c - 1.0) * penalty_k);
        const double ps = pen * (double)mix * (1.0 - window_influence) + window[i] * window_influence;
        // a NaN score is the maximum for numpy's argmax (first NaN wins): map it to +inf so the smallest NaN index is returned
        const double key = ps != ps ? DBL_MAX : ps;
        if (key > best) { best = key; best_i = i; }
    }
    block_argmax(best, best_i);
    if (threadIdx.x == 0) {
        // No cell compared greater than -DBL_MAX: every penalised score is NaN (non-finite weights / inf boxes).  numpy's argmax
        // (usot_tracker.py:157) returns the first NaN cell, which is cell 0 when all are NaN: never index out of bounds.
        const int i = (best_i >= 0 && best_i < cells) ? best_i : 0, r = i / R, c = i % R;
        const float s0 = 1.f / (1.f + expf(-cls[i])), s1 = 1.f / (1.f + expf(-cls_mem[i]));
        const float mix = ratio * s0 + (1.f - ratio) * s1;
        const double gx = (double)(c - R / 2) * 8.0 + half, gy = (double)(r - R / 2) * 8.0 + half;
        const double x1 = gx - (double)bbox[i], y1 = gy - (double)bbox[cells + i];
        const double x2 = gx + (double)bbox[2 * cells + i], y2 = gy + (double)bbox[3 * cells + i];
        const double s_c = change(szf(x2 - x1, y2 - y1) / sz_t);
        const double r_c = change((tw / th) / ((x2 - x1) / (y2 - y1)));
        result[0] = r; result[1] = c; result[2] = x1; result[3] = y1; result[4] = x2; result[5] = y2;
        result[6] = exp(-(r_c * s_c - 1.0) * penalty_k);
        result[7] = (double)mix;
    }
}

int launch_tracker_post(const float* cls, const float* cls_mem, const float* bbox, const double* window, int R, int instance_size,
                        double tw, double th, float ratio, double penalty_k, double window_influence, double* result, cudaStream_t st) {
    USOT_REQUIRE(R > 0 && R <= 64, "tracker postprocess: bad score size");
    tracker_post_kernel<<<1, 256, 0, st>>>(cls, cls_mem, bbox, window, R, instance_size, tw, th, ratio, penalty_k, window_influence, result);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- backward of the depth-wise cross-correlation (training path) -----------------------------------
// Forward (connect.py:147-157): out[b,c,i,j] = sum_{u,v} x[b,c,i+u,j+v] * k[b',c,u,v], b' = b or 0 (kernel batch 1 broadcast).
//   grad_x[b,c,y,x] = sum_{u,v} gout[b,c,y-u,x-v] * k[b',c,u,v]                 (full correlation, zero outside gout)
//   grad_k[b',c,u,v] = sum_{b -> b'} sum_{i,j} gout[b,c,i,j] * x[b,c,i+u,j+v]    (reduced over positions and, when the kernel is
//                                                                                broadcast, over the samples that share it)
// One block per (sample, channel) plane, like the forward op: x, gout and the taps are staged in shared memory.
__global__ void __launch_bounds__(256) xcorr_backward_kernel(const float* __restrict__ x, const float* __restrict__ k, const float* __restrict__ gout,
                                                             float* __restrict__ gx, float* __restrict__ gk, int nk, int C, int hx, int wx,
                                                             int hk, int wk) {
    extern __shared__ float xb_sm[];
    const int ho = hx - hk + 1, wo = wx - wk + 1;
    float* xs = xb_sm;
    float* gs = xs + hx * wx;
    float* ks = gs + ho * wo;
    float* red = ks + hk * wk;  // [8] warp partials
    const int plane = blockIdx.x, b = plane / C, c = plane % C;
    const int kb = (nk == 1) ? 0 : b;
    const float* xp = x + (size_t)plane * hx * wx;
    const float* gp = gout + (size_t)plane * ho * wo;
    const float* kp = k + ((size_t)kb * C + c) * hk * wk;
    for (int i = threadIdx.x; i < hx * wx; i += 256) xs[i] = __ldg(xp + i);
    for (int i = threadIdx.x; i < ho * wo; i += 256) gs[i] = __ldg(gp + i);
    for (int i = threadIdx.x; i < hk * wk; i += 256) ks[i] = __ldg(kp + i);
    __syncthreads();
    if (gx) {
        for (int o = threadIdx.x; o < hx * wx; o += 256) {
            const int y = o / wx, xx = o % wx;
            float s = 0.f;
            for (int u = 0; u < hk; ++u) {
                const int i = y - u;
                if (i < 0 || i >= ho) continue;
                for (int v = 0; v < wk; ++v) {
                    const int j = xx - v;
                    if (j >= 0 && j < wo) s = fmaf(gs[i * wo + j], ks[u * wk + v], s);
                }
            }
            gx[(size_t)plane * hx * wx + o] = s;
        }
    }
    if (gk) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int t = 0; t < hk * wk; ++t) {
            const int u = t / wk, v = t % wk;
            float s = 0.f;
            for (int o = threadIdx.x; o < ho * wo; o += 256) {
                const int i = o / wo, j = o % wo;
                s = fmaf(gs[o], xs[(i + u) * wx + j + v], s);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            if (threadIdx.x == 0) {
                float tot = 0.f;
                for (int w8 = 0; w8 < 8; ++w8) tot += red[w8];
                float* dst = gk + ((size_t)kb * C + c) * hk * wk + t;
                if (nk == 1) atomicAdd(dst, tot);  // the broadcast kernel collects every sample's contribution
                else *dst = tot;
            }
            __syncthreads();
        }
    }
}

int launch_xcorr_backward(const float* x, const float* k, const float* gout, float* gx, float* gk, int nx, int nk, int C, int hx, int wx,
                          int hk, int wk, cudaStream_t st) {
    USOT_REQUIRE(nk == 1 || nk == nx, "xcorr backward: kernel batch must be 1 or equal to the search batch");
    USOT_REQUIRE(hx >= hk && wx >= wk, "xcorr backward: kernel larger than search map");
    const int ho = hx - hk + 1, wo = wx - wk + 1;
    const size_t smem = ((size_t)hx * wx + (size_t)ho * wo + (size_t)hk * wk + 8) * sizeof(float);
    USOT_REQUIRE(smem <= 48 * 1024, "xcorr backward: plane too large");
    if (gk && nk == 1) USOT_CUDA_OK(cudaMemsetAsync(gk, 0, (size_t)C * hk * wk * sizeof(float), st));
    if (nx * C == 0) return 0;
    xcorr_backward_kernel<<<nx * C, 256, smem, st>>>(x, k, gout, gx, gk, nk, C, hx, wx, hk, wk);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- one-call tracker frame helpers ------------------------------------------------------------------------------
struct RowList { int r[16]; };
__global__ void gather_rows_kernel(const float4* __restrict__ buf, RowList rows, size_t row4, float4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < row4) out[(size_t)blockIdx.y * row4 + i] = __ldg(buf + (size_t)rows.r[blockIdx.y] * row4 + i);
}
// Memory templates of this frame: out[k] = buf[rows[k]] (device-resident queue, usot_tracker.py:222-261 without the host round trip)
int launch_gather_rows(const float* buf, const int* rows_host, int n_rows, size_t row_floats, float* out, cudaStream_t st) {
    USOT_REQUIRE(n_rows > 0 && n_rows <= 16 && row_floats % 4 == 0, "gather: 1..16 rows of a multiple of 4 floats");
    RowList rl;
    for (int k = 0; k < 16; ++k) rl.r[k] = k < n_rows ? rows_host[k] : 0;
    const size_t row4 = row_floats / 4;
    dim3 grid((unsigned)((row4 + 255) / 256), (unsigned)n_rows);
    gather_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(buf), rl, row4, reinterpret_cast<float4*>(out));
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// pool_label_search (usot_tracker.py:339-362) of the winning box, in the reference's float32 arithmetic:
//   bbox = clip(float32([x1,y1,x2,y2]), reg_min - gap, reg_max + gap);  (bbox - reg_min) * slope
__global__ void pool_box_kernel(const double* __restrict__ result, float reg_min, float reg_max, float slope, float gap, float* __restrict__ box4) {
    const int k = threadIdx.x;
    if (k < 4) {
        float v = (float)result[2 + k];
        v = fminf(fmaxf(v, reg_min - gap), reg_max + gap);
        box4[k] = __fmul_rn(__fsub_rn(v, reg_min), slope);
    }
}
int launch_pool_box_from_result(const double* result, int score_size, int instance_size, int total_stride, float* box4, cudaStream_t st) {
    const int sf = score_size;
    const double reg_min = (double)(0 - sf / 2) * total_stride + instance_size / 2;
    const double reg_max = (double)(sf - 1 - sf / 2) * total_stride + instance_size / 2;
    const double slope = (2 * (sf / 2)) / (reg_max - reg_min);
    pool_box_kernel<<<1, 32, 0, st>>>(result, (float)reg_min, (float)reg_max, (float)slope, (float)(1.0 / slope), box4);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- cycle-memory glue: one block per memory sample --------------------------------------------------------------
// res = r*off_cls + (1-r)*mem_cls ; idx = argmax ; box = grid(idx) -/+ off_bbox[:, idx] ; pool_box = map to PrPool coordinates
__global__ void __launch_bounds__(256) cycle_glue_kernel(const float* __restrict__ off_cls, const float* __restrict__ mem_cls,
                                                         const float* __restrict__ off_bbox, int R, int search_size, int sf_size,
                                                         float ratio, float* __restrict__ pool_box, float* __restrict__ best_score,
                                                         int* __restrict__ best_idx) {
    const int s = blockIdx.x, cells = R * R;
    const float* oc = off_cls + (size_t)s * cells;
    const float* mc = mem_cls + (size_t)s * cells;
    float best = -FLT_MAX;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const float v = ratio * oc[i] + (1.f - ratio) * mc[i];
        if (v > best) { best = v; bi = i; }
    }
    block_argmax(best, bi);
    if (threadIdx.x == 0) {
        const int r = bi / R, c = bi % R;
        const float half = (float)(search_size / 2);
        const float gx = (float)(c - R / 2) * 8.f + half, gy = (float)(r - R / 2) * 8.f + half;  // models.py:107-117
        const float* bb = off_bbox + (size_t)s * 4 * cells;
        float box[4] = {gx - bb[bi], gy - bb[cells + bi], gx + bb[2 * cells + bi], gy + bb[3 * cells + bi]};
        // image_bbox_to_prpool_bbox, models.py:150-162
        const double reg_min = (double)(0 - sf_size / 2) * 8.0 + (double)(search_size / 2);
        const double reg_max = (double)(sf_size - 1 - sf_size / 2) * 8.0 + (double)(search_size / 2);
        const double gap = (reg_max - reg_min) / (double)(2 * (sf_size / 2));
        const float lo = (float)(reg_min - 2 * gap), hi = (float)(reg_max + 2 * gap), slope = (float)(1.0 / gap), rmin = (float)reg_min;
        for (int k = 0; k < 4; ++k) pool_box[(size_t)s * 4 + k] = (fminf(fmaxf(box[k], lo), hi) - rmin) * slope;
        if (best_score) best_score[s] = best;
        if (best_idx) best_idx[s] = bi;
    }
}

int launch_cycle_glue(const float* off_cls, const float* mem_cls, const float* off_bbox, int n, int R, int search_size, int sf_size,
                      float ratio, float* pool_box, float* best_score, int* best_idx, cudaStream_t st) {
    if (n == 0) return 0;
    cycle_glue_kernel<<<n, 256, 0, st>>>(off_cls, mem_cls, off_bbox, R, search_size, sf_size, ratio, pool_box, best_score, best_idx);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- losses (single block, double accumulation) ---------------------------------------------------------------------
static __device__ double block_sum(double v) {
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (warp == 0)
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) sm[0] = v;
    __syncthreads();
    v = sm[0];
    __syncthreads();
    return v;
}

// 0.5 * mean_{label==1} BCEWithLogits(pred,1) + 0.5 * mean_{label==0} BCEWithLogits(pred,0); a class with exactly one member
// contributes 0 (models.py:43-44 quirk), an empty class contributes NaN (mean of an empty tensor).
__global__ void __launch_bounds__(1024) bce_kernel(const float* __restrict__ pred, const float* __restrict__ label, int count,
                                                   float* __restrict__ out) {
    double sp = 0, sn = 0, cp = 0, cn = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const float x = pred[i], y = label[i];
        const float sl = log1pf(expf(-fabsf(x)));
        if (y == 1.f) { sp += (double)(fmaxf(-x, 0.f) + sl); cp += 1; }
        else if (y == 0.f) { sn += (double)(fmaxf(x, 0.f) + sl); cn += 1; }
    }
    sp = block_sum(sp); sn = block_sum(sn); cp = block_sum(cp); cn = block_sum(cn);
    if (threadIdx.x == 0) {
        const double lp = cp == 1 ? 0.0 : sp / cp, ln = cn == 1 ? 0.0 : sn / cn;
        *out = (float)(lp * 0.5 + ln * 0.5);
    }
}

// bbox_pred (n,4,R,R) nchw ; reg_target (n,R,R,4) ; reg_weight (n,R,R)
__global__ void __launch_bounds__(1024) iou_kernel(const float* __restrict__ bbox, const float* __restrict__ target,
                                                   const float* __restrict__ weight, int n, int cells, float* __restrict__ out) {
    double s = 0, cnt = 0;
    for (int i = threadIdx.x; i < n * cells; i += blockDim.x) {
        if (!(weight[i] > 0.f)) continue;
        const int b = i / cells, c = i % cells;
        const float* p = bbox + (size_t)b * 4 * cells + c;
        const float pl = p[0], pt = p[cells], pr = p[2 * cells], pb = p[3 * cells];
        const float* t = target + (size_t)i * 4;
        const float tl = t[0], tt = t[1], tr = t[2], tb = t[3];
        const float ta = (tl + tr) * (tt + tb), pa = (pl + pr) * (pt + pb);
        const float wi = fminf(pl, tl) + fminf(pr, tr), hi = fminf(pb, tb) + fminf(pt, tt);
        const float ai = wi * hi, au = ta + pa - ai;
        s += (double)(-logf((ai + 1.0f) / (au + 1.0f)));
        cnt += 1;
    }
    s = block_sum(s); cnt = block_sum(cnt);
    if (threadIdx.x == 0) *out = (float)(s / cnt);
}

// =============================================================================================
// PrRoIPool backward (reference layouts: NCHW features, rois (n,5))      prroi_pooling_gpu_impl.cu:214-379
// =============================================================================================
static __device__ __forceinline__ float pr_g(float lim, float a) { return lim - 0.5f * lim * lim - a + 0.5f * a * a; }
static __device__ __forceinline__ void pr_scatter(float* g, int h, int w, int H, int W, float v) {
    if (h >= 0 && w >= 0 && h < H && w < W) atomicAdd(g + h * W + w, v);
}
struct PrBin { float ws_w, ws_h, we_w, we_h, win; };
static __device__ __forceinline__ PrBin pr_bin(const float* r, int ph, int pw, int PH, int PW, float scale) {
    const float sw = r[1] * scale, sh = r[2] * scale, ew = r[3] * scale, eh = r[4] * scale;
    const float bw = fmaxf(ew - sw, 0.f) / (float)PW, bh = fmaxf(eh - sh, 0.f) / (float)PH;
    PrBin b;
    b.ws_w = sw + bw * pw; b.ws_h = sh + bh * ph; b.we_w = b.ws_w + bw; b.we_h = b.ws_h + bh;
    b.win = fmaxf(0.f, bw * bh);
    return b;
}

// d(out)/d(features): the forward's corner coefficients, scattered with atomics (several bins / rois touch the same cell)
__global__ void prroi_backward_kernel(const float* __restrict__ rois, const float* __restrict__ top_diff, float* __restrict__ bottom_diff,
                                      size_t total, int C, int H, int W, int PH, int PW, float scale) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int pw = idx % PW, ph = (idx / PW) % PH, c = (idx / ((size_t)PW * PH)) % C;
    const int n = idx / ((size_t)PW * PH * C);
    const float* r = rois + (size_t)n * 5;
    const PrBin b = pr_bin(r, ph, pw, PH, PW, scale);
    if (b.win == 0.f) return;
    const float g = top_diff[idx] / b.win;
    float* dst = bottom_diff + ((size_t)(int)r[0] * C + c) * H * W;
    const int s_w = (int)floorf(b.ws_w), e_w = (int)ceilf(b.we_w), s_h = (int)floorf(b.ws_h), e_h = (int)ceilf(b.we_h);
    for (int wi = s_w; wi < e_w; ++wi)
        for (int hi = s_h; hi < e_h; ++hi) {
            const float y0 = fmaxf(b.ws_h, (float)hi), x0 = fmaxf(b.ws_w, (float)wi);
            const float y1 = fminf(b.we_h, (float)hi + 1.f), x1 = fminf(b.we_w, (float)wi + 1.f);
            const float ax = pr_g(x1 - wi, x0 - wi), ay = pr_g(y1 - hi, y0 - hi);
            const float bx = pr_g((float)(wi + 1) - x0, (float)(wi + 1) - x1), by = pr_g((float)(hi + 1) - y0, (float)(hi + 1) - y1);
            pr_scatter(dst, hi, wi, H, W, g * (ax * ay));
            pr_scatter(dst, hi, wi + 1, H, W, g * (bx * ay));
            pr_scatter(dst, hi + 1, wi, H, W, g * (ax * by));
            pr_scatter(dst, hi + 1, wi + 1, H, W, g * (bx * by));
        }
}

static __device__ __forceinline__ float pr_at(const float* d, int h, int w, int H, int W) {
    return (h < 0 || w < 0 || h >= H || w >= W) ? 0.f : d[h * W + w];
}
static __device__ float pr_bilinear(const float* d, float h, float w, int H, int W) {  // zero outside the map
    const int h0 = (int)floorf(h), w0 = (int)floorf(w);
    const float dh = h - (float)h0, dw = w - (float)w0;
    return pr_at(d, h0, w0, H, W) * ((1.f - dh) * (1.f - dw)) + pr_at(d, h0 + 1, w0, H, W) * (dh * (1.f - dw)) +
           pr_at(d, h0, w0 + 1, H, W) * ((1.f - dh) * dw) + pr_at(d, h0 + 1, w0 + 1, H, W) * (dh * dw);
}
// integral over [s,t] of the linear interpolant between c1 (at 0) and c2 (at 1); double sub-expressions as in the reference
static __device__ __forceinline__ float pr_edge(float s, float t, float c1, float c2) {
    return (float)(0.5 * (double)(t * t - s * s) * (double)c2 + ((double)t - 0.5 * (double)(t * t) - (double)s + 0.5 * (double)(s * s)) * (double)c1);
}

// d(out)/d(x1,y1,x2,y2): line integrals of the interpolant along the four bin edges (Leibniz rule)
__global__ void prroi_coor_backward_kernel(const float* __restrict__ feat, const float* __restrict__ rois, const float* __restrict__ top,
                                           const float* __restrict__ top_diff, float* __restrict__ rois_diff, size_t total, int C, int H,
                                           int W, int PH, int PW, float scale) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int pw = idx % PW, ph = (idx / PW) % PH, c = (idx / ((size_t)PW * PH)) % C;
    const int n = idx / ((size_t)PW * PH * C);
    const float* r = rois + (size_t)n * 5;
    const PrBin b = pr_bin(r, ph, pw, PH, PW, scale);
    const float go = top_diff[idx];
    if (b.win == 0.f || go / b.win == 0.f) return;  // (the reference skips zero upstream gradients, :323-325)
    const float* d = feat + ((size_t)(int)r[0] * C + c) * H * W;
    const int s_w = (int)floorf(b.ws_w), e_w = (int)ceilf(b.we_w), s_h = (int)floorf(b.ws_h), e_h = (int)ceilf(b.we_h);
    float gx1 = 0.f, gx2 = 0.f, gy1 = 0.f, gy2 = 0.f;
    for (int hi = s_h; hi < e_h; ++hi) {
        const float s = fmaxf(b.ws_h, (float)hi) - hi, t = fminf(b.we_h, (float)(hi + 1)) - hi;
        gx1 += pr_edge(s, t, pr_bilinear(d, (float)hi, b.ws_w, H, W), pr_bilinear(d, (float)(hi + 1), b.ws_w, H, W));
        gx2 += pr_edge(s, t, pr_bilinear(d, (float)hi, b.we_w, H, W), pr_bilinear(d, (float)(hi + 1), b.we_w, H, W));
    }
    for (int wi = s_w; wi < e_w; ++wi) {
        const float s = fmaxf(b.ws_w, (float)wi) - wi, t = fminf(b.we_w, (float)(wi + 1)) - wi;
        gy1 += pr_edge(s, t, pr_bilinear(d, b.ws_h, (float)wi, H, W), pr_bilinear(d, b.ws_h, (float)(wi + 1), H, W));
        gy2 += pr_edge(s, t, pr_bilinear(d, b.we_h, (float)wi, H, W), pr_bilinear(d, b.we_h, (float)(wi + 1), H, W));
    }
    const float o = top[idx];
    float px1 = -gx1 + (b.we_h - b.ws_h) * o, py1 = -gy1 + (b.we_w - b.ws_w) * o;
    float px2 = gx2 - (b.we_h - b.ws_h) * o, py2 = gy2 - (b.we_w - b.ws_w) * o;
    px1 = px1 / b.win * scale; px2 = px2 / b.win * scale; py1 = py1 / b.win * scale; py2 = py2 / b.win * scale;
    float* g = rois_diff + (size_t)n * 5;
    const double fw0 = (double)((float)pw / PW), fw1 = (double)((float)(pw + 1) / PW);
    const double fh0 = (double)((float)ph / PH), fh1 = (double)((float)(ph + 1) / PH);
    atomicAdd(g + 1, (float)(((double)px1 * (1.0 - fw0) + (double)px2 * (1.0 - fw1)) * (double)go));
    atomicAdd(g + 2, (float)(((double)py1 * (1.0 - fh0) + (double)py2 * (1.0 - fh1)) * (double)go));
    atomicAdd(g + 3, (float)(((double)(px2 * (float)(pw + 1) / PW) + (double)(px1 * (float)pw / PW)) * (double)go));
    atomicAdd(g + 4, (float)(((double)(py2 * (float)(ph + 1) / PH) + (double)(py1 * (float)ph / PH)) * (double)go));
}

int launch_prroi_backward(const float* rois, const float* top_diff, float* bottom_diff, int n_features, int n_rois, int C, int H, int W,
                          int PH, int PW, float scale, cudaStream_t st) {
    USOT_CUDA_OK(cudaMemsetAsync(bottom_diff, 0, (size_t)n_features * C * H * W * sizeof(float), st));
    const size_t total = (size_t)n_rois * C * PH * PW;
    if (total == 0) return 0;
    prroi_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rois, top_diff, bottom_diff, total, C, H, W, PH, PW, scale);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_prroi_coor_backward(const float* feat, const float* rois, const float* top, const float* top_diff, float* rois_diff, int n_rois,
                               int C, int H, int W, int PH, int PW, float scale, cudaStream_t st) {
    USOT_CUDA_OK(cudaMemsetAsync(rois_diff, 0, (size_t)n_rois * 5 * sizeof(float), st));
    const size_t total = (size_t)n_rois * C * PH * PW;
    if (total == 0) return 0;
    prroi_coor_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, rois, top, top_diff, rois_diff, total, C, H, W, PH, PW, scale);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_bce(const float* pred, const float* label, int count, float* out, cudaStream_t st) {
    bce_kernel<<<1, 1024, 0, st>>>(pred, label, count, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_iou(const float* bbox, const float* target, const float* weight, int n, int cells, float* out, cudaStream_t st) {
    iou_kernel<<<1, 1024, 0, st>>>(bbox, target, weight, n, cells, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace usot
