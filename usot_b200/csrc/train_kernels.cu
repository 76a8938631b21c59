// Training path: the kernels the BACKWARD of USOT_.forward (lib/models/models.py:208-295, driven by
// scripts/train_usot.py:229-236) needs beyond the forward kernels, plus train-mode BatchNorm.  All tensors NHWC fp32.
//
//   conv wgrad            autograd of nn.Conv2d w.r.t. the weight       (any stride / dilation / kernel size / channel count)
//   conv dgrad (generic)  autograd of nn.Conv2d w.r.t. the input        (thin convs: the 256->1/4 prediction heads; every wide
//                                                                        layer runs its dgrad on the forward conv kernels, ops.py)
//   BatchNorm2d           batch statistics + normalise (+bias)(+residual)(+ReLU), and its backward (train and eval flavour)
//   MaxPool 3x3/2 p1      backward (first-maximum rule of torch's max_pool2d)
//   Conf_Fusion           backward of the exp / normalise / weighted-sum reduction   lib/models/connect.py:130-144
//   weighted sum of 3     GroupDW's softmax-weighted sum of the three correlations   lib/models/connect.py:96-102
//   _weighted_BCE / IoU   backward of the two losses                                 lib/models/models.py:42-100
//   column sums           bias gradients
#include "common.cuh"
#include "conv_tc.cuh"
#include "../../include/usot_b200.h"

#include <cfloat>

namespace usot {

// =============================================================================================================================
// conv wgrad: dW[(t*cin+ci)][co] += sum_pix X[n, oy*s-ph+kh*dh, ox*s-pw+kw*dw, ci] * dY[n,oy,ox,co]
// One GEMM per filter tap: M = ci, N = co, K = output pixels; both operands are pixel-major in memory, i.e. already in the
// [k][m] / [k][n] shared-memory layout the register-tiled inner product wants -- no transposes.  64x64 tiles, 256 threads, 4x4 per
// thread, pixel chunks of 16, split-K over pixel ranges with fp32 atomics into the (pre-zeroed) result.
// =============================================================================================================================
constexpr int WG_T = 64, WG_K = 16;

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, ConvGeom g, int ci_tiles,
                                                         int co_tiles, int pix_per_split, float* __restrict__ dw) {
    __shared__ __align__(16) float As[WG_K][WG_T + 4];
    __shared__ __align__(16) float Bs[WG_K][WG_T + 4];
    const int tid = threadIdx.x;
    int bt = blockIdx.x;
    const int cot = bt % co_tiles; bt /= co_tiles;
    const int cit = bt % ci_tiles;
    const int tap = bt / ci_tiles;
    const int kh = tap / g.kw, kw = tap % g.kw;
    const int ci0 = cit * WG_T, co0 = cot * WG_T;
    const int M = g.n * g.ho * g.wo;
    const int p_begin = blockIdx.y * pix_per_split, p_end = min(M, p_begin + pix_per_split);
    const int lrow = tid >> 4, lcol = (tid & 15) * 4;   // loader: pixel row of the chunk, 4 consecutive channels
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool vec_a = (g.cin % 4 == 0), vec_b = (g.cout % 4 == 0);
    for (int p0 = p_begin; p0 < p_end; p0 += WG_K) {
        const int p = p0 + lrow;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (p < p_end) {
            const int n = p / (g.ho * g.wo), rem = p % (g.ho * g.wo);
            const int oy = rem / g.wo, ox = rem % g.wo;
            const int iy = oy * g.stride - g.ph + kh * g.dh, ix = ox * g.stride - g.pw + kw * g.dw;
            if (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w) {
                const float* xp = x + (((size_t)n * g.h + iy) * g.w + ix) * g.cin + ci0 + lcol;
                if (vec_a && ci0 + lcol + 3 < g.cin) a = __ldg(reinterpret_cast<const float4*>(xp));
                else {
                    if (ci0 + lcol + 0 < g.cin) a.x = __ldg(xp + 0);
                    if (ci0 + lcol + 1 < g.cin) a.y = __ldg(xp + 1);
                    if (ci0 + lcol + 2 < g.cin) a.z = __ldg(xp + 2);
                    if (ci0 + lcol + 3 < g.cin) a.w = __ldg(xp + 3);
                }
            }
            const float* yp = dy + (size_t)p * g.cout + co0 + lcol;
            if (vec_b && co0 + lcol + 3 < g.cout) b = __ldg(reinterpret_cast<const float4*>(yp));
            else {
                if (co0 + lcol + 0 < g.cout) b.x = __ldg(yp + 0);
                if (co0 + lcol + 1 < g.cout) b.y = __ldg(yp + 1);
                if (co0 + lcol + 2 < g.cout) b.z = __ldg(yp + 2);
                if (co0 + lcol + 3 < g.cout) b.w = __ldg(yp + 3);
            }
        }
        __syncthreads();  // previous chunk fully consumed
        *reinterpret_cast<float4*>(&As[lrow][lcol]) = a;
        *reinterpret_cast<float4*>(&Bs[lrow][lcol]) = b;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WG_K; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ty * 4 + i;
        if (ci >= g.cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co < g.cout) atomicAdd(dw + ((size_t)tap * g.cin + ci) * g.cout + co, acc[i][j]);
        }
    }
}

int launch_conv_wgrad(const float* x, const float* dy, const ConvGeom& g, float* dw_kn, cudaStream_t st) {
    const size_t numel = (size_t)g.kh * g.kw * g.cin * g.cout;
    USOT_CUDA_OK(cudaMemsetAsync(dw_kn, 0, numel * sizeof(float), st));
    const int M = g.n * g.ho * g.wo;
    if (M == 0) return 0;
    const int ci_tiles = (g.cin + WG_T - 1) / WG_T, co_tiles = (g.cout + WG_T - 1) / WG_T;
    const int tiles = g.kh * g.kw * ci_tiles * co_tiles;
    // enough blocks to fill the device a few times over, at least 256 pixels per block
    int splits = (device_sm_count() * 8 + tiles - 1) / tiles;
    splits = std::max(1, std::min(splits, (M + 255) / 256));
    int pps = (M + splits - 1) / splits;
    pps = (pps + WG_K - 1) / WG_K * WG_K;
    splits = (M + pps - 1) / pps;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    conv_wgrad_kernel<<<grid, 256, 0, st>>>(x, dy, g, ci_tiles, co_tiles, pps, dw_kn);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================================================
// stem wgrad: conv1 is 7x7 / stride 2 / pad 0 with Cin = 3 -- as a per-tap GEMM its M dimension would be 3 rows of a 64-row tile.  Here
// (channel, tap) = 147 values form the M dimension instead (im2col on the fly from the NCHW image): dW[k][co] = sum_p patch[p][k] * dY[p][co],
// k = (c*7 + kh)*7 + kw.  Block = a slab of output pixels; 256 threads = 16 (k groups of 10) x 16 (co groups of 4); chunks of 16 pixels
// staged in shared memory; fp32 atomics into the (pre-zeroed) [147][64] result.
// =============================================================================================================================
constexpr int SW_K = 147, SW_KP = 160, SW_CHUNK = 16;

__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x /*nchw (n,3,S,S)*/, const float* __restrict__ dy /*nhwc (n,HO,HO,64)*/,
                                                         int S, int HO, int total_pix, int pix_per_block, float* __restrict__ dw /*[147][64]*/) {
    __shared__ float As[SW_CHUNK][SW_KP];
    __shared__ __align__(16) float Bs[SW_CHUNK][64];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int p_begin = blockIdx.x * pix_per_block, p_end = min(total_pix, p_begin + pix_per_block);
    float acc[10][4];
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int p0 = p_begin; p0 < p_end; p0 += SW_CHUNK) {
        __syncthreads();
        for (int i = tid; i < SW_CHUNK * SW_KP; i += 256) {
            const int r = i / SW_KP, k = i - r * SW_KP, p = p0 + r;
            float v = 0.f;
            if (p < p_end && k < SW_K) {
                const int n = p / (HO * HO), rem = p - n * HO * HO, oy = rem / HO, ox = rem - oy * HO;
                const int c = k / 49, t = k - c * 49, kh = t / 7, kw = t - kh * 7;
                v = __ldg(x + (((size_t)n * 3 + c) * S + oy * 2 + kh) * S + ox * 2 + kw);
            }
            As[r][k] = v;
        }
        for (int i = tid; i < SW_CHUNK * 16; i += 256) {
            const int r = i >> 4, q = i & 15, p = p0 + r;
            reinterpret_cast<float4*>(&Bs[r][0])[q] = p < p_end ? __ldg(reinterpret_cast<const float4*>(dy + (size_t)p * 64) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < SW_CHUNK; ++r) {
            const float4 b = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const float a = As[r][ty * 10 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, bb[j], acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = ty * 10 + i;
        if (k >= SW_K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dw + (size_t)k * 64 + tx * 4 + j, acc[i][j]);
    }
}

int launch_stem_wgrad(const float* x_nchw, const float* dy, int n, int S, float* dw_k64, cudaStream_t st) {
    USOT_CUDA_OK(cudaMemsetAsync(dw_k64, 0, SW_K * 64 * sizeof(float), st));
    const int HO = (S - 7) / 2 + 1, total = n * HO * HO;
    if (total == 0) return 0;
    int blocks = std::min(device_sm_count() * 4, (total + 255) / 256);
    int ppb = (total + blocks - 1) / blocks;
    ppb = (ppb + SW_CHUNK - 1) / SW_CHUNK * SW_CHUNK;
    blocks = (total + ppb - 1) / ppb;
    stem_wgrad_kernel<<<blocks, 256, 0, st>>>(x_nchw, dy, S, HO, total, ppb, dw_k64);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================================================
// generic conv dgrad (gather form): dX[n,y,x,ci] = sum_{taps hitting (y,x)} sum_co dY[n,oy,ox,co] * W[(t*cin+ci)][co]
// Thread per (input pixel, ci).  Used for the thin prediction convs (cout 1 / 4); wide layers use the forward GEMM kernels.
// =============================================================================================================================
__global__ void conv_dgrad_gather_kernel(const float* __restrict__ dy, const float* __restrict__ wkn, ConvGeom g, float* __restrict__ dx) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)g.n * g.h * g.w * g.cin;
    if (idx >= total) return;
    const int ci = idx % g.cin;
    size_t t = idx / g.cin;
    const int xx = t % g.w; t /= g.w;
    const int yy = t % g.h;
    const int n = t / g.h;
    float acc = 0.f;
    for (int kh = 0; kh < g.kh; ++kh) {
        const int ny = yy + g.ph - kh * g.dh;
        if (ny < 0 || ny % g.stride) continue;
        const int oy = ny / g.stride;
        if (oy >= g.ho) continue;
        for (int kw = 0; kw < g.kw; ++kw) {
            const int nx = xx + g.pw - kw * g.dw;
            if (nx < 0 || nx % g.stride) continue;
            const int ox = nx / g.stride;
            if (ox >= g.wo) continue;
            const float* gp = dy + (((size_t)n * g.ho + oy) * g.wo + ox) * g.cout;
            const float* wp = wkn + ((size_t)(kh * g.kw + kw) * g.cin + ci) * g.cout;
            for (int co = 0; co < g.cout; ++co) acc = fmaf(__ldg(gp + co), __ldg(wp + co), acc);
        }
    }
    dx[idx] = acc;
}

int launch_conv_dgrad_gather(const float* dy, const float* wkn, const ConvGeom& g, float* dx, cudaStream_t st) {
    const size_t total = (size_t)g.n * g.h * g.w * g.cin;
    if (total == 0) return 0;
    conv_dgrad_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dy, wkn, g, dx);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================================================
// Per-channel reductions over the pixels of an [M][C] map (double accumulation; partial sums per block, then a tiny finalize pass).
//   mode 0  bn statistics   : s0 = sum (x + bias), s1 = sum (x + bias)^2
//   mode 1  bn backward     : s0 = sum dz, s1 = sum dz * xhat        with dz = dy * (y > 0 if relu) and xhat = (x + bias - mean) * invstd
//   mode 2  column sum      : s0 = sum x
// Block = 32 channels x 8 pixel lanes (256 threads); grid.x = channel groups, grid.y = pixel slabs.
// =============================================================================================================================
struct RedArgs {
    const float* x; const float* bias; const float* dy; const float* y; const float* mean; const float* invstd;
    int M, C, mode, relu, slab;
};

__global__ void __launch_bounds__(256) chan_reduce_kernel(RedArgs a, double* __restrict__ part /*[slabs][2][C]*/) {
    __shared__ double s0s[8][32], s1s[8][32];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int m0 = blockIdx.y * a.slab, m1 = min(a.M, m0 + a.slab);
    double s0 = 0.0, s1 = 0.0;
    if (c < a.C) {
        const float b = a.bias ? __ldg(a.bias + c) : 0.f;
        const float mu = a.mode == 1 ? __ldg(a.mean + c) : 0.f, is = a.mode == 1 ? __ldg(a.invstd + c) : 0.f;
        for (int m = m0 + pl; m < m1; m += 8) {
            const size_t o = (size_t)m * a.C + c;
            if (a.mode == 0) {
                const float v = __ldg(a.x + o) + b;
                s0 += (double)v; s1 += (double)v * (double)v;
            } else if (a.mode == 1) {
                float dz = __ldg(a.dy + o);
                if (a.relu && !(__ldg(a.y + o) > 0.f)) dz = 0.f;
                const float xh = (__ldg(a.x + o) + b - mu) * is;
                s0 += (double)dz; s1 += (double)dz * (double)xh;
            } else {
                s0 += (double)__ldg(a.x + o);
            }
        }
    }
    s0s[pl][cl] = s0; s1s[pl][cl] = s1;
    __syncthreads();
    if (pl == 0 && c < a.C) {
        for (int i = 1; i < 8; ++i) { s0 += s0s[i][cl]; s1 += s1s[i][cl]; }
        part[((size_t)blockIdx.y * 2 + 0) * a.C + c] = s0;
        part[((size_t)blockIdx.y * 2 + 1) * a.C + c] = s1;
    }
}

// Vector variant (C % 4 == 0, C >= 16): a thread owns one float4 of channels, a warp reads 32 consecutive float4 = 512 contiguous bytes
// of a row (or 2 / 4 shorter rows), rows are unrolled by two for memory-level parallelism.  Same partial-sum layout as above.
__global__ void __launch_bounds__(256) chan_reduce_vec_kernel(RedArgs a, int qb /*float4 lanes along the channels: min(C/4, 32)*/,
                                                              double* __restrict__ part) {
    __shared__ double sm[256][9];   // (+1: bank spread)
    const int tid = threadIdx.x, q = tid % qb, rl = tid / qb, rows = 256 / qb;
    const int c4 = blockIdx.x * qb + q, C4 = a.C >> 2, c = c4 * 4;
    const int m0 = blockIdx.y * a.slab, m1 = min(a.M, m0 + a.slab);
    double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    if (c4 < C4) {
        float b[4] = {0, 0, 0, 0}, mu[4] = {0, 0, 0, 0}, is[4] = {0, 0, 0, 0};
        for (int j = 0; j < 4; ++j) {
            if (a.bias) b[j] = __ldg(a.bias + c + j);
            if (a.mode == 1) { mu[j] = __ldg(a.mean + c + j); is[j] = __ldg(a.invstd + c + j); }
        }
        const float4* x4 = reinterpret_cast<const float4*>(a.x);
        const float4* d4 = reinterpret_cast<const float4*>(a.dy);
        const float4* y4 = reinterpret_cast<const float4*>(a.y);
        auto accum = [&](const float4& xv, const float4& dv, const float4& yv) {
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (a.mode == 0) {
                    const float v = xs[j] + b[j];
                    s0[j] += (double)v; s1[j] += (double)v * (double)v;
                } else if (a.mode == 1) {
                    const float dz = (a.relu && !(ys[j] > 0.f)) ? 0.f : ds[j];
                    const float xh = (xs[j] + b[j] - mu[j]) * is[j];
                    s0[j] += (double)dz; s1[j] += (double)dz * (double)xh;
                } else {
                    s0[j] += (double)xs[j];
                }
            }
        };
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int m = m0 + rl;
        for (; m + rows < m1; m += 2 * rows) {
            const size_t o0 = (size_t)m * C4 + c4, o1 = (size_t)(m + rows) * C4 + c4;
            const float4 xa = __ldg(x4 + o0), xb = __ldg(x4 + o1);
            const float4 da = a.mode == 1 ? __ldg(d4 + o0) : z4, db = a.mode == 1 ? __ldg(d4 + o1) : z4;
            const float4 ya = (a.mode == 1 && a.relu) ? __ldg(y4 + o0) : z4, yb = (a.mode == 1 && a.relu) ? __ldg(y4 + o1) : z4;
            accum(xa, da, ya);
            accum(xb, db, yb);
        }
        if (m < m1) {
            const size_t o0 = (size_t)m * C4 + c4;
            accum(__ldg(x4 + o0), a.mode == 1 ? __ldg(d4 + o0) : z4, (a.mode == 1 && a.relu) ? __ldg(y4 + o0) : z4);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { sm[tid][j] = s0[j]; sm[tid][4 + j] = s1[j]; }
    __syncthreads();
    if (rl == 0 && c4 < C4) {
        for (int r = 1; r < rows; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s0[j] += sm[r * qb + q][j]; s1[j] += sm[r * qb + q][4 + j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            part[((size_t)blockIdx.y * 2 + 0) * a.C + c + j] = s0[j];
            part[((size_t)blockIdx.y * 2 + 1) * a.C + c + j] = s1[j];
        }
    }
}

// finalize: out0/out1 from the partial sums.  mode 0: mean, biased variance.  mode 1: dbeta (= sum dz), dgamma (= sum dz*xhat).  mode 2: sum.
__global__ void __launch_bounds__(256) chan_finalize_kernel(const double* __restrict__ part, int slabs, int C, int M, int mode, float* __restrict__ out0,
                                                            float* __restrict__ out1) {
    // block = 32 channels x 8 slab lanes: the (up to several hundred) per-slab partial sums of a channel are added by 8 threads in parallel
    __shared__ double sm0[8][32], sm1[8][32];
    const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double s0 = 0.0, s1 = 0.0;
    if (c < C)
        for (int s = sl; s < slabs; s += 8) { s0 += part[((size_t)s * 2 + 0) * C + c]; s1 += part[((size_t)s * 2 + 1) * C + c]; }
    sm0[sl][cl] = s0; sm1[sl][cl] = s1;
    __syncthreads();
    if (sl != 0 || c >= C) return;
    for (int i = 1; i < 8; ++i) { s0 += sm0[i][cl]; s1 += sm1[i][cl]; }
    if (mode == 0) {
        const double mean = s0 / M;
        out0[c] = (float)mean;
        out1[c] = (float)fmax(s1 / M - mean * mean, 0.0);
    } else {
        out0[c] = (float)s0;
        if (out1) out1[c] = (float)s1;
    }
}

static int chan_reduce(RedArgs a, float* out0, float* out1, cudaStream_t st) {
    USOT_REQUIRE(a.M > 0 && a.C > 0, "empty reduction");
    const bool vec = a.C % 4 == 0 && a.C >= 16 && (a.C / 4 <= 32 ? 256 % (a.C / 4) == 0 : true);
    const int qb = vec ? std::min(a.C / 4, 32) : 0;
    const int groups = vec ? (a.C / 4 + qb - 1) / qb : (a.C + 31) / 32;
    int slabs = std::max(1, std::min((device_sm_count() * 4 + groups - 1) / groups, (a.M + 63) / 64));
    a.slab = (a.M + slabs - 1) / slabs;
    slabs = (a.M + a.slab - 1) / a.slab;
    double* part = nullptr;
    ensure_async_pool();
    USOT_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&part), (size_t)slabs * 2 * a.C * sizeof(double), st));
    if (vec) chan_reduce_vec_kernel<<<dim3(groups, slabs), 256, 0, st>>>(a, qb, part);
    else chan_reduce_kernel<<<dim3(groups, slabs), 256, 0, st>>>(a, part);
    chan_finalize_kernel<<<(a.C + 31) / 32, 256, 0, st>>>(part, slabs, a.C, a.M, a.mode, out0, out1);
    USOT_CUDA_OK(cudaGetLastError());
    USOT_CUDA_OK(cudaFreeAsync(part, st));
    return 0;
}

// y = (x + bias - mean) * invstd * gamma + beta (+ residual) ; optional ReLU.  invstd = rsqrt(var + eps) computed here in double.
__global__ void bn_apply_kernel(const float4* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ mean,
                                const float* __restrict__ var, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float4* __restrict__ residual, int relu, size_t total4, int C4, float4* __restrict__ y,
                                float* __restrict__ invstd_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)(i % C4) * 4;
    float4 v = __ldg(x + i);
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float is = (float)(1.0 / sqrt((double)__ldg(var + c + j) + (double)eps));
        if (invstd_out && i < (size_t)C4) invstd_out[c + j] = is;
        const float b = bias ? __ldg(bias + c + j) : 0.f;
        o[j] = (o[j] + b - __ldg(mean + c + j)) * is * __ldg(gamma + c + j) + __ldg(beta + c + j);
    }
    if (residual) { const float4 r = __ldg(residual + i); o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w; }
    if (relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    y[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// dx = gamma * invstd * (dz - [train] (dbeta + xhat * dgamma) / M) ; dres = dz (optional)
__global__ void bn_backward_apply_kernel(const float4* __restrict__ dy, const float4* __restrict__ yout, const float4* __restrict__ x,
                                         const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ invstd,
                                         const float* __restrict__ gamma, const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                         int train, int relu, float inv_m, size_t total4, int C4, float4* __restrict__ dx,
                                         float4* __restrict__ dres) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)(i % C4) * 4;
    const float4 g4 = __ldg(dy + i), x4 = __ldg(x + i);
    float dz[4] = {g4.x, g4.y, g4.z, g4.w};
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
    if (relu) {
        const float4 y4 = __ldg(yout + i);
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) if (!(yv[j] > 0.f)) dz[j] = 0.f;
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float is = __ldg(invstd + c + j), ga = __ldg(gamma + c + j);
        float t = dz[j];
        if (train) {
            const float b = bias ? __ldg(bias + c + j) : 0.f;
            const float xh = (xv[j] + b - __ldg(mean + c + j)) * is;
            t = dz[j] - (__ldg(dbeta + c + j) + xh * __ldg(dgamma + c + j)) * inv_m;
        }
        o[j] = ga * is * t;
    }
    dx[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (dres) dres[i] = make_float4(dz[0], dz[1], dz[2], dz[3]);
}

// =============================================================================================================================
// MaxPool 3x3 / stride 2 / pad 1 backward, NHWC: every input pixel gathers from the (<= 4) windows that contain it and whose FIRST
// maximum (row-major scan, strict >, as torch's max_pool2d) it is.
// =============================================================================================================================
__global__ void maxpool_backward_kernel(const float* __restrict__ in, const float* __restrict__ dout, int n, int h, int w, int c, int ho, int wo,
                                        float* __restrict__ din) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * h * w * c;
    if (idx >= total) return;
    const int cc = idx % c;
    size_t t = idx / c;
    const int x = t % w; t /= w;
    const int y = t % h;
    const int b = t / h;
    const float* base = in + (size_t)b * h * w * c + cc;
    float acc = 0.f;
    for (int oy = (y + 1) / 2 - ((y + 1) % 2 == 0 ? 1 : 0); oy <= (y + 1) / 2; ++oy) {   // windows with oy*2-1 <= y <= oy*2+1
        if (oy < 0 || oy >= ho) continue;
        for (int ox = (x + 1) / 2 - ((x + 1) % 2 == 0 ? 1 : 0); ox <= (x + 1) / 2; ++ox) {
            if (ox < 0 || ox >= wo) continue;
            float best = -FLT_MAX;
            int by = -1, bx = -1;
            for (int dy = 0; dy < 3; ++dy) {
                const int yy = oy * 2 - 1 + dy;
                if (yy < 0 || yy >= h) continue;
                for (int dx = 0; dx < 3; ++dx) {
                    const int xx = ox * 2 - 1 + dx;
                    if (xx < 0 || xx >= w) continue;
                    const float v = __ldg(base + ((size_t)yy * w + xx) * c);
                    if (v > best || by < 0) { best = v; by = yy; bx = xx; }
                }
            }
            if (by == y && bx == x) acc += __ldg(dout + (((size_t)b * ho + oy) * wo + ox) * c + cc);
        }
    }
    din[idx] = acc;
}

// =============================================================================================================================
// Conf_Fusion reduction backward.  e_q = exp(clamp(c_q,-6,4)); S = sum_q e_q; out = sum_q e_q v_q / S.
//   dv_q = dout * e_q / S ;  dc_q = dout * (v_q - out) * e_q / S * [ -6 < c_q < 4 ]     (torch.clamp passes gradient on the closed interval)
// =============================================================================================================================
__global__ void conf_fusion_backward_kernel(const float* __restrict__ conf, const float* __restrict__ value, const float* __restrict__ dout, int nq,
                                            size_t per_map, size_t total, float* __restrict__ dconf, float* __restrict__ dvalue) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t b = idx / per_map, e = idx % per_map;
    float s = 0.f, acc = 0.f;
    for (int q = 0; q < nq; ++q) {
        const size_t off = (b * nq + q) * per_map + e;
        const float ex = expf(fminf(fmaxf(__ldg(conf + off), -6.f), 4.f));
        s += ex;
        acc = fmaf(ex, __ldg(value + off), acc);
    }
    const float out = acc / s, g = __ldg(dout + idx);
    for (int q = 0; q < nq; ++q) {
        const size_t off = (b * nq + q) * per_map + e;
        const float cf = __ldg(conf + off), v = __ldg(value + off);
        const float wq = expf(fminf(fmaxf(cf, -6.f), 4.f)) / s;
        dvalue[off] = g * wq;
        dconf[off] = (cf >= -6.f && cf <= 4.f) ? g * (v - out) * wq : 0.f;
    }
}

// =============================================================================================================================
// out = w0*x0 + w1*x1 + w2*x2 (GroupDW's weighted sum) and its backward: dx_i = w_i * dout ; dw_i = <dout, x_i>
// =============================================================================================================================
__global__ void wsum3_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ w,
                             size_t n, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // same association as the reference's `s = 0; s = s + w[i] * x_i` loop (connect.py:96-102)
    out[i] = ((0.f + __ldg(w) * __ldg(x0 + i)) + __ldg(w + 1) * __ldg(x1 + i)) + __ldg(w + 2) * __ldg(x2 + i);
}

__global__ void __launch_bounds__(256) wsum3_backward_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ x2,
                                                             const float* __restrict__ w, const float* __restrict__ dout, size_t n,
                                                             float* __restrict__ dx0, float* __restrict__ dx1, float* __restrict__ dx2,
                                                             double* __restrict__ dw_acc /*[3], pre-zeroed*/) {
    __shared__ double sm[3][8];
    double a0 = 0, a1 = 0, a2 = 0;
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float g = __ldg(dout + i);
        a0 += (double)g * (double)__ldg(x0 + i); a1 += (double)g * (double)__ldg(x1 + i); a2 += (double)g * (double)__ldg(x2 + i);
        dx0[i] = w0 * g; dx1[i] = w1 * g; dx2[i] = w2 * g;
    }
    for (int off = 16; off > 0; off >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, off); a1 += __shfl_down_sync(0xffffffffu, a1, off); a2 += __shfl_down_sync(0xffffffffu, a2, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm[0][warp] = a0; sm[1][warp] = a1; sm[2][warp] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0;
        for (int i = 0; i < 8; ++i) s += sm[threadIdx.x][i];
        atomicAdd(dw_acc + threadIdx.x, s);
    }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// =============================================================================================================================
// Loss backward.  _weighted_BCE (models.py:42-58): L = 0.5 * mean_pos bce + 0.5 * mean_neg bce, a class with exactly one member
// contributes 0 (and gets no gradient).  d bce(x, y) / dx = sigmoid(x) - y.
// =============================================================================================================================
__global__ void __launch_bounds__(1024) bce_backward_kernel(const float* __restrict__ pred, const float* __restrict__ label, int count,
                                                            const float* __restrict__ gloss, float* __restrict__ dpred) {
    __shared__ int cnt[2];
    if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
    __syncthreads();
    int cp = 0, cn = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) { const float y = label[i]; cp += y == 1.f; cn += y == 0.f; }
    atomicAdd(&cnt[0], cp);
    atomicAdd(&cnt[1], cn);
    __syncthreads();
    const float g = __ldg(gloss);
    const float wp = cnt[0] > 1 ? 0.5f / (float)cnt[0] : 0.f, wn = cnt[1] > 1 ? 0.5f / (float)cnt[1] : 0.f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const float x = pred[i], y = label[i];
        const float sg = 1.f / (1.f + expf(-x));
        dpred[i] = y == 1.f ? g * wp * (sg - 1.f) : (y == 0.f ? g * wn * sg : 0.f);
    }
}

// IoU loss (models.py:60-100): L = mean over cells with weight > 0 of -log((I + 1) / (U + 1)).  bbox (n,4,R,R) nchw, target (n,R,R,4).
__global__ void __launch_bounds__(1024) iou_backward_kernel(const float* __restrict__ bbox, const float* __restrict__ target,
                                                            const float* __restrict__ weight, int n, int cells, const float* __restrict__ gloss,
                                                            float* __restrict__ dbbox) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int c0 = 0;
    for (int i = threadIdx.x; i < n * cells; i += blockDim.x) c0 += weight[i] > 0.f;
    atomicAdd(&cnt, c0);
    __syncthreads();
    const float g = __ldg(gloss) / (float)cnt;
    for (int i = threadIdx.x; i < n * cells; i += blockDim.x) {
        const int b = i / cells, c = i % cells;
        float* d = dbbox + (size_t)b * 4 * cells + c;
        if (!(weight[i] > 0.f)) { d[0] = 0.f; d[cells] = 0.f; d[2 * cells] = 0.f; d[3 * cells] = 0.f; continue; }
        const float* p = bbox + (size_t)b * 4 * cells + c;
        const float pl = p[0], pt = p[cells], pr = p[2 * cells], pb = p[3 * cells];
        const float* t = target + (size_t)i * 4;
        const float tl = t[0], tt = t[1], tr = t[2], tb = t[3];
        const float wi = fminf(pl, tl) + fminf(pr, tr), hi = fminf(pb, tb) + fminf(pt, tt);
        const float ai = wi * hi, pa = (pl + pr) * (pt + pb), ta = (tl + tr) * (tt + tb);
        const float au = ta + pa - ai;
        // L = log(U+1) - log(I+1); dL/dI = -1/(I+1) - 1/(U+1) (U depends on I with coefficient -1); dL/dPA = 1/(U+1)
        const float dI = -1.f / (ai + 1.f) - 1.f / (au + 1.f), dPA = 1.f / (au + 1.f);
        // torch.min(a, b) routes the gradient to a where a <= b (ties: torch gives 0.5 to each; measure-zero for float data)
        const float ml = pl < tl ? 1.f : (pl == tl ? 0.5f : 0.f), mr = pr < tr ? 1.f : (pr == tr ? 0.5f : 0.f);
        const float mt = pt < tt ? 1.f : (pt == tt ? 0.5f : 0.f), mb = pb < tb ? 1.f : (pb == tb ? 0.5f : 0.f);
        d[0] = g * (dI * hi * ml + dPA * (pt + pb));
        d[2 * cells] = g * (dI * hi * mr + dPA * (pt + pb));
        d[cells] = g * (dI * wi * mt + dPA * (pl + pr));
        d[3 * cells] = g * (dI * wi * mb + dPA * (pl + pr));
    }
}

}  // namespace usot

using namespace usot;

// =============================================================================================================================
// C ABI
// =============================================================================================================================
extern "C" {

static int make_geom(ConvGeom* g, int n, int h, int w, int cin, int cout, int kh, int kw, int stride, int ph, int pw, int dh, int dw) {
    USOT_REQUIRE(n >= 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && kh > 0 && kw > 0 && stride > 0 && ph >= 0 && pw >= 0 && dh > 0 && dw > 0, "bad conv shape");
    *g = ConvGeom{n, h, w, cin, cout, kh, kw, stride, ph, pw, dh, dw, conv_out(h, kh, stride, ph, dh), conv_out(w, kw, stride, pw, dw)};
    USOT_REQUIRE(g->ho > 0 && g->wo > 0, "conv output is empty");
    return 0;
}

int usot_conv2d_wgrad_nhwc(const float* in, const float* grad_out, int n, int h, int w, int cin, int cout, int kh, int kw, int stride, int pad_h,
                           int pad_w, int dil_h, int dil_w, float* grad_weight_kn, int precision, void* stream) {
    USOT_REQUIRE(grad_weight_kn && (n == 0 || (in && grad_out)), "null pointer");
    USOT_REQUIRE(precision >= USOT_PREC_FP32_SIMT && precision <= USOT_PREC_FP16_TC, "unknown precision mode");
    ConvGeom g;
    if (int rc = make_geom(&g, n, h, w, cin, cout, kh, kw, stride, pad_h, pad_w, dil_h, dil_w)) return rc;
    count_op_launch(OPFAM_WGRAD, 1);
    if (precision != USOT_PREC_FP32_SIMT && wgrad_tc_supported(g))
        return launch_conv_wgrad_tc(in, grad_out, g, grad_weight_kn, precision == USOT_PREC_FP16X3_TC, (cudaStream_t)stream);
    return launch_conv_wgrad(in, grad_out, g, grad_weight_kn, (cudaStream_t)stream);
}

int usot_stem_conv_wgrad(const float* x, const float* grad_out, int n, int size, float* grad_weight_kn, void* stream) {
    USOT_REQUIRE(grad_weight_kn && (n == 0 || (x && grad_out)), "null pointer");
    USOT_REQUIRE(n >= 0 && size >= 7, "bad shape");
    count_op_launch(OPFAM_WGRAD, 1);
    return launch_stem_wgrad(x, grad_out, n, size, grad_weight_kn, (cudaStream_t)stream);
}

int usot_pow2_scale(const float* x, int64_t numel, int target_log2, float* y, float* scale2, void* stream) {
    USOT_REQUIRE(x && scale2 && numel > 0 && numel % 4 == 0, "bad argument");
    USOT_REQUIRE(target_log2 >= -20 && target_log2 <= 14, "target exponent out of range");
    count_op_launch(OPFAM_TRAIN, y ? 3 : 2);
    return launch_pow2_scale(x, (size_t)numel, target_log2, y, scale2, (cudaStream_t)stream);
}

int usot_conv2d_dgrad_nhwc(const float* grad_out, const float* weight_kn, int n, int h, int w, int cin, int cout, int kh, int kw, int stride,
                           int pad_h, int pad_w, int dil_h, int dil_w, float* grad_in, void* stream) {
    USOT_REQUIRE(n == 0 || (grad_out && weight_kn && grad_in), "null pointer");
    ConvGeom g;
    if (int rc = make_geom(&g, n, h, w, cin, cout, kh, kw, stride, pad_h, pad_w, dil_h, dil_w)) return rc;
    count_op_launch(OPFAM_TRAIN, 1);
    return launch_conv_dgrad_gather(grad_out, weight_kn, g, grad_in, (cudaStream_t)stream);
}

int usot_bn_stats(const float* x, const float* bias, int64_t m, int channels, float* mean, float* var, void* stream) {
    USOT_REQUIRE(x && mean && var && m > 0 && m < (1ll << 31) && channels > 0, "bad argument");
    count_op_launch(OPFAM_TRAIN, 2);
    RedArgs a{x, bias, nullptr, nullptr, nullptr, nullptr, (int)m, channels, 0, 0, 0};
    return chan_reduce(a, mean, var, (cudaStream_t)stream);
}

int usot_bn_apply(const float* x, const float* bias, const float* mean, const float* var, float eps, const float* gamma, const float* beta,
                  const float* residual, int relu, int64_t m, int channels, float* y, float* invstd_out, void* stream) {
    USOT_REQUIRE(x && mean && var && gamma && beta && y && m > 0 && channels > 0 && channels % 4 == 0, "bad argument");
    count_op_launch(OPFAM_TRAIN, 1);
    const size_t total4 = (size_t)m * channels / 4;
    bn_apply_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), bias, mean, var, eps, gamma, beta, reinterpret_cast<const float4*>(residual), relu, total4, channels / 4,
        reinterpret_cast<float4*>(y), invstd_out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int usot_bn_backward(const float* grad_y, const float* y, const float* x, const float* bias, const float* mean, const float* invstd,
                     const float* gamma, int train, int relu, int64_t m, int channels, float* grad_x, float* grad_gamma, float* grad_beta,
                     float* grad_residual, void* stream) {
    USOT_REQUIRE(grad_y && x && mean && invstd && gamma && grad_x && grad_gamma && grad_beta && (!relu || y), "null pointer");
    count_op_launch(OPFAM_TRAIN, 3);
    USOT_REQUIRE(m > 0 && m < (1ll << 31) && channels > 0 && channels % 4 == 0, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    RedArgs a{x, bias, grad_y, y, mean, invstd, (int)m, channels, 1, relu, 0};
    if (int rc = chan_reduce(a, grad_beta, grad_gamma, st)) return rc;
    const size_t total4 = (size_t)m * channels / 4;
    bn_backward_apply_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(grad_y), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(x), bias, mean, invstd, gamma,
        grad_beta, grad_gamma, train, relu, 1.0f / (float)m, total4, channels / 4, reinterpret_cast<float4*>(grad_x),
        reinterpret_cast<float4*>(grad_residual));
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int usot_channel_sum(const float* x, int64_t m, int channels, float* out, void* stream) {
    USOT_REQUIRE(x && out && m > 0 && m < (1ll << 31) && channels > 0, "bad argument");
    count_op_launch(OPFAM_TRAIN, 2);
    RedArgs a{x, nullptr, nullptr, nullptr, nullptr, nullptr, (int)m, channels, 2, 0, 0};
    return chan_reduce(a, out, nullptr, (cudaStream_t)stream);
}

int usot_maxpool3x3s2p1_backward_nhwc(const float* in, const float* grad_out, int n, int h, int w, int channels, float* grad_in, void* stream) {
    USOT_REQUIRE(n == 0 || (in && grad_out && grad_in), "null pointer");
    count_op_launch(OPFAM_TRAIN, 1);
    USOT_REQUIRE(n >= 0 && h > 0 && w > 0 && channels > 0, "bad shape");
    const size_t total = (size_t)n * h * w * channels;
    if (total == 0) return 0;
    maxpool_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, grad_out, n, h, w, channels, (h - 1) / 2 + 1,
                                                                                           (w - 1) / 2 + 1, grad_in);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int usot_conf_fusion_backward(const float* conf, const float* value, const float* grad_out, int batch, int nq, int64_t per_map, float* grad_conf,
                              float* grad_value, void* stream) {
    USOT_REQUIRE(batch == 0 || (conf && value && grad_out && grad_conf && grad_value), "null pointer");
    count_op_launch(OPFAM_TRAIN, 1);
    USOT_REQUIRE(batch >= 0 && nq > 0 && per_map > 0, "bad shape");
    const size_t total = (size_t)batch * per_map;
    if (total == 0) return 0;
    conf_fusion_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(conf, value, grad_out, nq, (size_t)per_map, total,
                                                                                               grad_conf, grad_value);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int usot_weighted_sum3(const float* x0, const float* x1, const float* x2, const float* w3, int64_t numel, float* out, void* stream) {
    USOT_REQUIRE(x0 && x1 && x2 && w3 && out && numel > 0, "bad argument");
    count_op_launch(OPFAM_TRAIN, 1);
    wsum3_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x0, x1, x2, w3, (size_t)numel, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int usot_weighted_sum3_backward(const float* x0, const float* x1, const float* x2, const float* w3, const float* grad_out, int64_t numel,
                                float* grad_x0, float* grad_x1, float* grad_x2, float* grad_w3, void* stream) {
    USOT_REQUIRE(x0 && x1 && x2 && w3 && grad_out && grad_x0 && grad_x1 && grad_x2 && grad_w3 && numel > 0, "bad argument");
    count_op_launch(OPFAM_TRAIN, 2);
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = nullptr;
    ensure_async_pool();
    USOT_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&acc), 3 * sizeof(double), st));
    USOT_CUDA_OK(cudaMemsetAsync(acc, 0, 3 * sizeof(double), st));
    const unsigned blocks = (unsigned)std::min<int64_t>((numel + 255) / 256, (int64_t)device_sm_count() * 8);
    wsum3_backward_kernel<<<blocks, 256, 0, st>>>(x0, x1, x2, w3, grad_out, (size_t)numel, grad_x0, grad_x1, grad_x2, acc);
    f64_to_f32_kernel<<<1, 32, 0, st>>>(acc, 3, grad_w3);
    USOT_CUDA_OK(cudaGetLastError());
    USOT_CUDA_OK(cudaFreeAsync(acc, st));
    return 0;
}

int usot_weighted_bce_backward(const float* pred, const float* label, int count, const float* grad_loss, float* grad_pred, void* stream) {
    USOT_REQUIRE(pred && label && grad_loss && grad_pred && count > 0, "bad argument");
    count_op_launch(OPFAM_TRAIN, 1);
    bce_backward_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, label, count, grad_loss, grad_pred);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int usot_iou_loss_backward(const float* bbox, const float* reg_target, const float* reg_weight, int n, int cells, const float* grad_loss,
                           float* grad_bbox, void* stream) {
    USOT_REQUIRE(bbox && reg_target && reg_weight && grad_loss && grad_bbox && n > 0 && cells > 0, "bad argument");
    count_op_launch(OPFAM_TRAIN, 1);
    iou_backward_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(bbox, reg_target, reg_weight, n, cells, grad_loss, grad_bbox);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
