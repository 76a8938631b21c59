// fp32 SIMT kernels of the USOT forward path (NHWC activations).
//
//   stem_kernel          conv1 7x7/2 + BN + ReLU            lib/models/modules.py:70-74,138-140
//   maxpool_kernel       maxpool 3x3/2 p1                   lib/models/modules.py:74,141
//   conv_simt_kernel     dense conv (1x1 / 3x3, any stride/pad/dilation) as an fp32 implicit GEMM with the
//                        folded-BN / bias / residual / ReLU epilogue   modules.py:37-58, connect.py:20-53,112-121,178-209
//   groupdw_kernel       fused 3-scale depth-wise xcorr + softmax-weighted sum   connect.py:86-102,147-157
//   xcorr_nchw_kernel    the stand-alone reference op, NCHW   connect.py:147-157
//   pred_conv_kernel     Cout<=4 prediction convs + exp/scale epilogue   connect.py:212-219,236-241,275
//   conf_fusion_kernel   clamp/exp/normalise/weighted-sum over N_q   connect.py:128-142
//   prroi_*_kernel       PrRoIPool forward   prroi_pool/src/prroi_pooling_gpu_impl.cu:149-212
//
// The dense convs here are the exact-fp32 fallback and on-device cross-check for the tcgen05 path (conv_tc.cu).
#include "common.cuh"
#include <cfloat>

namespace usot {

static __device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// =============================================================================================
// Stem: NCHW fp32 (n,3,S,S) -> NHWC (n,ho,ho,64), 7x7 stride 2 pad 0, folded BN + ReLU.
// Block = 16x16 output pixels; thread = 4 consecutive pixels x 16 channels.
// =============================================================================================
constexpr int STEM_T = 16;
constexpr int STEM_P = STEM_T * 2 + 5;  // 37 input rows/cols per tile

__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ x, int S, int HO, const float* __restrict__ wk,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   float* __restrict__ out, int relu) {
    extern __shared__ float smem[];
    float* ws = smem;                   // [147][64]
    float* ps = smem + 147 * 64;        // [3][37][37]
    const int n = blockIdx.z, ty0 = blockIdx.y * STEM_T, tx0 = blockIdx.x * STEM_T;
    const int tid = threadIdx.x;
    for (int i = tid; i < 147 * 64 / 4; i += 256) reinterpret_cast<float4*>(ws)[i] = ldg4(wk + i * 4);
    const float* xn = x + (size_t)n * 3 * S * S;
    for (int i = tid; i < 3 * STEM_P * STEM_P; i += 256) {
        int c = i / (STEM_P * STEM_P), r = (i / STEM_P) % STEM_P, q = i % STEM_P;
        int gy = ty0 * 2 + r, gx = tx0 * 2 + q;
        ps[i] = (gy < S && gx < S) ? __ldg(xn + ((size_t)c * S + gy) * S + gx) : 0.f;
    }
    __syncthreads();
    const int q = tid >> 2, cg = tid & 3, row = q >> 2, colq = q & 3;
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[p][j] = 0.f;
    for (int c = 0; c < 3; ++c) {
#pragma unroll 1
        for (int kh = 0; kh < 7; ++kh) {
            float in[13];
            const float* pr = ps + (c * STEM_P + row * 2 + kh) * STEM_P + colq * 8;
#pragma unroll
            for (int i = 0; i < 13; ++i) in[i] = pr[i];
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
                const float4* wp = reinterpret_cast<const float4*>(ws + ((c * 7 + kh) * 7 + kw) * 64 + cg * 16);
                float w[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 v = wp[j];
                    w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
                }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(in[2 * p + kw], w[j], acc[p][j]);
            }
        }
    }
    const int oy = ty0 + row;
    if (oy >= HO) return;
    float sc[16], sh[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { sc[j] = scale ? __ldg(scale + cg * 16 + j) : 1.f; sh[j] = shift ? __ldg(shift + cg * 16 + j) : 0.f; }
    const float lo = relu ? 0.f : -FLT_MAX;  // (raw conv output for the training path: no ReLU, identity affine)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int ox = tx0 + colq * 4 + p;
        if (ox >= HO) continue;
        float4* o = reinterpret_cast<float4*>(out + (((size_t)n * HO + oy) * HO + ox) * 64 + cg * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 v;
            v.x = fmaxf(fmaf(acc[p][4 * j], sc[4 * j], sh[4 * j]), lo);
            v.y = fmaxf(fmaf(acc[p][4 * j + 1], sc[4 * j + 1], sh[4 * j + 1]), lo);
            v.z = fmaxf(fmaf(acc[p][4 * j + 2], sc[4 * j + 2], sh[4 * j + 2]), lo);
            v.w = fmaxf(fmaf(acc[p][4 * j + 3], sc[4 * j + 3], sh[4 * j + 3]), lo);
            o[j] = v;
        }
    }
}

int launch_stem(const float* x, int n, int s, const float* wk, const float* scale, const float* shift, float* out,
                cudaStream_t st, int relu) {
    const int ho = (s - 7) / 2 + 1;
    const size_t smem = (147 * 64 + 3 * STEM_P * STEM_P) * sizeof(float);
    static SmemAttrCache attr;
    if (int rc = attr.ensure(stem_kernel, (int)smem)) return rc;
    dim3 grid((ho + STEM_T - 1) / STEM_T, (ho + STEM_T - 1) / STEM_T, n);
    stem_kernel<<<grid, 256, smem, st>>>(x, s, ho, wk, scale, shift, out, relu);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// MaxPool 3x3 stride 2 pad 1, NHWC, one float4 of channels per thread.
// =============================================================================================
// out != nullptr: fp32 result.  hi != nullptr: the result leaves as the split-fp16 planes the tensor-core convs read (lo may be
// null in single-fp16 mode), which saves the fp32 round trip through HBM and the separate conversion launch.
__global__ void maxpool_kernel(const float* __restrict__ in, int n, int h, int w, int c4, int ho, int wo,
                               float* __restrict__ out, uint2* __restrict__ hi, uint2* __restrict__ lo) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n * ho * wo * c4;
    if (idx >= total) return;
    int cc = idx % c4;
    size_t t = idx / c4;
    int ox = t % wo; t /= wo;
    int oy = t % ho;
    int b = t / ho;
    float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        int y = oy * 2 - 1 + dy;
        if (y < 0 || y >= h) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            int xx = ox * 2 - 1 + dx;
            if (xx < 0 || xx >= w) continue;
            float4 v = ldg4(in + (((size_t)b * h + y) * w + xx) * c4 * 4 + cc * 4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    if (out) reinterpret_cast<float4*>(out)[idx] = m;
    if (hi) {
        const __half2 h0 = __floats2half2_rn(m.x, m.y), h1 = __floats2half2_rn(m.z, m.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(m.x - f0.x, m.y - f0.y), l1 = __floats2half2_rn(m.z - f1.x, m.w - f1.y);
        uint2 a, b;
        a.x = *reinterpret_cast<const uint32_t*>(&h0); a.y = *reinterpret_cast<const uint32_t*>(&h1);
        b.x = *reinterpret_cast<const uint32_t*>(&l0); b.y = *reinterpret_cast<const uint32_t*>(&l1);
        hi[idx] = a;
        if (lo) lo[idx] = b;
    }
}

int launch_maxpool3x3s2p1(const float* in, int n, int h, int w, int c, float* out, __half* out_hi, __half* out_lo, cudaStream_t st) {
    USOT_REQUIRE(c % 4 == 0, "maxpool needs C % 4 == 0");
    USOT_REQUIRE(out || out_hi, "maxpool needs an output");
    int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    size_t total = (size_t)n * ho * wo * (c / 4);
    maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, n, h, w, c / 4, ho, wo, out, reinterpret_cast<uint2*>(out_hi),
                                                                    reinterpret_cast<uint2*>(out_lo));
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// Dense conv as fp32 implicit GEMM.  M = n*ho*wo pixels, N = cout, K = kh*kw*cin (cin % 16 == 0).
// 128 x (64*NH) x 16 tiles, 256 threads, 8 x (4*NH) outputs per thread, register-staged double buffer.
// =============================================================================================
constexpr int CS_BM = 128, CS_BK = 16, CS_LDA = CS_BM + 4;

template <int NH>
__global__ void __launch_bounds__(256, 2) conv_simt_kernel(const float* __restrict__ in, ConvGeom g, const float* __restrict__ wkn,
                                                        Epilogue ep, float* __restrict__ out) {
    constexpr int BN = 64 * NH;
    __shared__ __align__(16) float As[2][CS_BK][CS_LDA];
    __shared__ __align__(16) float Bs[2][CS_BK][BN];
    const int tid = threadIdx.x;
    const int M = g.n * g.ho * g.wo;
    const int m0 = blockIdx.x * CS_BM, n0 = blockIdx.y * BN;
    const int nK = g.kh * g.kw * g.cin / CS_BK;

    // ---- A loader state: rows (tid/4) and (tid/4 + 64), 4 floats at k-offset (tid%4)*4 ----
    const int a_kq = tid & 3;
    int a_hi0[2], a_wi0[2];
    const float* a_base[2];
    bool a_rowok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int m = m0 + (tid >> 2) + r * 64;
        a_rowok[r] = m < M;
        int mm = a_rowok[r] ? m : 0;
        int b = mm / (g.ho * g.wo), rem = mm % (g.ho * g.wo);
        int oy = rem / g.wo, ox = rem % g.wo;
        a_hi0[r] = oy * g.stride - g.ph;
        a_wi0[r] = ox * g.stride - g.pw;
        a_base[r] = in + (size_t)b * g.h * g.w * g.cin;
    }
    // ---- B loader state: rows (tid/(BN/4)) + i*(1024/BN), one float4 each ----
    constexpr int B_TPR = BN / 4;            // threads per k-row
    constexpr int B_ROWS = 256 / B_TPR;      // rows per pass
    constexpr int B_PASS = CS_BK / B_ROWS;   // passes
    const int b_row = tid / B_TPR, b_col = (tid % B_TPR) * 4;

    int k_c0 = 0, k_kh = 0, k_kw = 0;  // decomposition of the NEXT chunk to load
    float4 a_reg[2], b_reg[B_PASS];

    auto load_chunk = [&](int kc) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int hi = a_hi0[r] + k_kh * g.dh, wi = a_wi0[r] + k_kw * g.dw;
            bool ok = a_rowok[r] && hi >= 0 && hi < g.h && wi >= 0 && wi < g.w;
            a_reg[r] = ok ? ldg4(a_base[r] + ((size_t)hi * g.w + wi) * g.cin + k_c0 + a_kq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int p = 0; p < B_PASS; ++p)
            b_reg[p] = ldg4(wkn + (size_t)(kc * CS_BK + b_row + p * B_ROWS) * g.cout + n0 + b_col);
        k_c0 += CS_BK;
        if (k_c0 == g.cin) { k_c0 = 0; if (++k_kw == g.kw) { k_kw = 0; ++k_kh; } }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int row = (tid >> 2) + r * 64;
            As[buf][a_kq * 4 + 0][row] = a_reg[r].x;
            As[buf][a_kq * 4 + 1][row] = a_reg[r].y;
            As[buf][a_kq * 4 + 2][row] = a_reg[r].z;
            As[buf][a_kq * 4 + 3][row] = a_reg[r].w;
        }
#pragma unroll
        for (int p = 0; p < B_PASS; ++p) *reinterpret_cast<float4*>(&Bs[buf][b_row + p * B_ROWS][b_col]) = b_reg[p];
    };

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][4 * NH];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NH; ++j) acc[i][j] = 0.f;

    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int kc = 0; kc < nK; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nK) load_chunk(kc + 1);
#pragma unroll
        for (int k = 0; k < CS_BK; ++k) {
            float a[8], b[4 * NH];
            float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
            v = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float4 u = *reinterpret_cast<const float4*>(&Bs[buf][k][h * 64 + tx * 4]);
                b[4 * h] = u.x; b[4 * h + 1] = u.y; b[4 * h + 2] = u.z; b[4 * h + 3] = u.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4 * NH; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nK) store_chunk(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----
#pragma unroll
    for (int h = 0; h < NH; ++h) {
        const int col = n0 + h * 64 + tx * 4;
        const float4 sc = ldg4(ep.scale + col), sh = ldg4(ep.shift + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
            if (m >= M) continue;
            float4 y;
            y.x = fmaf(acc[i][4 * h + 0], sc.x, sh.x);
            y.y = fmaf(acc[i][4 * h + 1], sc.y, sh.y);
            y.z = fmaf(acc[i][4 * h + 2], sc.z, sh.z);
            y.w = fmaf(acc[i][4 * h + 3], sc.w, sh.w);
            if (ep.residual) {
                float4 r = ldg4(ep.residual + (size_t)m * g.cout + col);
                y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
            }
            if (ep.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
            *reinterpret_cast<float4*>(out + (size_t)m * g.cout + col) = y;
        }
    }
}

int launch_conv_simt(const float* in, const ConvGeom& g, const float* w_kn, const Epilogue& ep, float* out, cudaStream_t st) {
    USOT_REQUIRE(g.cin % CS_BK == 0, "conv_simt needs Cin % 16 == 0");
    USOT_REQUIRE(g.cout % 64 == 0, "conv_simt needs Cout % 64 == 0");
    const int M = g.n * g.ho * g.wo;
    if (M == 0) return 0;
    if (g.cout % 128 == 0) {
        dim3 grid((M + CS_BM - 1) / CS_BM, g.cout / 128);
        conv_simt_kernel<2><<<grid, 256, 0, st>>>(in, g, w_kn, ep, out);
    } else {
        dim3 grid((M + CS_BM - 1) / CS_BM, g.cout / 64);
        conv_simt_kernel<1><<<grid, 256, 0, st>>>(in, g, w_kn, ep, out);
    }
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// Fused GroupDW (bandwidth-bound).  One thread per (output sample, channel, column strip); lanes run
// over consecutive channels so every global access is a full 128-byte line.  Each x value is read from
// HBM exactly once and lives in registers; a ring of 5 output rows is kept in registers and the three
// correlations (5x5 on x11, 3x5 on x12, 5x3 on x21) accumulate into it, taps pre-scaled by softmax(w).
// =============================================================================================
template <int SW>
__global__ void __launch_bounds__(64) groupdw_kernel(GroupDWArgs a, int nstrips, float w0, float w1, float w2) {
    __shared__ float zs[55][64];
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int strip = bid % nstrips; bid /= nstrips;
    const int cblocks = a.C / 64;
    const int cblk = bid % cblocks;
    const int n = bid / cblocks;
    const int c = cblk * 64 + tid;
    const int rep = a.n_out / a.nx;
    const int zb = n / (a.n_out / a.nz);  // nz == n_out: own kernel; nz == 1: broadcast; otherwise each kernel serves n_out/nz samples
    const int xb = n / rep;
    const int C = a.C, F = a.F, R = F - 6;
    for (int t = 0; t < 25; ++t) zs[t][tid] = w0 * __ldg(a.z11 + ((size_t)zb * 25 + t) * C + c);
    for (int t = 0; t < 15; ++t) zs[25 + t][tid] = w1 * __ldg(a.z12 + ((size_t)zb * 15 + t) * C + c);
    for (int t = 0; t < 15; ++t) zs[40 + t][tid] = w2 * __ldg(a.z21 + ((size_t)zb * 15 + t) * C + c);
    // (no __syncthreads needed: each thread only ever reads the column it wrote)
    const int j0 = strip * SW;
    const int jn = min(SW, R - j0);
    const int W11 = F - 2, H11 = F - 2, W12 = F - 2, H12 = F - 4, W21 = F - 4;
    const float* x11 = a.x11 + (size_t)xb * H11 * W11 * C + c;
    const float* x12 = a.x12 + (size_t)xb * H12 * W12 * C + c;
    const float* x21 = a.x21 + (size_t)xb * H11 * W21 * C + c;
    float* out = a.out + (size_t)n * R * R * C + c;

    float acc[5][SW];  // acc[k] = output row t-4+k at step t
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int j = 0; j < SW; ++j) acc[k][j] = 0.f;

    for (int t = 0; t < H11; ++t) {
        {   // 5x5 on x11 row t -> output rows t-u
            float xr[SW + 4];
#pragma unroll
            for (int q = 0; q < SW + 4; ++q) xr[q] = (j0 + q < W11) ? __ldg(x11 + ((size_t)t * W11 + j0 + q) * C) : 0.f;
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const float z = zs[u * 5 + v][tid];
#pragma unroll
                    for (int j = 0; j < SW; ++j) acc[4 - u][j] = fmaf(xr[j + v], z, acc[4 - u][j]);
                }
        }
        {   // 5x3 on x21 row t -> output rows t-u
            float xr[SW + 2];
#pragma unroll
            for (int q = 0; q < SW + 2; ++q) xr[q] = (j0 + q < W21) ? __ldg(x21 + ((size_t)t * W21 + j0 + q) * C) : 0.f;
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    const float z = zs[40 + u * 3 + v][tid];
#pragma unroll
                    for (int j = 0; j < SW; ++j) acc[4 - u][j] = fmaf(xr[j + v], z, acc[4 - u][j]);
                }
        }
        if (t >= 2 && t - 2 < H12) {  // 3x5 on x12 row t-2 -> output rows t-2-u  (ring slots 2-u)
            float xr[SW + 4];
#pragma unroll
            for (int q = 0; q < SW + 4; ++q) xr[q] = (j0 + q < W12) ? __ldg(x12 + ((size_t)(t - 2) * W12 + j0 + q) * C) : 0.f;
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const float z = zs[25 + u * 5 + v][tid];
#pragma unroll
                    for (int j = 0; j < SW; ++j) acc[2 - u][j] = fmaf(xr[j + v], z, acc[2 - u][j]);
                }
        }
        if (t >= 4) {  // output row t-4 is complete
            float* o = out + ((size_t)(t - 4) * R + j0) * C;
#pragma unroll
            for (int j = 0; j < SW; ++j)
                if (j < jn) o[(size_t)j * C] = acc[0][j];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < SW; ++j) acc[k][j] = acc[k + 1][j];
#pragma unroll
        for (int j = 0; j < SW; ++j) acc[4][j] = 0.f;
    }
}

int launch_groupdw(const GroupDWArgs& a, cudaStream_t st) {
    USOT_REQUIRE(a.C % 64 == 0, "groupdw needs C % 64 == 0");
    USOT_REQUIRE(a.nx > 0 && a.n_out % a.nx == 0, "groupdw: n_out must be a multiple of nx");
    USOT_REQUIRE(a.nz > 0 && a.n_out % a.nz == 0, "groupdw: n_out must be a multiple of the kernel batch");
    USOT_REQUIRE(a.F >= 7, "groupdw: feature size too small");
    float w[3];
    USOT_CUDA_OK(cudaMemcpyAsync(w, a.dw_weight, sizeof(w), cudaMemcpyDeviceToHost, st));
    USOT_CUDA_OK(cudaStreamSynchronize(st));
    float mx = fmaxf(w[0], fmaxf(w[1], w[2]));
    float e0 = expf(w[0] - mx), e1 = expf(w[1] - mx), e2 = expf(w[2] - mx), s = e0 + e1 + e2;
    return launch_groupdw_w(a, e0 / s, e1 / s, e2 / s, st);
}

Tunable g_groupdw_strips = 3;  // tunable (usot_set_tunable("groupdw_strips", 2|3)) of the register-staged variant
Tunable g_groupdw_tma = 2;     // tunable: 2 = TMA ring + packed FFMA2 (xcorr_tma.cu, default), 1 = TMA ring + scalar FMA, 0 = register-staged kernel below

bool groupdw_split_output_supported(int F) { return g_groupdw_tma >= 2 && (F - 6 + 8) / 9 == 3; }

int launch_groupdw_w(const GroupDWArgs& a, float w0, float w1, float w2, cudaStream_t st) {
    USOT_REQUIRE(a.nx > 0 && a.nz > 0 && a.n_out % a.nx == 0 && a.n_out % a.nz == 0, "groupdw: n_out must be a multiple of both batches");
    const int R = a.F - 6;
    if (g_groupdw_tma && (R + 8) / 9 == 3) return launch_groupdw_tma(a, w0, w1, w2, st);
    USOT_REQUIRE(!a.out_hi, "split-fp16 GroupDW output needs the FFMA2 kernel (see groupdw_split_output_supported)");
    const int nstrips = (g_groupdw_strips == 2 && R <= 28) ? 2 : (R + 8) / 9;
    const int sw = (R + nstrips - 1) / nstrips;
    const unsigned grid = (unsigned)(a.n_out * (a.C / 64) * nstrips);
    if (grid == 0) return 0;
    if (sw <= 9) groupdw_kernel<9><<<grid, 64, 0, st>>>(a, nstrips, w0, w1, w2);
    else if (sw <= 13) groupdw_kernel<13><<<grid, 64, 0, st>>>(a, nstrips, w0, w1, w2);
    else if (sw <= 14) groupdw_kernel<14><<<grid, 64, 0, st>>>(a, nstrips, w0, w1, w2);
    else { USOT_REQUIRE(false, "groupdw: unsupported response size"); }
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// Stand-alone depth-wise xcorr in the reference's NCHW layout: one block per (sample, channel) plane.
// =============================================================================================
__global__ void __launch_bounds__(256) xcorr_nchw_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                         float* __restrict__ out, int nk, int C, int hx, int wx, int hk, int wk) {
    extern __shared__ float sm[];
    float* xs = sm;
    float* ks = sm + hx * wx;
    const int plane = blockIdx.x;  // b*C + c
    const int b = plane / C, c = plane % C;
    const int kb = (nk == 1) ? 0 : b;
    const float* xp = x + (size_t)plane * hx * wx;
    const float* kp = k + ((size_t)kb * C + c) * hk * wk;
    for (int i = threadIdx.x; i < hx * wx; i += 256) xs[i] = __ldg(xp + i);
    for (int i = threadIdx.x; i < hk * wk; i += 256) ks[i] = __ldg(kp + i);
    __syncthreads();
    const int ho = hx - hk + 1, wo = wx - wk + 1;
    for (int o = threadIdx.x; o < ho * wo; o += 256) {
        int oy = o / wo, ox = o % wo;
        float s = 0.f;
        for (int u = 0; u < hk; ++u)
            for (int v = 0; v < wk; ++v) s = fmaf(xs[(oy + u) * wx + ox + v], ks[u * wk + v], s);
        out[(size_t)plane * ho * wo + o] = s;
    }
}

int launch_xcorr_nchw(const float* x, const float* k, float* out, int nx, int nk, int C, int hx, int wx, int hk, int wk,
                      cudaStream_t st) {
    USOT_REQUIRE(nk == 1 || nk == nx, "xcorr: kernel batch must be 1 or equal to the search batch");
    USOT_REQUIRE(hx >= hk && wx >= wk, "xcorr: kernel larger than search map");
    size_t smem = ((size_t)hx * wx + (size_t)hk * wk) * sizeof(float);
    USOT_REQUIRE(smem <= 48 * 1024, "xcorr: plane too large");
    if (nx * C == 0) return 0;
    xcorr_nchw_kernel<<<nx * C, 256, smem, st>>>(x, k, out, nk, C, hx, wx, hk, wk);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// Conf_Fusion epilogue: out[b] = sum_q exp(clamp(conf[b,q],-6,4)) * value[b,q] / sum_q exp(clamp(conf[b,q],-6,4))
// =============================================================================================
__global__ void conf_fusion_kernel(const float4* __restrict__ conf, const float4* __restrict__ value, int nq, size_t per_map4,
                                   size_t total4, float4* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    size_t b = idx / per_map4, e = idx % per_map4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), acc = s;
    for (int q = 0; q < nq; ++q) {
        size_t off = (b * nq + q) * per_map4 + e;
        float4 cf = __ldg(conf + off), v = __ldg(value + off);
        float ex = expf(fminf(fmaxf(cf.x, -6.f), 4.f)), ey = expf(fminf(fmaxf(cf.y, -6.f), 4.f));
        float ez = expf(fminf(fmaxf(cf.z, -6.f), 4.f)), ew = expf(fminf(fmaxf(cf.w, -6.f), 4.f));
        s.x += ex; s.y += ey; s.z += ez; s.w += ew;
        acc.x = fmaf(ex, v.x, acc.x); acc.y = fmaf(ey, v.y, acc.y); acc.z = fmaf(ez, v.z, acc.z); acc.w = fmaf(ew, v.w, acc.w);
    }
    out[idx] = make_float4(acc.x / s.x, acc.y / s.y, acc.z / s.z, acc.w / s.w);
}

int launch_conf_fusion(const float* conf, const float* value, int b, int nq, size_t per_map, float* out, cudaStream_t st) {
    USOT_REQUIRE(per_map % 4 == 0, "conf fusion: map size must be a multiple of 4");
    size_t total4 = (size_t)b * per_map / 4;
    if (total4 == 0) return 0;
    conf_fusion_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(conf), reinterpret_cast<const float4*>(value), nq, per_map / 4, total4,
        reinterpret_cast<float4*>(out));
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// PrRoIPool forward.  Same cell loop as the reference kernel; `cs`/`ps` are the channel / pixel strides of the
// feature tensor so that one body serves NHWC (engine) and NCHW (stand-alone op).
// =============================================================================================
static __device__ __forceinline__ float prroi_get(const float* d, int h, int w, int H, int W, size_t ps) {
    return (h < 0 || w < 0 || h >= H || w >= W) ? 0.f : __ldg(d + ((size_t)h * W + w) * ps);
}
static __device__ __forceinline__ float prroi_g(float lim, float a) { return lim - 0.5f * lim * lim - a + 0.5f * a * a; }

static __device__ float prroi_bin(const float* d, int H, int W, size_t ps, float x1, float y1, float x2, float y2, int ph, int pw,
                                  int PH, int PW, float scale) {
    const float sw = x1 * scale, sh = y1 * scale, ew = x2 * scale, eh = y2 * scale;
    const float roi_w = fmaxf(ew - sw, 0.f), roi_h = fmaxf(eh - sh, 0.f);
    const float bin_h = roi_h / (float)PH, bin_w = roi_w / (float)PW;
    const float ws_w = sw + bin_w * pw, ws_h = sh + bin_h * ph;
    const float we_w = ws_w + bin_w, we_h = ws_h + bin_h;
    const float win = fmaxf(0.f, bin_w * bin_h);
    if (win == 0.f) return 0.f;
    float sum = 0.f;
    const int s_w = (int)floorf(ws_w), e_w = (int)ceilf(we_w), s_h = (int)floorf(ws_h), e_h = (int)ceilf(we_h);
    for (int wi = s_w; wi < e_w; ++wi)
        for (int hi = s_h; hi < e_h; ++hi) {
            const float y0 = fmaxf(ws_h, (float)hi), x0 = fmaxf(ws_w, (float)wi);
            const float yy1 = fminf(we_h, (float)hi + 1.f), xx1 = fminf(we_w, (float)wi + 1.f);
            float al = x0 - (float)wi, be = y0 - (float)hi, la = xx1 - (float)wi, lb = yy1 - (float)hi;
            float cell = prroi_get(d, hi, wi, H, W, ps) * (prroi_g(la, al) * prroi_g(lb, be));
            float al2 = (float)(wi + 1) - xx1, la2 = (float)(wi + 1) - x0;
            cell += prroi_get(d, hi, wi + 1, H, W, ps) * (prroi_g(la2, al2) * prroi_g(lb, be));
            float be2 = (float)(hi + 1) - yy1, lb2 = (float)(hi + 1) - y0;
            cell += prroi_get(d, hi + 1, wi, H, W, ps) * (prroi_g(la, al) * prroi_g(lb2, be2));
            cell += prroi_get(d, hi + 1, wi + 1, H, W, ps) * (prroi_g(la2, al2) * prroi_g(lb2, be2));
            sum += cell;
        }
    return sum / win;
}

// NHWC features, boxes (n,4), implicit batch index = roi index (models.py:164-171), output NHWC (n,7,7,C)
__global__ void prroi_nhwc_kernel(const float* __restrict__ feat, int H, int W, int C, const float* __restrict__ boxes, int n_rois,
                                  float* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n_rois * 49 * C;
    if (idx >= total) return;
    int c = idx % C;
    int bin = (idx / C) % 49;
    int n = idx / ((size_t)C * 49);
    const float* b = boxes + (size_t)n * 4;
    out[idx] = prroi_bin(feat + (size_t)n * H * W * C + c, H, W, C, __ldg(b), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3), bin / 7,
                         bin % 7, 7, 7, 1.0f);
}

int launch_prroi_nhwc(const float* feat, int n_feat, int h, int w, int c, const float* boxes4, int n_rois, float* out,
                      cudaStream_t st) {
    USOT_REQUIRE(n_rois <= n_feat, "prroi: more rois than feature maps");
    size_t total = (size_t)n_rois * 49 * c;
    if (total == 0) return 0;
    prroi_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, h, w, c, boxes4, n_rois, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// reference layout: NCHW features, rois (n,5) = [batch_idx, x1, y1, x2, y2], output NCHW
__global__ void prroi_nchw_kernel(const float* __restrict__ feat, int C, int H, int W, const float* __restrict__ rois, int n_rois,
                                  int PH, int PW, float scale, float* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n_rois * C * PH * PW;
    if (idx >= total) return;
    int pw = idx % PW, ph = (idx / PW) % PH;
    int c = (idx / ((size_t)PW * PH)) % C;
    int n = idx / ((size_t)PW * PH * C);
    const float* r = rois + (size_t)n * 5;
    int bi = (int)__ldg(r);
    out[idx] = prroi_bin(feat + ((size_t)bi * C + c) * H * W, H, W, 1, __ldg(r + 1), __ldg(r + 2), __ldg(r + 3), __ldg(r + 4), ph,
                         pw, PH, PW, scale);
}

int launch_prroi_nchw(const float* feat, int c, int h, int w, const float* rois5, int n_rois, int ph, int pw, float scale,
                      float* out, cudaStream_t st) {
    size_t total = (size_t)n_rois * c * ph * pw;
    if (total == 0) return 0;
    prroi_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, c, h, w, rois5, n_rois, ph, pw, scale, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// Layout helpers (boundary only)
// =============================================================================================
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int HW, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const float* ip = in + (size_t)n * C * HW;
    float* op = out + (size_t)n * C * HW;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? ip[(size_t)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) op[(size_t)p * C + c] = tile[threadIdx.x][i];
    }
}

int launch_nchw_to_nhwc(const float* in, int n, int c, int h, int w, float* out, cudaStream_t st) {
    if ((size_t)n * c * h * w == 0) return 0;
    dim3 grid((h * w + 31) / 32, (c + 31) / 32, n), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, st>>>(in, c, h * w, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// NHWC -> NCHW is the same transpose with the roles of (C, HW) swapped.
int launch_nhwc_to_nchw(const float* in, int n, int h, int w, int c, float* out, cudaStream_t st) {
    if ((size_t)n * c * h * w == 0) return 0;
    dim3 grid((c + 31) / 32, (h * w + 31) / 32, n), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, st>>>(in, h * w, c, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void center_crop_kernel(const float4* __restrict__ in, int h, int w, int c4, int l, int ho, int wo, size_t total,
                                   float4* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int cc = idx % c4;
    size_t t = idx / c4;
    int x = t % wo; t /= wo;
    int y = t % ho;
    int b = t / ho;
    out[idx] = __ldg(in + (((size_t)b * h + y + l) * w + x + l) * c4 + cc);
}

int launch_center_crop_nhwc(const float* in, int n, int h, int w, int c, int l, float* out, cudaStream_t st) {
    int ho = h - 2 * l, wo = w - 2 * l;
    USOT_REQUIRE(ho > 0 && wo > 0 && c % 4 == 0, "center crop: bad geometry");
    size_t total = (size_t)n * ho * wo * (c / 4);
    if (total == 0) return 0;
    center_crop_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in), h, w, c / 4, l, ho, wo,
                                                                         total, reinterpret_cast<float4*>(out));
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace usot
