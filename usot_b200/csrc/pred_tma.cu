// Skinny prediction convs of the head (bbox_pred 256->4, cls_pred / cls_memory_pred 256->1; 3x3, pad 1, bias):
//   lib/models/connect.py:235-241,274-275 (`0.1 * cls_pred(...)`, `exp(adjust * bbox_pred(...) + bias)`).
//
// Bandwidth-shaped work (164 MB of fp32 activations per batch-256 launch against 0.4-1.5 GFMA), so the input must be read from
// HBM exactly once and the FMA pipe must not become the limiter.  One CTA = one image:
//   * a producer warp streams the image's (R x R x 256) NHWC fp32 map in eight 32-channel slices through a 2-stage ring with
//     ONE 4-D TMA box {32 ch, R, R, 1} per slice (128B swizzle: pixel rows of 128 B, conflict-free 128-bit reads when
//     lane = pixel);
//   * consumer lanes own two PIXELS each and compute the per-pixel partial products P[pixel][tap*COUT+co] = <in[pixel,:],
//     w[tap,co,:]> for all 9 taps at once (a 625 x 36 x 256 GEMM per image) with packed `fma.rn.f32x2`: the (w_a, w_b) pairs
//     are broadcast 128-bit shared-memory loads shared by both pixels, the activation is duplicated into both halves;
//   * the 9-tap neighbourhood sum, bias and the exp / 0.1x epilogue run from shared memory; NCHW output stores are coalesced.
// Small batches (fewer images than a third of the SMs) run pred_dot_kernel instead, which spreads one image over many SMs: one
// thread per (output pixel, tap, co) dot product.  Both kernels build every partial product as the SAME sequential fma.rn chain
// over the 256 channels and add the nine taps in the same order, so they agree bit for bit and a sample's result does not
// depend on the batch it is computed in (tests/test_gpu_model.py::test_batch_256_independence, sharded == single device).
#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace usot {

constexpr int PD_STAGES = 2, PD_CH = 32;

struct PredParams {
    CUtensorMap xmap;
    const float* w;       // [9][COUT][C]
    const float* b;       // [COUT]
    const float* adjust;  // [1]
    const float* bias4;   // [4]
    float* out;           // [n][COUT][R][R]
    int R, C, mode, cwarps, stage_bytes;
    float mul;
};

static __device__ __forceinline__ void pd_ffma2(float2& d, const float2& a, const float2& b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}
static __device__ __forceinline__ void pd_bar_consumers(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <int COUT>
__global__ void __launch_bounds__(13 * 32, 1) pred_gemm_kernel(const __grid_constant__ PredParams p) {
    constexpr int NP = (9 * COUT + 1) / 2 * 2;  // partial products per pixel, padded to an even count: 36 / 10
    constexpr int NPAIR = NP / 2;
    extern __shared__ __align__(1024) uint8_t psm_raw[];
    const uint32_t base = (smem_u32(psm_raw) + 1023u) & ~1023u;
    uint8_t* sm = psm_raw + (base - smem_u32(psm_raw));
    const int R = p.R, C = p.C, npix = R * R, nchunk = C / PD_CH;
    const int cthreads = p.cwarps * 32;
    float* ws = reinterpret_cast<float*>(sm + PD_STAGES * p.stage_bytes);  // [C][NP]: weights of channel c, (tap, co) fastest
    const uint32_t bar_full = base + PD_STAGES * p.stage_bytes + C * NP * 4, bar_empty = bar_full + 8 * PD_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int img = blockIdx.x;

    if (tid == 0) {
        for (int s = 0; s < PD_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, p.cwarps); }
        fence_barrier_init();
    }
    for (int i = tid; i < NP * C; i += blockDim.x) {  // coalesced read of [k][c], transposed store
        const int k = i / C, c = i - k * C;
        ws[c * NP + k] = (k < 9 * COUT) ? __ldg(p.w + i) : 0.f;
    }
    __syncthreads();

    if (warp == p.cwarps) {
        // ------------------------------- producer -------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&p.xmap);
            for (int k = 0; k < nchunk; ++k) {
                const int s = k % PD_STAGES;
                if (k >= PD_STAGES) mbar_wait(bar_empty + 8 * s, ((k / PD_STAGES) + 1) & 1);
                mbar_expect_tx(bar_full + 8 * s, (uint32_t)(npix * PD_CH * 4));
                tma_load_4d(base + s * p.stage_bytes, &p.xmap, bar_full + 8 * s, k * PD_CH, 0, 0, img);
            }
        }
        return;
    }

    // ------------------------------- consumers: lane = two pixels -------------------------------
    const int q0 = tid, q1 = tid + cthreads;
    const int r0 = min(q0, npix - 1), r1 = min(q1, npix - 1);  // clamped rows for the loads of padding lanes
    const uint32_t off0 = r0 * 128, sw0 = r0 & 7, off1 = r1 * 128, sw1 = r1 & 7;
    float2 acc0[NPAIR], acc1[NPAIR];
#pragma unroll
    for (int k = 0; k < NPAIR; ++k) acc0[k] = acc1[k] = make_float2(0.f, 0.f);

    for (int kc = 0; kc < nchunk; ++kc) {
        const int s = kc % PD_STAGES;
        mbar_wait(bar_full + 8 * s, (kc / PD_STAGES) & 1);
        const uint8_t* st = sm + s * p.stage_bytes;
        const float* wk = ws + kc * PD_CH * NP;
#pragma unroll
        for (int j = 0; j < PD_CH / 4; ++j) {
            const float4 xa = *reinterpret_cast<const float4*>(st + off0 + ((j ^ sw0) << 4));
            const float4 xb = *reinterpret_cast<const float4*>(st + off1 + ((j ^ sw1) << 4));
            const float xav[4] = {xa.x, xa.y, xa.z, xa.w}, xbv[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float2 a2 = make_float2(xav[cc], xav[cc]), b2 = make_float2(xbv[cc], xbv[cc]);
                const float* wrow = wk + (j * 4 + cc) * NP;  // warp-uniform address: broadcast loads
                if constexpr (NP % 4 == 0) {  // rows of 144 B: 128-bit loads
#pragma unroll
                    for (int k = 0; k < NP / 4; ++k) {
                        const float4 w4 = reinterpret_cast<const float4*>(wrow)[k];
                        const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
                        pd_ffma2(acc0[2 * k], a2, wa);
                        pd_ffma2(acc1[2 * k], b2, wa);
                        pd_ffma2(acc0[2 * k + 1], a2, wb);
                        pd_ffma2(acc1[2 * k + 1], b2, wb);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NPAIR; ++k) {
                        const float2 w2 = reinterpret_cast<const float2*>(wrow)[k];
                        pd_ffma2(acc0[k], a2, w2);
                        pd_ffma2(acc1[k], b2, w2);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    }

    // ---- partial products -> shared memory S[k][pixel slot] (aliases the operand ring, which every consumer has finished reading)
    pd_bar_consumers(cthreads);
    float* S = reinterpret_cast<float*>(sm);
    const int SP = 2 * cthreads;
#pragma unroll
    for (int k = 0; k < NPAIR; ++k) {
        S[(2 * k) * SP + q0] = acc0[k].x;
        S[(2 * k + 1) * SP + q0] = acc0[k].y;
        S[(2 * k) * SP + q1] = acc1[k].x;
        S[(2 * k + 1) * SP + q1] = acc1[k].y;
    }
    pd_bar_consumers(cthreads);

    // ---- 3x3 neighbourhood sum (zero padding), bias, epilogue; out is NCHW
    const float adj = (p.mode == 1) ? __ldg(p.adjust) : 0.f;
    for (int idx = tid; idx < COUT * npix; idx += cthreads) {
        const int co = idx / npix, pix = idx - co * npix;
        const int y = pix / R, x = pix - y * R;
        float v = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int yy = y + dy - 1, xx = x + dx - 1;
                if (yy >= 0 && yy < R && xx >= 0 && xx < R) v += S[((dy * 3 + dx) * COUT + co) * SP + yy * R + xx];
            }
        v += __ldg(p.b + co);
        v = (p.mode == 0) ? p.mul * v : expf(fmaf(adj, v, __ldg(p.bias4 + co)));
        p.out[((size_t)img * COUT + co) * npix + pix] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// small-batch variant: thread = (output pixel, k = tap*COUT+co); w4 is the [C/4][9*COUT][4] repack of the weights so that the
// lanes of a warp (consecutive k) read consecutive 16-byte groups; the activation row of the tap's neighbour pixel is a
// broadcast read.
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(144) pred_dot_kernel(const float* __restrict__ in, int total_pix, int R, int C, const float* __restrict__ w4,
                                                       const float* __restrict__ b, int mode, float mul, const float* __restrict__ adjust,
                                                       const float* __restrict__ bias4, float* __restrict__ out) {
    constexpr int K = 9 * COUT, PXB = 144 / K;
    __shared__ float P[PXB][K];
    __shared__ int valid[PXB][9];
    const int t = threadIdx.x, px = t / K, k = t - px * K;
    const int tap = k / COUT;
    const int npix = R * R;
    const int o = blockIdx.x * PXB + px;  // flat (image, y, x)
    bool ok = o < total_pix;
    int img = 0, y = 0, x = 0;
    if (ok) {
        img = o / npix;
        const int pix = o - img * npix;
        y = pix / R; x = pix - y * R;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        ok = yy >= 0 && yy < R && xx >= 0 && xx < R;
        if (ok) {
            const float4* xr = reinterpret_cast<const float4*>(in + (((size_t)img * R + yy) * R + xx) * C);
            const float4* wr = reinterpret_cast<const float4*>(w4) + k;
            float acc = 0.f;
#pragma unroll 8
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 xv = __ldg(xr + c4), wv = __ldg(wr + (size_t)c4 * K);
                acc = fmaf(xv.x, wv.x, acc);
                acc = fmaf(xv.y, wv.y, acc);
                acc = fmaf(xv.z, wv.z, acc);
                acc = fmaf(xv.w, wv.w, acc);
            }
            P[px][k] = acc;
        }
    }
    if (k % COUT == 0) valid[px][tap] = ok ? 1 : 0;
    __syncthreads();
    if (t < PXB * COUT) {
        const int p2 = t / COUT, co = t - p2 * COUT;
        const int o2 = blockIdx.x * PXB + p2;
        if (o2 < total_pix) {
            float v = 0.f;
#pragma unroll
            for (int tp = 0; tp < 9; ++tp)
                if (valid[p2][tp]) v += P[p2][tp * COUT + co];
            v += __ldg(b + co);
            v = (mode == 0) ? mul * v : expf(fmaf(__ldg(adjust), v, __ldg(bias4 + co)));
            const int img2 = o2 / npix, pix2 = o2 - img2 * npix;
            out[((size_t)img2 * COUT + co) * npix + pix2] = v;
        }
    }
}

__global__ void pred_repack_kernel(const float* __restrict__ w, int K, int C, float* __restrict__ w4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over [k][c]
    if (i >= K * C) return;
    const int k = i / C, c = i - k * C;
    w4[((size_t)(c / 4) * K + k) * 4 + (c & 3)] = w[i];
}

int launch_pred_repack(const float* w, int cout, int C, float* w4, cudaStream_t st) {
    const int total = 9 * cout * C;
    pred_repack_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, 9 * cout, C, w4);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

Tunable g_pred_tma_min_batch = 48;  // tunable "pred_tma_min_batch": batches of at least this many images use the kernel above (0 = never)

bool pred_tma_supported(int n, int r, int C, int cout) {
    return g_pred_tma_min_batch > 0 && n >= g_pred_tma_min_batch && C == 256 && (cout == 1 || cout == 4) && r >= 8 && r <= 27;
}

int launch_pred_tma(const float* in, int n, int r, int C, const float* w, const float* b, int cout, int mode, float mul,
                    const float* adjust, const float* bias4, float* out, cudaStream_t st) {
    USOT_REQUIRE(pred_tma_supported(n, r, C, cout), "pred_tma: unsupported shape");
    PredParams p;
    memset(&p, 0, sizeof(p));
    const int npix = r * r;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)r * C * 4, (cuuint64_t)npix * C * 4};
    cuuint32_t box[4] = {PD_CH, (cuuint32_t)r, (cuuint32_t)r, 1};
    if (int rc = encode_tmap(&p.xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    p.w = w; p.b = b; p.adjust = adjust; p.bias4 = bias4; p.out = out;
    p.R = r; p.C = C; p.mode = mode; p.mul = mul;
    p.cwarps = (npix + 63) / 64;
    p.stage_bytes = (npix * PD_CH * 4 + 1023) & ~1023;
    const int np = (9 * cout + 1) / 2 * 2;
    USOT_REQUIRE(np * 2 * p.cwarps * 32 * 4 <= PD_STAGES * p.stage_bytes, "pred_tma: partial-product buffer does not fit the ring");
    const int smem = PD_STAGES * p.stage_bytes + C * np * 4 + 2 * PD_STAGES * 8 + 1024;
    USOT_REQUIRE(smem <= 227 * 1024, "pred_tma: shared memory budget exceeded");
    const int threads = (p.cwarps + 1) * 32;
    if (cout == 4) {
        static SmemAttrCache cache;
        if (int rc = cache.ensure(pred_gemm_kernel<4>, smem)) return rc;
        pred_gemm_kernel<4><<<n, threads, smem, st>>>(p);
    } else {
        static SmemAttrCache cache;
        if (int rc = cache.ensure(pred_gemm_kernel<1>, smem)) return rc;
        pred_gemm_kernel<1><<<n, threads, smem, st>>>(p);
    }
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// Dispatcher.  w = [9][cout][C] (tap-major, the layout pred_gemm_kernel transposes itself); w4 = its [C/4][9*cout][4] repack for the
// small-batch kernel, or nullptr (stand-alone op): then the repack runs into a stream-ordered temporary.
int launch_pred_conv(const float* in, int n, int r, int C, const float* w, const float* w4, const float* b, int cout, int mode, float mul,
                     const float* adjust, const float* bias4, float* out, cudaStream_t st) {
    USOT_REQUIRE(C % 4 == 0 && (cout == 1 || cout == 4), "pred conv needs C % 4 == 0 and Cout in {1, 4}");
    const int total = n * r * r;
    if (total == 0) return 0;
    if (pred_tma_supported(n, r, C, cout)) return launch_pred_tma(in, n, r, C, w, b, cout, mode, mul, adjust, bias4, out, st);
    float* tmp = nullptr;
    if (!w4) {
        USOT_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&tmp), (size_t)9 * cout * C * sizeof(float), st));
        if (int rc = launch_pred_repack(w, cout, C, tmp, st)) return rc;
        w4 = tmp;
    }
    const int pxb = 144 / (9 * cout);
    const unsigned grid = (unsigned)((total + pxb - 1) / pxb);
    if (cout == 1) pred_dot_kernel<1><<<grid, 144, 0, st>>>(in, total, r, C, w4, b, mode, mul, adjust, bias4, out);
    else pred_dot_kernel<4><<<grid, 144, 0, st>>>(in, total, r, C, w4, b, mode, mul, adjust, bias4, out);
    USOT_CUDA_OK(cudaGetLastError());
    if (tmp) USOT_CUDA_OK(cudaFreeAsync(tmp, st));
    return 0;
}

}  // namespace usot
