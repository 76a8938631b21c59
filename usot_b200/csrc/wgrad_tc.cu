// Conv weight gradient on the 5th-generation tensor cores (training path; the reference gets it from torch autograd
// over cuDNN under loss.backward(), scripts/train_usot.py:229-233).
//
//     dW[(tap*cin + ci)][co] = sum over output pixels p of  X[p shifted by the tap][ci] * dY[p][co]
//
// One GEMM per filter tap with  M = ci (128 rows), N = co (64 / 128 columns), K = output pixels.  Both operands are PIXEL-major in
// memory (NHWC: the channel index is contiguous, the reduction index is not), i.e. "MN-major" in tcgen05 terms: a 4-D TMA box
// {64 channels, bw, bh, 1} (bw*bh = 64 pixels) lands in shared memory as 64 rows of 128 bytes under the 128-byte swizzle, which is
// exactly the canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 64 channels contiguous, pixel rows
// 128 B apart, 8-pixel groups 1024 B apart (SBO), 64-channel blocks one box apart (LBO) -- so no transpose pass is needed: the
// instruction descriptor's a_major / b_major bits select MN-major for both operands.  The X box is the forward kernel's activation
// box (tap shift / dilation as a coordinate offset, padding as out-of-bounds zero fill, stride 2 as four parity-decimated maps); the
// dY box uses the un-shifted output coordinates, and pixels of a patch that lie outside the output map are zero-filled too.
//
// Precision: fp16x3 like the forward pass (operands split hi + lo, hi*hi into `main`, hi*lo + lo*hi into `cross`, summed in fp32 in
// the epilogue).  dY is multiplied by a per-tensor power of two first (gradients sit far below fp16's normal range); the factor is
// computed on the device (absmax reduction) and divided out in the epilogue.  Split-K over pixel patches, fp32 atomics into dW.
#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstring>

namespace usot {

constexpr int WG_BM = 128, WG_KP = 64;                 // ci rows per tile, pixels per pipeline stage
constexpr int WG_BOX_BYTES = WG_KP * 128;              // one {64 ch, 64 px} fp16 box
constexpr int WG_THREADS = 256;

struct WgParams {
    CUtensorMap a[2][4];   // X planes [hi/lo][stride-2 parity]
    CUtensorMap b[2];      // dY planes [hi/lo]
    float* dw;
    const float* inv_scale;  // device scalar: 1 / (power-of-two factor applied to dY)
    int cin, cout, taps, kw, stride, ph, pw, dh, dw_dil;
    int n_img, tiles_h, tiles_w, bh, bw;
    int ci_tiles, co_tiles, total_patches, patches_per_split;
    int stages, split;     // split: fp16x3 (1) or single fp16 (0)
};

// MN-major operand, 128-byte swizzle: 64-channel blocks `lbo` bytes apart, 8-pixel groups 1024 B apart.
static __device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;   // bits [16,30) leading byte offset >> 4
    d |= (uint64_t)(1024 >> 4) << 32;             // bits [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// kind::f16, D = F32, A = B = F16, BOTH MN-major (bits 15 / 16)
static __device__ __forceinline__ uint32_t make_idesc_mn(int m, int n) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
    constexpr int PL_MAX = 2;
    constexpr int A_BYTES = PL_MAX * 2 * WG_BOX_BYTES;            // planes x two 64-ci blocks
    constexpr int B_BYTES = PL_MAX * (BN / 64) * WG_BOX_BYTES;    // planes x BN/64 blocks (hi blocks first, then lo blocks)
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = 2 * BN;                             // main | cross
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + p.stages * STAGE_BYTES);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * 4, bar_done = bar_empty + 8 * 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bt = blockIdx.x;
    const int cot = bt % p.co_tiles; bt /= p.co_tiles;
    const int cit = bt % p.ci_tiles;
    const int tap = bt / p.ci_tiles;
    const int ci0 = cit * WG_BM, co0 = cot * BN;
    const int a_blocks = (p.cin - ci0) >= 128 ? 2 : 1;           // a 64-channel layer fills only the first block (rows 64.. are never stored)
    const int p_begin = blockIdx.y * p.patches_per_split, p_end = min(p.total_patches, p_begin + p.patches_per_split);
    const int planes = p.split ? 2 : 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a[0][0]); tma_prefetch_desc(&p.b[0]);
        if (p.split) { tma_prefetch_desc(&p.a[1][0]); tma_prefetch_desc(&p.b[1]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (p_begin < p_end) {   // (all 32 lanes walk the loop; the single-thread instructions are predicated on elect_one(), see tc_ptx.cuh)
            const int kh = tap / p.kw, kwi = tap - kh * p.kw;
            int offh = kh * p.dh - p.ph, offw = kwi * p.dw_dil - p.pw, par = 0;
            if (p.stride == 2) {
                const int py = offh & 1, px = offw & 1;
                par = py * 2 + px;
                offh = (offh - py) >> 1;
                offw = (offw - px) >> 1;
            }
            const uint32_t tx = (uint32_t)planes * (a_blocks + BN / 64) * WG_BOX_BYTES;
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = p_begin; pt < p_end; ++pt) {
                int t = pt;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h;
                const int img = t / p.tiles_h;
                const int h0 = th * p.bh, w0 = tw * p.bw;
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t full = bar_full + 8 * stage;
                const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(full, tx);
                    for (int pl = 0; pl < planes; ++pl) {
                        for (int blk = 0; blk < a_blocks; ++blk)
                            tma_load_4d(sa + (pl * 2 + blk) * WG_BOX_BYTES, &p.a[pl][par], full, ci0 + blk * 64, w0 + offw, h0 + offh, img);
                        for (int blk = 0; blk < BN / 64; ++blk)
                            tma_load_4d(sb + (pl * (BN / 64) + blk) * WG_BOX_BYTES, &p.b[pl], full, co0 + blk * 64, w0, h0, img);
                    }
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (p_begin < p_end) {   // (all lanes wait on the barriers; one elected lane issues the MMAs and commits of a stage)
            const uint32_t idesc = make_idesc_mn(WG_BM, BN), idesc2 = make_idesc_mn(WG_BM, 2 * BN);
            const uint32_t tmem_d = tmem_base, tmem_x = tmem_base + BN;
            int stage = 0;
            uint32_t phase = 0;
            bool first = true;
            for (int pt = p_begin; pt < p_end; ++pt) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < WG_KP / 16; ++k) {   // 16 pixels per MMA = two 8-pixel groups = 2048 B
                        const uint64_t a_hi = make_smem_desc_mn(sa + k * 2048, WG_BOX_BYTES), b_hi = make_smem_desc_mn(sb + k * 2048, WG_BOX_BYTES);
                        const uint32_t acc = (first && k == 0) ? 0u : 1u;
                        if (p.split) {
                            // [b_hi blocks | b_lo blocks] are contiguous with a uniform block stride: ONE N = 2*BN MMA gives x_hi*dy_hi -> main, x_hi*dy_lo -> cross
                            const uint64_t a_lo = make_smem_desc_mn(sa + 2 * WG_BOX_BYTES + k * 2048, WG_BOX_BYTES);
                            umma_f16(tmem_d, a_hi, b_hi, idesc2, acc);
                            umma_f16(tmem_x, a_lo, b_hi, idesc, 1u);
                        } else {
                            umma_f16(tmem_d, a_hi, b_hi, idesc, acc);
                        }
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (pt == p_end - 1) umma_commit(bar_done);
                }
                __syncwarp();
                first = false;
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue: TMEM -> scale -> atomic add into dW ================================
        if (p_begin < p_end) {
            const int ew = warp & 3, row = ew * 32 + lane;
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
            const bool row_ok = ci0 + row < p.cin;
            float* dst = p.dw + ((size_t)tap * p.cin + ci0 + row) * p.cout + co0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                if (p.split) {
                    uint32_t x[32];
                    tmem_ld32(taddr + BN + c0, x);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(x[j]));
                } else {
                    tmem_ld_wait();
                }
                if (row_ok) {   // 128-bit vector reductions (red.global.add.v4.f32): 8 per 32-column chunk instead of 32 scalar atomics
                    float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        atomicAdd(d4 + j, make_float4(__uint_as_float(v[4 * j]) * inv, __uint_as_float(v[4 * j + 1]) * inv,
                                                      __uint_as_float(v[4 * j + 2]) * inv, __uint_as_float(v[4 * j + 3]) * inv));
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---- |x|_max -> power-of-two factor s with s * |x|_max in [2^target, 2^(target+1)) ; out2 = {s, 1/s} -------------------------------
__global__ void absmax_kernel(const float4* __restrict__ x, size_t n4, unsigned* __restrict__ bits) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(bits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}
__global__ void pow2_scale_kernel(const unsigned* __restrict__ bits, int target_log2, float* __restrict__ out2) {
    const float m = __uint_as_float(*bits);
    float s = 1.f;
    if (m > 0.f && isfinite(m)) {
        int ex;
        frexpf(m, &ex);                   // m = f * 2^ex, f in [0.5, 1)
        s = ldexpf(1.f, target_log2 + 1 - ex);
    }
    out2[0] = s;
    out2[1] = 1.f / s;
}
__global__ void scale_kernel(const float4* __restrict__ x, size_t n4, const float* __restrict__ s, float4* __restrict__ y) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float f = __ldg(s);
    const float4 v = __ldg(x + i);
    y[i] = make_float4(v.x * f, v.y * f, v.z * f, v.w * f);
}

int launch_pow2_scale(const float* x, size_t n, int target_log2, float* y /*or null*/, float* out2, cudaStream_t st) {
    USOT_REQUIRE(n % 4 == 0 && n > 0, "pow2_scale: element count must be a positive multiple of 4");
    unsigned* bits = nullptr;
    ensure_async_pool();
    USOT_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&bits), sizeof(unsigned), st));
    USOT_CUDA_OK(cudaMemsetAsync(bits, 0, sizeof(unsigned), st));
    const size_t n4 = n / 4;
    const unsigned blocks = (unsigned)std::min<size_t>((n4 + 255) / 256, (size_t)device_sm_count() * 8);
    absmax_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(x), n4, bits);
    pow2_scale_kernel<<<1, 1, 0, st>>>(bits, target_log2, out2);
    if (y) scale_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n4, out2, reinterpret_cast<float4*>(y));
    USOT_CUDA_OK(cudaGetLastError());
    USOT_CUDA_OK(cudaFreeAsync(bits, st));
    return 0;
}

static void choose_patch(int ho, int wo, int* bw, int* bh) {
    double best = -1;
    for (int w = 64; w >= 4; w /= 2) {
        const int h = WG_KP / w;
        const double eff = (double)ho * wo / ((double)((wo + w - 1) / w) * ((ho + h - 1) / h) * WG_KP);
        if (eff > best + 1e-9) { best = eff; *bw = w; *bh = h; }
    }
}

bool wgrad_tc_supported(const ConvGeom& g) {
    return g.cin % 64 == 0 && g.cout % 64 == 0 && (g.stride == 1 || g.stride == 2) && (size_t)g.n * g.ho * g.wo >= 64;
}

int launch_conv_wgrad_tc(const float* x, const float* dy, const ConvGeom& g, float* dw_kn, bool split, cudaStream_t st) {
    USOT_REQUIRE(wgrad_tc_supported(g), "wgrad_tc needs Cin % 64 == 0, Cout % 64 == 0, stride 1 or 2");
    const size_t n_x = (size_t)g.n * g.h * g.w * g.cin, n_y = (size_t)g.n * g.ho * g.wo * g.cout, n_w = (size_t)g.kh * g.kw * g.cin * g.cout;
    USOT_CUDA_OK(cudaMemsetAsync(dw_kn, 0, n_w * sizeof(float), st));
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t planes = split ? 2 : 1;
    const size_t bytes = planes * (al(n_x * 2) + al(n_y * 2)) + 256;
    char* ws = nullptr;
    ensure_async_pool();
    USOT_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), bytes, st));
    char* cur = ws;
    auto take = [&](size_t b) { char* q = cur; cur += al(b); return q; };
    __half* x_hi = reinterpret_cast<__half*>(take(n_x * 2));
    __half* x_lo = split ? reinterpret_cast<__half*>(take(n_x * 2)) : nullptr;
    __half* y_hi = reinterpret_cast<__half*>(take(n_y * 2));
    __half* y_lo = split ? reinterpret_cast<__half*>(take(n_y * 2)) : nullptr;
    float* sc2 = reinterpret_cast<float*>(take(8));
    int rc = 0;
    do {
        if ((rc = launch_pow2_scale(dy, n_y, 10, nullptr, sc2, st))) break;      // s with s*|dY|_max in [1024, 2048): far inside fp16's normal range
        if ((rc = launch_f32_to_split(x, n_x, x_hi, x_lo, st))) break;
        if ((rc = launch_f32_to_split(dy, n_y, y_hi, y_lo, st, sc2))) break;     // the multiplication by s rides on the conversion (exact: power of two)
        WgParams p;
        memset(&p, 0, sizeof(p));
        choose_patch(g.ho, g.wo, &p.bw, &p.bh);
        p.tiles_w = (g.wo + p.bw - 1) / p.bw; p.tiles_h = (g.ho + p.bh - 1) / p.bh;
        p.n_img = g.n; p.total_patches = g.n * p.tiles_h * p.tiles_w;
        p.cin = g.cin; p.cout = g.cout; p.taps = g.kh * g.kw; p.kw = g.kw; p.stride = g.stride; p.ph = g.ph; p.pw = g.pw; p.dh = g.dh; p.dw_dil = g.dw;
        p.dw = dw_kn; p.inv_scale = sc2 + 1; p.split = split ? 1 : 0;
        const int bn = g.cout % 128 == 0 ? 128 : 64;
        p.ci_tiles = (g.cin + WG_BM - 1) / WG_BM; p.co_tiles = g.cout / bn;
        const int tiles = p.taps * p.ci_tiles * p.co_tiles;
        // split-K: enough CTAs for ~3 waves, at least 8 patches per CTA -- and at most 48 (192 accumulating MMAs per TMEM tile): the tensor core
        // truncates on every accumulate, so the error of one accumulator grows with the number of MMAs added into it; the
        // partial sums of the splits meet in fp32 round-to-nearest atomics.
        int splits = std::max(1, std::min((device_sm_count() * 3 + tiles - 1) / tiles, (p.total_patches + 7) / 8));
        splits = std::max(splits, (p.total_patches + 47) / 48);
        p.patches_per_split = (p.total_patches + splits - 1) / splits;
        splits = (p.total_patches + p.patches_per_split - 1) / p.patches_per_split;
        const int stage_bytes = 2 * 2 * WG_BOX_BYTES + 2 * (bn / 64) * WG_BOX_BYTES;
        p.stages = std::min(4, (227 * 1024 - 2048) / stage_bytes);
        // ---- X maps: dims {C, W', H', N} per (plane, parity), box {64, bw, bh, 1} ----
        const int npar = g.stride == 2 ? 4 : 1;
        for (int pl = 0; pl < (int)planes && !rc; ++pl) {
            const __half* base = pl == 0 ? x_hi : x_lo;
            for (int par = 0; par < npar && !rc; ++par) {
                const int py = par >> 1, px = par & 1;
                cuuint64_t dims[4], strides[3];
                cuuint32_t box[4] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1};
                const __half* b = base;
                if (g.stride == 1) {
                    dims[0] = g.cin; dims[1] = g.w; dims[2] = g.h; dims[3] = g.n;
                    strides[0] = (cuuint64_t)g.cin * 2; strides[1] = (cuuint64_t)g.w * g.cin * 2; strides[2] = (cuuint64_t)g.h * g.w * g.cin * 2;
                } else {
                    if (py >= g.h || px >= g.w) continue;
                    b = base + ((size_t)py * g.w + px) * g.cin;
                    dims[0] = g.cin; dims[1] = (g.w - px + 1) / 2; dims[2] = (g.h - py + 1) / 2; dims[3] = g.n;
                    strides[0] = (cuuint64_t)g.cin * 4; strides[1] = (cuuint64_t)g.w * g.cin * 4; strides[2] = (cuuint64_t)g.h * g.w * g.cin * 2;
                }
                rc = encode_tmap(&p.a[pl][par], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, b, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            }
            if (rc) break;
            cuuint64_t yd[4] = {(cuuint64_t)g.cout, (cuuint64_t)g.wo, (cuuint64_t)g.ho, (cuuint64_t)g.n};
            cuuint64_t ys[3] = {(cuuint64_t)g.cout * 2, (cuuint64_t)g.wo * g.cout * 2, (cuuint64_t)g.ho * g.wo * g.cout * 2};
            cuuint32_t yb[4] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1};
            rc = encode_tmap(&p.b[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, pl == 0 ? y_hi : y_lo, 4, yd, ys, yb, CU_TENSOR_MAP_SWIZZLE_128B);
        }
        if (rc) break;
        if (!split) { for (int par = 0; par < 4; ++par) p.a[1][par] = p.a[0][par]; p.b[1] = p.b[0]; }
        const int smem = p.stages * stage_bytes + 1024 + 256;
        dim3 grid((unsigned)tiles, (unsigned)splits);
        if (bn == 128) {
            static SmemAttrCache attr;
            if ((rc = attr.ensure(wgrad_tc_kernel<128>, 227 * 1024))) break;
            wgrad_tc_kernel<128><<<grid, WG_THREADS, smem, st>>>(p);
        } else {
            static SmemAttrCache attr;
            if ((rc = attr.ensure(wgrad_tc_kernel<64>, 227 * 1024))) break;
            wgrad_tc_kernel<64><<<grid, WG_THREADS, smem, st>>>(p);
        }
        if (cudaGetLastError() != cudaSuccess) { set_error("wgrad_tc: launch failed"); rc = 1; }
    } while (0);
    cudaFreeAsync(ws, st);
    return rc;
}

}  // namespace usot
