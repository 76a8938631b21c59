// Stem on the tensor cores: conv1 7x7 / stride 2 / pad 0 (Cin = 3) + folded BN + ReLU   lib/models/modules.py:70-74,138-140
//
// As a GEMM the stem is M = output pixels, N = 64, K = 3*7*7 = 147 (laid out as 21 x 8 = 168, padded to 176 = eleven k16 steps).  Cin = 3 is far too thin
// for a TMA-fed implicit GEMM, so the A operand is built in shared memory by the CTA itself ("software im2col"):
//   1. the 7 input rows x 3 channels a tile of one output row needs are staged as fp32 with cp.async (double buffered),
//   2. all 256 threads convert them into the split-fp16 (hi, lo) A tile, written directly in the 128-byte-swizzled K-major
//      layout tcgen05.mma reads; K is ordered k' = (c*7+kh)*8 + kw (kw padded to 8) so that one task = 7 contiguous patch floats
//      -> one 16-byte store,
//   3. one thread issues the 10 x 3 MMAs (hi*hi + hi*lo + lo*hi) against the weight tile that stays resident in shared memory,
//      accumulating in one of two TMEM buffers (main + cross-term accumulator each, as in conv_tc.cu),
//   4. while those MMAs run, all 8 warps drain the PREVIOUS tile's accumulator: scale/shift (BN) + ReLU -> NHWC fp32.
// Input is the reference's NCHW fp32 image in the raw 0..255 range (lib/utils/track_utils.py:24-27).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cmath>
#include <cstring>
#include <vector>

namespace usot {

constexpr int ST_KUSED = 176, ST_PW = 261, ST_PWP = 264, ST_PROWS = 21;  // K' = (c*7+kh)*8 + kw (kw padded 7 -> 8): 168 -> 176 = 11 k16 steps
constexpr int ST_A_PLANE = 3 * 128 * 128;   // 3 chunks x 128 rows x 128 B
constexpr int ST_B_PLANE = 3 * 64 * 128;    // 3 chunks x 64 rows x 128 B
constexpr int ST_A_OFF = 0, ST_B_OFF = 2 * ST_A_PLANE, ST_P_OFF = ST_B_OFF + 2 * ST_B_PLANE;
constexpr int ST_PATCH_BYTES = ST_PROWS * ST_PWP * 4;
constexpr int ST_SS_OFF = ST_P_OFF + 2 * ST_PATCH_BYTES, ST_BAR_OFF = ST_SS_OFF + 512;
constexpr int ST_SMEM = ST_BAR_OFF + 64 + 1024;


struct StemParams {
    const float* x;       // (n,3,S,S) nchw
    const uint4* w_img;   // packed weight tile: exact shared-memory image (2 planes x 3 chunks x 64 rows x 128 B, swizzled)
    const float* scale;   // folded BN scale * 2^-e
    const float* shift;
    float* out;           // (n,HO,HO,64) nhwc fp32
    int S, HO, tiles_per_row, num_tiles;
};

template <bool SPLIT>
__global__ void __launch_bounds__(256, 1) stem_tc_kernel(const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    float* s_scale = reinterpret_cast<float*>(sm + ST_SS_OFF);
    float* s_shift = s_scale + 64;
    const uint32_t bar_done = base + ST_BAR_OFF;  // [2] MMAs of tile t complete (accumulator ready, A tile free)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + ST_BAR_OFF + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { mbar_init(bar_done, 1); mbar_init(bar_done + 8, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 2 * ST_B_PLANE / 16; i += 256) reinterpret_cast<uint4*>(sm + ST_B_OFF)[i] = __ldg(p.w_img + i);
    if (tid < 64) { s_scale[tid] = __ldg(p.scale + tid); s_shift[tid] = __ldg(p.shift + tid); }
    fence_proxy_async_smem();  // the weight tile was written through the generic proxy, tcgen05.mma reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int t, int& b, int& oy, int& x0) {
        const int seg = t % p.tiles_per_row;
        const int r = t / p.tiles_per_row;
        oy = r % p.HO;
        b = r / p.HO;
        x0 = seg * 128;
    };
    auto load_patch = [&](int t, int buf) {
        int b, oy, x0;
        decode(t, b, oy, x0);
        float* dst = reinterpret_cast<float*>(sm + ST_P_OFF + buf * ST_PATCH_BYTES);
        const float* src = p.x + (size_t)b * 3 * p.S * p.S;
        // one patch row (c, kh) per warp pass, lanes stride over its 261 columns: the address arithmetic is per row, not per element
        // (the flat-index version spent two integer divisions per 4-byte copy and made this loop the largest instruction consumer
        // of the kernel)
        for (int r = warp; r < ST_PROWS; r += 8) {
            const int c = r / 7, kh = r - c * 7;
            const int gy = 2 * oy + kh;
            const bool row_ok = gy < p.S;
            const float* g = src + ((size_t)c * p.S + gy) * p.S + 2 * x0;
            const uint32_t d = smem_u32(dst + r * ST_PWP);
            const int ncol = min(ST_PW, p.S - 2 * x0);  // columns of this row that exist in the image
            for (int col = lane; col < ST_PW; col += 32) {
                if (row_ok && col < ncol) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * col), "l"(g + col) : "memory");
                } else {
                    dst[r * ST_PWP + col] = 0.f;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto epilogue = [&](int t, int it) {
        int b, oy, x0;
        decode(t, b, oy, x0);
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        uint32_t v[32];
        const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (it & 1) * 128 + half * 32;
        tmem_ld32(ta, v);
        if (SPLIT) {  // main (hi*hi) + cross (hi*lo + lo*hi) accumulators, added in fp32 here (see conv_tc.cu, XACC)
            uint32_t x[32];
            tmem_ld32(ta + 64, x);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(x[j]));
        } else {
            tmem_ld_wait();
        }
        if (x0 + row < p.HO) {
            float4* o = reinterpret_cast<float4*>(p.out + (((size_t)b * p.HO + oy) * p.HO + x0 + row) * 64 + half * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 y;
                y.x = fmaxf(fmaf(__uint_as_float(v[4 * q + 0]), s_scale[half * 32 + 4 * q + 0], s_shift[half * 32 + 4 * q + 0]), 0.f);
                y.y = fmaxf(fmaf(__uint_as_float(v[4 * q + 1]), s_scale[half * 32 + 4 * q + 1], s_shift[half * 32 + 4 * q + 1]), 0.f);
                y.z = fmaxf(fmaf(__uint_as_float(v[4 * q + 2]), s_scale[half * 32 + 4 * q + 2], s_shift[half * 32 + 4 * q + 2]), 0.f);
                y.w = fmaxf(fmaf(__uint_as_float(v[4 * q + 3]), s_scale[half * 32 + 4 * q + 3], s_shift[half * 32 + 4 * q + 3]), 0.f);
                o[q] = y;
            }
        }
        tc_fence_before();
    };

    // k' 168..175 (16-byte group 21 of every row) is padding that no build task ever writes: zero it once
    for (int m = tid; m < 128; m += 256) {
        const uint32_t a = ST_A_OFF + (21 >> 3) * (128 * 128) + m * 128 + (((21 & 7) ^ (m & 7)) << 4);
        *reinterpret_cast<uint4*>(sm + a) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sm + ST_A_PLANE + a) = make_uint4(0, 0, 0, 0);
    }
    const uint32_t idesc = make_idesc(128, 64);
    int it = 0, prev_tile = -1;
    if ((int)blockIdx.x < p.num_tiles) load_patch(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // patch(it) is complete and visible; everyone is done with patch(it-1) and epilogue(it-2)
        if (tile + (int)gridDim.x < p.num_tiles) load_patch(tile + gridDim.x, (it + 1) & 1);
        if (it > 0) {  // MMAs of the previous tile have finished: the A tile may be overwritten, its accumulator is ready
            mbar_wait(bar_done + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
            tc_fence_after();
        }
        // ---- software im2col: patch (fp32) -> A tile (split fp16, 128B swizzle, K-major) ----
        const float* patch = reinterpret_cast<const float*>(sm + ST_P_OFF + (it & 1) * ST_PATCH_BYTES);
#pragma unroll 3
        for (int task = tid; task < 128 * ST_PROWS; task += 256) {
            const int m = task & 127, r = task >> 7;  // r = c*7 + kh: the 7 taps kw = 0..6 are 7 contiguous patch floats
            const float2* src = reinterpret_cast<const float2*>(patch + r * ST_PWP + 2 * m);
            const float2 p0 = src[0], p1 = src[1], p2 = src[2];
            const float p6 = patch[r * ST_PWP + 2 * m + 6];
            const float v[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p6, 0.f};
            uint4 hi, lo;
            __half2* hh = reinterpret_cast<__half2*>(&hi);
            __half2* ll = reinterpret_cast<__half2*>(&lo);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                const float2 hf = __half22float2(h);
                hh[e] = h;
                ll[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
            }
            const uint32_t a = ST_A_OFF + (r >> 3) * (128 * 128) + m * 128 + (((r & 7) ^ (m & 7)) << 4);
            *reinterpret_cast<uint4*>(sm + a) = hi;
            if (SPLIT) *reinterpret_cast<uint4*>(sm + ST_A_PLANE + a) = lo;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (it & 1) * 128, tmem_x = tmem_d + 64;
#pragma unroll
            for (int k = 0; k < ST_KUSED / 16; ++k) {
                const uint32_t ao = base + ST_A_OFF + (k >> 2) * (128 * 128) + (k & 3) * 32;
                const uint32_t bo = base + ST_B_OFF + (k >> 2) * (64 * 128) + (k & 3) * 32;
                const uint64_t a_hi = make_smem_desc(ao), b_hi = make_smem_desc(bo);
                umma_f16(tmem_d, a_hi, b_hi, idesc, k ? 1u : 0u);
                if (SPLIT) {
                    umma_f16(tmem_x, a_hi, make_smem_desc(bo + ST_B_PLANE), idesc, k ? 1u : 0u);
                    umma_f16(tmem_x, make_smem_desc(ao + ST_A_PLANE), b_hi, idesc, 1u);
                }
            }
            umma_commit(bar_done + 8 * (it & 1));
        }
        if (it > 0) epilogue(prev_tile, it - 1);  // overlaps the MMAs just issued
        prev_tile = tile;
    }
    if (it > 0) {
        mbar_wait(bar_done + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
        tc_fence_after();
        epilogue(prev_tile, it - 1);
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

int launch_stem_tc(const float* x, int n, int s, const void* w_img, const float* scale_tc, const float* shift, float* out, bool split,
                   cudaStream_t st) {
    static SmemAttrCache attr_split, attr_single;
    if (int rc = attr_split.ensure(stem_tc_kernel<true>, ST_SMEM)) return rc;
    if (int rc = attr_single.ensure(stem_tc_kernel<false>, ST_SMEM)) return rc;
    StemParams p;
    p.x = x; p.w_img = static_cast<const uint4*>(w_img); p.scale = scale_tc; p.shift = shift; p.out = out;
    p.S = s; p.HO = (s - 7) / 2 + 1;
    p.tiles_per_row = (p.HO + 127) / 128;
    p.num_tiles = n * p.HO * p.tiles_per_row;
    if (p.num_tiles == 0) return 0;
    const int num_sms = device_sm_count();
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    if (split) stem_tc_kernel<true><<<grid, 256, ST_SMEM, st>>>(p);
    else stem_tc_kernel<false><<<grid, 256, ST_SMEM, st>>>(p);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t stem_tc_image_bytes() { return (size_t)2 * ST_B_PLANE; }

// w: OIHW (64,3,7,7) flattened [64][147].  Produces the shared-memory image of the B tile (hi plane then lo plane, each
// 3 chunks x 64 rows x 128 B with the 128B swizzle) of w * 2^e[co], and scale_out = scale_in * 2^-e.
void pack_stem_tc_host(const float* w, const float* scale_in, std::vector<uint8_t>& img, std::vector<float>& scale_out) {
    img.assign(2 * ST_B_PLANE, 0);
    scale_out.resize(64);
    for (int co = 0; co < 64; ++co) {
        float mx = 0.f;
        for (int k = 0; k < 147; ++k) mx = std::fmax(mx, std::fabs(w[co * 147 + k]));
        int e = 0;
        if (mx > 0.f && std::isfinite(mx)) { int ex; std::frexp(mx, &ex); e = 8 - ex; }
        const float s = std::ldexp(1.0f, e);
        for (int k = 0; k < 147; ++k) {
            const float v = w[co * 147 + k] * s;
            const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
            const int kp = (k / 7) * 8 + k % 7;  // k = (c*7+kh)*7 + kw  ->  k' = (c*7+kh)*8 + kw
            const int chunk = kp / 64, kk = kp % 64, g = kk / 8, e8 = kk % 8;
            const size_t off = (size_t)chunk * (64 * 128) + co * 128 + ((g ^ (co & 7)) << 4) + e8 * 2;
            memcpy(&img[off], &h, 2);
            memcpy(&img[ST_B_PLANE + off], &l, 2);
        }
        scale_out[co] = scale_in[co] * std::ldexp(1.0f, -e);
    }
}

}  // namespace usot
