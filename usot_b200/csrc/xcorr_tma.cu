// Fused GroupDW depth-wise cross-correlation, TMA-pipelined (the bandwidth-bound half of the headline metric).
//   out[n] = sum_i softmax(w)_i * xcorr(x_i[n / rep_x], z_i[n / rep_z])      lib/models/connect.py:86-102,147-157
//
// One CTA = (output sample, 64-channel slab).  A producer warp streams the three encoded search maps row by row through a
// 4-stage shared-memory ring with 3-D..4-D TMA boxes {64 ch, W, 1 row, 1 sample} (fp32, no swizzle: lanes read consecutive
// channels, conflict-free), so ~3 rows x 21 KB per CTA are always in flight and every byte of x is fetched from HBM/L2 exactly
// once per output sample.  Six consumer warps = 3 column strips x 64 channels; each thread keeps its 55 pre-scaled taps and a
// ring of 5 output rows x 9 columns in registers, accumulates the 5x5 / 5x3 / 3x5 correlations of the current input row, and
// writes a finished output row (coalesced over channels) every step.
#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

namespace usot {

constexpr int GD_STAGES = 2, GD_SW = 9, GD_CONSUMERS = 192;  // 2 stages x ~21 KB + 14 KB taps = 58 KB per CTA -> 3 CTAs per SM

struct GdwParams {
    CUtensorMap m11, m12, m21;
    const float* z11; const float* z12; const float* z21;
    float* out;
    __half2* out_hi;  // FFMA2 kernel: if set, results leave as split-fp16 planes (hi = rn16(v), lo = rn16(v - hi)) instead of fp32
    __half2* out_lo;
    int nx, nz, n_out, C, F, nstrips;
    int row_groups, rows_per_group;  // FFMA2 kernel: output rows split over row_groups CTAs (small batches; 1 = whole map per CTA)
    float w0, w1, w2;
};

__global__ void __launch_bounds__(GD_CONSUMERS + 32, 3) groupdw_tma_kernel(const __grid_constant__ GdwParams p) {
    extern __shared__ __align__(128) uint8_t gsm_raw[];
    const uint32_t base = (smem_u32(gsm_raw) + 127u) & ~127u;
    uint8_t* sm = gsm_raw + (base - smem_u32(gsm_raw));
    const int F = p.F, R = F - 6, W11 = F - 2, H11 = F - 2, W12 = F - 2, H12 = F - 4, W21 = F - 4, C = p.C;
    const int row11 = W11 * 256, row21 = W21 * 256, row12 = W12 * 256;  // bytes of one staged row (64 ch x 4 B per pixel)
    const int stage_bytes = row11 + row21 + row12;
    float* zs = reinterpret_cast<float*>(sm + GD_STAGES * stage_bytes);  // [55][64] taps of this CTA's channels, pre-scaled
    const uint32_t bar_full = base + GD_STAGES * stage_bytes + 55 * 64 * 4, bar_empty = bar_full + 8 * GD_STAGES;

    const int cblocks = C / 64;
    const int cblk = blockIdx.x % cblocks, n = blockIdx.x / cblocks;
    const int xb = n / (p.n_out / p.nx), zb = n / (p.n_out / p.nz);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < GD_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, GD_CONSUMERS / 32); }
        fence_barrier_init();
    }
    if (tid < 64) {  // taps shared by the 3 column strips (registers are the scarce resource: 3 CTAs/SM need <= 96 per thread)
        const int cc = cblk * 64 + tid;
        for (int t = 0; t < 25; ++t) zs[t * 64 + tid] = p.w0 * __ldg(p.z11 + ((size_t)zb * 25 + t) * C + cc);
        for (int t = 0; t < 15; ++t) zs[(25 + t) * 64 + tid] = p.w1 * __ldg(p.z12 + ((size_t)zb * 15 + t) * C + cc);
        for (int t = 0; t < 15; ++t) zs[(40 + t) * 64 + tid] = p.w2 * __ldg(p.z21 + ((size_t)zb * 15 + t) * C + cc);
    }
    __syncthreads();

    if (warp == GD_CONSUMERS / 32) {
        // ------------------------------- producer -------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&p.m11); tma_prefetch_desc(&p.m12); tma_prefetch_desc(&p.m21);
            for (int t = 0; t < H11; ++t) {
                const int s = t % GD_STAGES;
                if (t >= GD_STAGES) mbar_wait(bar_empty + 8 * s, ((t / GD_STAGES) + 1) & 1);
                const bool has12 = t >= 2 && t - 2 < H12;
                const uint32_t full = bar_full + 8 * s, dst = base + s * stage_bytes;
                mbar_expect_tx(full, (uint32_t)(row11 + row21 + (has12 ? row12 : 0)));
                tma_load_4d(dst, &p.m11, full, cblk * 64, 0, t, xb);
                tma_load_4d(dst + row11, &p.m21, full, cblk * 64, 0, t, xb);
                if (has12) tma_load_4d(dst + row11 + row21, &p.m12, full, cblk * 64, 0, t - 2, xb);
            }
        }
        return;
    }

    // ------------------------------- consumers -------------------------------
    const int strip = tid >> 6, ch = tid & 63;
    const int c = cblk * 64 + ch;
    const int j0 = strip * GD_SW;
    const int jn = min(GD_SW, R - j0);
    const float* z = zs + ch;  // tap t of this thread's channel: z[t * 64]
    float* out = p.out + (size_t)n * R * R * C + c;

    float acc[5][GD_SW];  // acc[k] = output row t-4+k at step t
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int j = 0; j < GD_SW; ++j) acc[k][j] = 0.f;

    for (int t = 0; t < H11; ++t) {
        const int s = t % GD_STAGES;
        mbar_wait(bar_full + 8 * s, (t / GD_STAGES) & 1);
        const float* xs = reinterpret_cast<const float*>(sm + s * stage_bytes) + ch;
        {   // 5x5 on x11 row t -> output rows t-u
            float xr[GD_SW + 4];
#pragma unroll
            for (int q = 0; q < GD_SW + 4; ++q) xr[q] = (j0 + q < W11) ? xs[(j0 + q) * 64] : 0.f;
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const float zt = z[(u * 5 + v) * 64];
#pragma unroll
                    for (int j = 0; j < GD_SW; ++j) acc[4 - u][j] = fmaf(xr[j + v], zt, acc[4 - u][j]);
                }
        }
        {   // 5x3 on x21 row t -> output rows t-u
            const float* x2 = xs + W11 * 64;
            float xr[GD_SW + 2];
#pragma unroll
            for (int q = 0; q < GD_SW + 2; ++q) xr[q] = (j0 + q < W21) ? x2[(j0 + q) * 64] : 0.f;
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    const float zt = z[(40 + u * 3 + v) * 64];
#pragma unroll
                    for (int j = 0; j < GD_SW; ++j) acc[4 - u][j] = fmaf(xr[j + v], zt, acc[4 - u][j]);
                }
        }
        if (t >= 2 && t - 2 < H12) {  // 3x5 on x12 row t-2 -> output rows t-2-u (ring slots 2-u)
            const float* x3 = xs + (W11 + W21) * 64;
            float xr[GD_SW + 4];
#pragma unroll
            for (int q = 0; q < GD_SW + 4; ++q) xr[q] = (j0 + q < W12) ? x3[(j0 + q) * 64] : 0.f;
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const float zt = z[(25 + u * 5 + v) * 64];
#pragma unroll
                    for (int j = 0; j < GD_SW; ++j) acc[2 - u][j] = fmaf(xr[j + v], zt, acc[2 - u][j]);
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);  // this warp is done reading the stage
        if (t >= 4) {  // output row t-4 is complete
            float* o = out + ((size_t)(t - 4) * R + j0) * C;
#pragma unroll
            for (int j = 0; j < GD_SW; ++j)
                if (j < jn) o[(size_t)j * C] = acc[0][j];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < GD_SW; ++j) acc[k][j] = acc[k + 1][j];
#pragma unroll
        for (int j = 0; j < GD_SW; ++j) acc[4][j] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// v2: packed-FMA variant (default).  The v1 kernel above is FMA-pipe bound: a 3-register FFMA issues every other cycle per
// SM sub-partition on sm_100, so its 2.25 GFMA per batch-256 launch cost more than the HBM time of the 795 MB it streams.
// Here every consumer thread owns a PAIR of adjacent channels and all accumulators, search values and taps are float2
// operands of `fma.rn.f32x2` (SASS FFMA2): half the FMA-pipe instructions, and every shared-memory read is a conflict-free
// 64-bit load (lanes = consecutive channel pairs).  Each half of an FFMA2 is an IEEE fma.rn, and the accumulation order per
// output element is the same as v1 / the register-staged kernel, so the three variants agree bit for bit.
// CTA = (output sample, 64-channel slab): 3 consumer warps (one column strip each, 32 channel pairs) + 1 TMA producer warp.
// ---------------------------------------------------------------------------------------------
constexpr int G2_STAGES = 2, G2_CONSUMER_WARPS = 3;
// r02: FOUR consumer warps (column strips of 7 / 6) instead of three (9 / 9 / 7): one consumer warp per SM sub-partition.  The kernel is
// FFMA2-issue bound once the step's power cap has pulled the SM clock down (a 3-register FFMA2 issues every other cycle), and with 3 CTAs
// of 3 consumer warps per SM one scheduler carried 3 warps while the others carried 2.  Same fma sequence per output element: bit-identical.
constexpr int G2_CONSUMER_WARPS4 = 4;
Tunable g_groupdw_warps4 = 1;     // tunable "groupdw_warps4": 1 = four-strip variant for response sizes 25 / 27 (default), 0 = three strips
Tunable g_groupdw_row_split = 1;  // tunable "groupdw_row_split": small batches split the output rows of one map over several CTAs

static __device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

// One consumer warp: output columns [j0, j0+SW) (the first `jn` are stored), channel pair `cp` of the slab.
// MASK = false requires j0 + SW <= R (no column of any staged row is out of range), so no load needs a bounds check.
template <int SW, bool MASK>
static __device__ __forceinline__ void g2_consume(const GdwParams& p, const uint8_t* sm, uint32_t bar_full, uint32_t bar_empty,
                                                  int stage_bytes, const float2* z, float2* out, __half2* out_hi, __half2* out_lo, int j0,
                                                  int jn, int lane, int r0, int t_end) {
    const int F = p.F, R = F - 6, W11 = F - 2, W12 = F - 2, H12 = F - 4, W21 = F - 4, C2 = p.C / 2;
    const float2 zero = make_float2(0.f, 0.f);
    float2 acc[5][SW];  // acc[k] = output row t-4+k at step t
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int j = 0; j < SW; ++j) acc[k][j] = zero;

    // This CTA produces output rows [r0, t_end - 4): input rows t = r0 .. t_end-1 (rows below r0 only warm up the ring: what they add
    // to the not-yet-stored slots is what the full-map loop adds too, so every stored row sees the same fma sequence).
    for (int t = r0; t < t_end; ++t) {
        const int lt = t - r0, s = lt % G2_STAGES;
        mbar_wait(bar_full + 8 * s, (lt / G2_STAGES) & 1);
        const float2* xs = reinterpret_cast<const float2*>(sm + s * stage_bytes) + lane;  // pixel q of a staged row: xs[q * 32]
        {   // 5x5 on x11 row t -> output rows t-u
            float2 xr[SW + 4];
#pragma unroll
            for (int q = 0; q < SW + 4; ++q) xr[q] = (!MASK || j0 + q < W11) ? xs[(j0 + q) * 32] : zero;
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const float2 zt = z[(u * 5 + v) * 32];
#pragma unroll
                    for (int j = 0; j < SW; ++j) ffma2(acc[4 - u][j], xr[j + v], zt);
                }
        }
        {   // 5x3 on x21 row t -> output rows t-u
            const float2* x2 = xs + W11 * 32;
            float2 xr[SW + 2];
#pragma unroll
            for (int q = 0; q < SW + 2; ++q) xr[q] = (!MASK || j0 + q < W21) ? x2[(j0 + q) * 32] : zero;
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    const float2 zt = z[(40 + u * 3 + v) * 32];
#pragma unroll
                    for (int j = 0; j < SW; ++j) ffma2(acc[4 - u][j], xr[j + v], zt);
                }
        }
        if (t >= 2 && t - 2 < H12) {  // 3x5 on x12 row t-2 -> output rows t-2-u (ring slots 2-u)
            const float2* x3 = xs + (W11 + W21) * 32;
            float2 xr[SW + 4];
#pragma unroll
            for (int q = 0; q < SW + 4; ++q) xr[q] = (!MASK || j0 + q < W12) ? x3[(j0 + q) * 32] : zero;
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const float2 zt = z[(25 + u * 5 + v) * 32];
#pragma unroll
                    for (int j = 0; j < SW; ++j) ffma2(acc[2 - u][j], xr[j + v], zt);
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);  // this warp is done reading the stage
        if (t - 4 >= r0) {  // output row t-4 is complete: 256 contiguous bytes per warp and column
            const size_t o0 = ((size_t)(t - 4) * R + j0) * C2;
            if (out_hi) {  // 128 contiguous bytes per warp, column and plane
#pragma unroll
                for (int j = 0; j < SW; ++j)
                    if (!MASK || j < jn) {
                        const __half2 h = __floats2half2_rn(acc[0][j].x, acc[0][j].y);
                        const float2 hf = __half22float2(h);
                        out_hi[o0 + (size_t)j * C2] = h;
                        if (out_lo) out_lo[o0 + (size_t)j * C2] = __floats2half2_rn(acc[0][j].x - hf.x, acc[0][j].y - hf.y);
                    }
            } else {
                float2* o = out + o0;
#pragma unroll
                for (int j = 0; j < SW; ++j)
                    if (!MASK || j < jn) o[(size_t)j * C2] = acc[0][j];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < SW; ++j) acc[k][j] = acc[k + 1][j];
#pragma unroll
        for (int j = 0; j < SW; ++j) acc[4][j] = zero;
    }
}

template <int NCW>
__global__ void __launch_bounds__(NCW * 32 + 32, 3) groupdw_ffma2_kernel(const __grid_constant__ GdwParams p) {
    extern __shared__ __align__(128) uint8_t gsm_raw[];
    const uint32_t base = (smem_u32(gsm_raw) + 127u) & ~127u;
    uint8_t* sm = gsm_raw + (base - smem_u32(gsm_raw));
    const int F = p.F, R = F - 6, W11 = F - 2, H11 = F - 2, W12 = F - 2, H12 = F - 4, W21 = F - 4, C = p.C;
    const int row11 = W11 * 256, row21 = W21 * 256, row12 = W12 * 256;  // bytes of one staged row (64 ch x 4 B per pixel)
    const int stage_bytes = row11 + row21 + row12;
    float* zs = reinterpret_cast<float*>(sm + G2_STAGES * stage_bytes);  // [55][64] taps of this CTA's channels, pre-scaled
    const uint32_t bar_full = base + G2_STAGES * stage_bytes + 55 * 64 * 4, bar_empty = bar_full + 8 * G2_STAGES;

    const int cblocks = C / 64;
    const int rg = blockIdx.x % p.row_groups, cb = blockIdx.x / p.row_groups;
    const int cblk = cb % cblocks, n = cb / cblocks;
    const int xb = n / (p.n_out / p.nx), zb = n / (p.n_out / p.nz);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = rg * p.rows_per_group;                          // first output row of this CTA
    const int t_end = min(H11, min(R, r0 + p.rows_per_group) + 4);  // one past the last input row it needs

    if (tid == 0) {
        for (int s = 0; s < G2_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, NCW); }
        fence_barrier_init();
    }
    if (tid < 64) {
        const int cc = cblk * 64 + tid;
        for (int t = 0; t < 25; ++t) zs[t * 64 + tid] = p.w0 * __ldg(p.z11 + ((size_t)zb * 25 + t) * C + cc);
        for (int t = 0; t < 15; ++t) zs[(25 + t) * 64 + tid] = p.w1 * __ldg(p.z12 + ((size_t)zb * 15 + t) * C + cc);
        for (int t = 0; t < 15; ++t) zs[(40 + t) * 64 + tid] = p.w2 * __ldg(p.z21 + ((size_t)zb * 15 + t) * C + cc);
    }
    __syncthreads();

    if (warp == NCW) {
        // ------------------------------- producer -------------------------------
        // (all 32 lanes walk the loop; expect_tx + the three row loads are issued by one ELECTED lane, see elect_one() in tc_ptx.cuh)
        if (lane == 0) { tma_prefetch_desc(&p.m11); tma_prefetch_desc(&p.m12); tma_prefetch_desc(&p.m21); }
        __syncwarp();
        for (int t = r0; t < t_end; ++t) {
            const int lt = t - r0, s = lt % G2_STAGES;
            if (lt >= G2_STAGES) mbar_wait(bar_empty + 8 * s, ((lt / G2_STAGES) + 1) & 1);
            const bool has12 = t >= 2 && t - 2 < H12;
            const uint32_t full = bar_full + 8 * s, dst = base + s * stage_bytes;
            if (elect_one()) {
                mbar_expect_tx(full, (uint32_t)(row11 + row21 + (has12 ? row12 : 0)));
                tma_load_4d(dst, &p.m11, full, cblk * 64, 0, t, xb);
                tma_load_4d(dst + row11, &p.m21, full, cblk * 64, 0, t, xb);
                if (has12) tma_load_4d(dst + row11 + row21, &p.m12, full, cblk * 64, 0, t - 2, xb);
            }
            __syncwarp();
        }
        return;
    }

    // ------------------------------- consumers: warp = column strip, lane = channel pair -------------------------------
    const float2* z = reinterpret_cast<const float2*>(zs) + lane;                           // tap t of this pair: z[t * 32]
    const size_t obase = ((size_t)n * R * R * C + cblk * 64) / 2 + lane;  // in channel pairs
    float2* out = reinterpret_cast<float2*>(p.out) + obase;
    __half2* out_hi = p.out_hi ? p.out_hi + obase : nullptr;
    __half2* out_lo = p.out_lo ? p.out_lo + obase : nullptr;
    if (NCW == 4) {   // R = 25: strips 7 | 6 | 6 | 6;  R = 27: 7 | 7 | 7 | 6   (the launcher admits only these two sizes)
        const int wide = R - 24;                       // number of 7-wide strips (1 or 3)
        const int j4 = warp <= wide ? warp * 7 : wide * 7 + (warp - wide) * 6;
        if (warp < wide) g2_consume<7, false>(p, sm, bar_full, bar_empty, stage_bytes, z, out, out_hi, out_lo, j4, 7, lane, r0, t_end);
        else g2_consume<6, false>(p, sm, bar_full, bar_empty, stage_bytes, z, out, out_hi, out_lo, j4, 6, lane, r0, t_end);
        return;
    }
    const int j0 = warp * 9;
    const int jn = min(9, R - j0);
    if (jn == 9) g2_consume<9, false>(p, sm, bar_full, bar_empty, stage_bytes, z, out, out_hi, out_lo, j0, jn, lane, r0, t_end);
    else if (jn == 7) g2_consume<7, false>(p, sm, bar_full, bar_empty, stage_bytes, z, out, out_hi, out_lo, j0, jn, lane, r0, t_end);
    else g2_consume<9, true>(p, sm, bar_full, bar_empty, stage_bytes, z, out, out_hi, out_lo, j0, jn, lane, r0, t_end);
}

int launch_groupdw_tma(const GroupDWArgs& a, float w0, float w1, float w2, cudaStream_t st) {
    USOT_REQUIRE(a.C % 64 == 0, "groupdw needs C % 64 == 0");
    USOT_REQUIRE(a.nx > 0 && a.nz > 0 && a.n_out % a.nx == 0 && a.n_out % a.nz == 0, "groupdw: n_out must be a multiple of both batches");
    const int F = a.F, R = F - 6;
    USOT_REQUIRE(R > 0 && (R + GD_SW - 1) / GD_SW == 3, "groupdw_tma supports response sizes 19..27 (search 255 / 271)");
    if (a.n_out == 0) return 0;
    GdwParams p;
    memset(&p, 0, sizeof(p));
    auto mk = [&](CUtensorMap* m, const float* base, int h, int w) -> int {
        cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)a.nx};
        cuuint64_t strides[3] = {(cuuint64_t)a.C * 4, (cuuint64_t)w * a.C * 4, (cuuint64_t)h * w * a.C * 4};
        cuuint32_t box[4] = {64, (cuuint32_t)w, 1, 1};
        return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    };
    if (int rc = mk(&p.m11, a.x11, F - 2, F - 2)) return rc;
    if (int rc = mk(&p.m12, a.x12, F - 4, F - 2)) return rc;
    if (int rc = mk(&p.m21, a.x21, F - 2, F - 4)) return rc;
    p.z11 = a.z11; p.z12 = a.z12; p.z21 = a.z21; p.out = a.out;
    p.out_hi = reinterpret_cast<__half2*>(a.out_hi); p.out_lo = reinterpret_cast<__half2*>(a.out_lo);
    USOT_REQUIRE(!a.out_hi || g_groupdw_tma >= 2, "split-fp16 GroupDW output needs the FFMA2 kernel");
    p.nx = a.nx; p.nz = a.nz; p.n_out = a.n_out; p.C = a.C; p.F = F; p.nstrips = 3;
    p.w0 = w0; p.w1 = w1; p.w2 = w2;
    const int smem = GD_STAGES * (3 * F - 8) * 256 + 55 * 64 * 4 + 2 * GD_STAGES * 8 + 128;
    if (g_groupdw_tma >= 2) {
        const int smem2 = G2_STAGES * (3 * F - 8) * 256 + 55 * 64 * 4 + 2 * G2_STAGES * 8 + 128;
        static SmemAttrCache attr2, attr4;
        const bool four = g_groupdw_warps4 && (R == 25 || R == 27);
        if (int rc = four ? attr4.ensure(groupdw_ffma2_kernel<G2_CONSUMER_WARPS4>, smem2) : attr2.ensure(groupdw_ffma2_kernel<G2_CONSUMER_WARPS>, smem2)) return rc;
        // Latency mode (small batches): fewer (sample, slab) pairs than SMs -> split the output rows over several CTAs.  Each extra
        // CTA re-reads 4 warm-up input rows (L2 hits); results are bit-identical to the whole-map CTA (same fma sequence per row).
        const int pairs = a.n_out * (a.C / 64);
        p.row_groups = 1;
        const int sms = device_sm_count();
        if (g_groupdw_row_split && pairs * 2 <= sms) p.row_groups = std::min(R, sms / pairs);
        p.rows_per_group = (R + p.row_groups - 1) / p.row_groups;
        p.row_groups = (R + p.rows_per_group - 1) / p.rows_per_group;
        if (four) groupdw_ffma2_kernel<G2_CONSUMER_WARPS4><<<(unsigned)(pairs * p.row_groups), G2_CONSUMER_WARPS4 * 32 + 32, smem2, st>>>(p);
        else groupdw_ffma2_kernel<G2_CONSUMER_WARPS><<<(unsigned)(pairs * p.row_groups), G2_CONSUMER_WARPS * 32 + 32, smem2, st>>>(p);
        USOT_CUDA_OK(cudaGetLastError());
        return 0;
    }
    static SmemAttrCache attr1;
    if (int rc = attr1.ensure(groupdw_tma_kernel, smem)) return rc;
    groupdw_tma_kernel<<<(unsigned)(a.n_out * (a.C / 64)), GD_CONSUMERS + 32, smem, st>>>(p);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace usot
