// Dense convolutions on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
//
// Every 1x1 / 3x3 conv of the USOT forward path (lib/models/modules.py:37-58,121-126; connect.py:20-53,112-121,178-209,
// 287-290) is an implicit GEMM  D[m][n] = sum_k A[m][k] * B[n][k]  with
//     m = output pixel inside a (BH x BW) spatial patch of one image (<= 128 rows),
//     n = output channel,  k = (filter tap, input channel).
// A is never materialised: for each (tap, 64-channel chunk) ONE 4-D TMA box load [64 ch, BW, BH, 1 image] of the NHWC fp16
// activation tensor lands in shared memory already in the 128-byte-swizzled K-major layout that tcgen05.mma consumes;
// padding comes from TMA out-of-bounds zero fill (coordinates may be negative), dilation is a coordinate offset, and
// stride 2 uses four "parity-decimated" views of the same tensor (strides doubled) so that the box stays dense.
// B (weights, [cout][taps*cin] fp16, K-major) is a 2-D TMA box.  Accumulators live in TMEM (fp32), double buffered so the
// epilogue of tile i overlaps the MMAs of tile i+1.  Persistent CTAs, warp-specialised:
//     warp 0  TMA producer (one elected lane)      warp 1  MMA issuer (one elected lane)
//     warp 2  TMEM allocator                        warps 4-7  epilogue (tcgen05.ld -> scale/shift/residual/ReLU -> global)
//
// Precision modes
//   SPLIT (fp16x3, default): activations and weights are stored as value = hi + lo (two fp16 planes); three MMAs
//       hi*hi + hi*lo + lo*hi accumulate in fp32 in the same TMEM tile.  Per-product relative error ~2^-21, i.e. fp32-class,
//       which the 1e-3 / exact-argmax parity bar needs (single-pass TF32/fp16 does not meet it on random residual nets,
//       SURVEY.md App. C).  Weights are pre-scaled by a per-output-channel power of two (folded back in the epilogue scale)
//       so that both fp16 planes stay in the normal range.
//   single fp16: only the hi planes are read: 3x fewer MMAs, ~1e-3-class error per layer (fast mode).
#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace usot {

// =============================================================================================
// Kernel
// =============================================================================================
constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_RES_BUFS = 6;                           // residual-landed barriers per epilogue group (upper bound of TcParams::nbuf)
constexpr int TC_EPI_GROUPS = 2;                          // epilogue warp groups (4 warps each, one per TMEM lane quarter)
constexpr int TC_THREADS = 128 + 128 * TC_EPI_GROUPS;     // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4..: epilogue
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KiB per plane per stage

// PAIR (cta_group::2): a CTA stages only ITS half of the weight rows of the N block (the pair's MMA reads both halves), see conv_tc_kernel.
template <int BN, bool SPLIT, int PAIR = 0>
struct TcCfg {
    static constexpr int PLANES = SPLIT ? 2 : 1;
    static constexpr int B_ROWS = PAIR ? BN / 2 : BN;     // weight rows staged by one CTA per K-step and plane
    static constexpr int B_BYTES = B_ROWS * TC_BK * 2;
    static constexpr int B_REGION = PLANES * B_BYTES;
    static constexpr int STAGE_BYTES = PLANES * TC_A_BYTES + B_REGION;
    static constexpr int SMEM_BUDGET = 227 * 1024 - 1024 - 256;  // alignment slack + barriers
    static constexpr bool TMA_OUT = !(SPLIT && BN == 256);  // (the 2-stage 96 KiB ring of that config leaves no room for staging)
    static constexpr int OUT_STAGE_BYTES = TMA_OUT ? 2 * 2 * TC_BM * 64 : 0;  // 2 buffers x 2 planes x 128 rows x 64 B
    static constexpr int STAGES_RAW = (SMEM_BUDGET - 2 * BN * 4 - OUT_STAGE_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    // Split mode keeps TWO accumulators per tile when TMEM allows it (BN <= 128): `main` receives only the hi*hi products and
    // `cross` the 2^-11-times smaller hi*lo + lo*hi products.  The tensor core truncates on every accumulate, so the error of
    // an accumulator grows with the number of MMAs added into it; routing the cross terms elsewhere cuts the adds into `main`
    // 3x and makes the cross terms' own truncation negligible.  The epilogue adds the two in fp32 (round-to-nearest).
    static constexpr bool XACC = SPLIT && BN <= 128;
    static constexpr int ACC_COLS = XACC ? 2 * BN : BN;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;  // double buffered; power of two for BN in {64,128,256}
    static constexpr int BUF_BYTES = 2 * TC_BM * 64;  // one staging buffer: 2 planes x 128 rows x 64 B
    static constexpr int MISC_BYTES = 2 * BN * 4 + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int stages_with_bufs(int nbuf) { return (227 * 1024 - MISC_BYTES - TC_EPI_GROUPS * nbuf * BUF_BYTES) / STAGE_BYTES; }
    static constexpr int smem_bytes(int stages, int nbuf) { return stages * STAGE_BYTES + TC_EPI_GROUPS * nbuf * BUF_BYTES * (TMA_OUT ? 1 : 0) + MISC_BYTES; }
    // ring depth when every epilogue group rotates 3 staging buffers (TMA-prefetched residual)
    static constexpr int STAGES_RES = (227 * 1024 - MISC_BYTES - TC_EPI_GROUPS * 3 * BUF_BYTES) / STAGE_BYTES;
};

// One index of the persistent tile loop is a SEQUENCE of `count` accumulator tiles that the same CTA walks back to back:
//   ordinary conv            count = 1
//   fused Conf_Fusion        the p.group (= N_q) memory maps of one sample for a fixed (patch, N block): image advances
//   fused stem + max-pool    the conv rows first..last that one band of pooled rows of one image needs: the tile row advances
struct TileSeq { int nb, tw, th, img, count, dth, dimg; };
static __device__ __forceinline__ TileSeq decode_seq(const TcParams& p, int tile) {
    TileSeq s;
    if (p.pool_bands > 0) {
        const int band = tile % p.pool_bands;
        const int p0 = band * p.pool_band_rows, p1 = min(p.pool_po, p0 + p.pool_band_rows);
        const int first = p0 > 0 ? 2 * p0 - 1 : 0, last = min(2 * p1 - 1, p.ho - 1);   // pooled row py reads conv rows 2py-1 .. 2py+1
        s.nb = 0; s.tw = 0; s.th = first; s.img = tile / p.pool_bands; s.count = last - first + 1; s.dth = 1; s.dimg = 0;
        return s;
    }
    s.nb = tile % p.n_tiles_n;
    int mt = tile / p.n_tiles_n;
    s.tw = mt % p.tiles_w; mt /= p.tiles_w;
    s.th = mt % p.tiles_h;
    s.img = (mt / p.tiles_h) * p.group * p.bimg;   // first image of the tile (bimg == 1 whenever group > 1)
    s.count = p.group; s.dth = 0; s.dimg = 1;
    return s;
}

// EPI = 0: the ordinary epilogue (scale/shift, residual, ReLU -> split-fp16 / fp32 maps).
// EPI = 2: fused stem + max-pool 3x3/2 p1 (modules.py:70-74,138-141).  The stem runs as a TMA-fed implicit GEMM over the space-to-depth
//          image (see launch_stem_s2d_pool); a tile is ONE conv row (bh = 1), a CTA walks the rows of a band top to bottom, every epilogue
//          thread (pixel, 32 channels) keeps the running vertical maximum in registers, and every second row the row of vertical maxima
//          goes through a swizzled shared-memory buffer for the horizontal 3-tap / stride-2 maximum and leaves as the split-fp16 operand
//          planes layer1 reads.  The (n, 125, 125, 64) conv map never exists in HBM.
// EPI = 1: fused Conf_Fusion (connect.py:123-144).  The weight tile of N-block nb holds BN/2 channels of conf_gen followed by the SAME
//          BN/2 channels of value_gen, a CTA walks the p.group (= N_q) memory maps of one sample back to back for a fixed (patch, nb),
//          and the epilogue warps keep  sum_q e_q * v_q  and  sum_q e_q  (e = exp(clamp(conf)), v = value) in registers: neither the
//          confidence nor the value maps ever reach HBM.  Same arithmetic, in the same order, as the two convs + conf_fusion_kernel.
// PAIR = true (EPI = 0 only): the kernel is launched in clusters of two CTAs (the two SMs of a TPC) that execute every MMA together
// (tcgen05.mma.cta_group::2, M = 256): CTA rank r of the pair owns the M tile of image group 2*gp + r at the same (patch, N block), loads
// its own activation boxes and rows [r*BN/2, (r+1)*BN/2) of the weight tile; the leader (rank 0) issues the MMAs for both.  A weight byte
// then crosses L2 -> SM once per PAIR of M tiles and each SM reads only half of the B operand from its shared memory; the operand ring gets
// deeper (a stage is 16 + 16 KiB instead of 16 + 32 in single-fp16 mode, 32 + 16 instead of 32 + 32 in split mode).  Every output element
// sees the same K order and the same accumulators as in the one-CTA kernel: results are bit-identical (tests/test_gpu_tunables.py).
template <int BN, bool SPLIT, int EPI, int PAIR = 0>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
    static_assert(!PAIR || EPI == 0, "the CTA-pair variant exists for the ordinary epilogue only");
    using Cfg = TcCfg<BN, SPLIT, PAIR>;
    // pair mode: tile-loop indices count PAIR tiles; tile_of() gives this CTA's ordinary tile index (possibly a phantom tile past the
    // last image group when the group count is odd: its loads are zero-filled by TMA, its stores clipped)
    // (macros, not variables: each warp role evaluates them where it needs them, so the one-CTA kernels keep their register allocation)
#define cta_rank (PAIR ? cluster_ctarank() : 0u)
#define t_first (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x)
#define t_step (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x)
#define t_count (PAIR ? p.num_pair_tiles : p.num_tiles)
    auto tile_of = [&p](int t) -> int {
        if constexpr (!PAIR) return t;
        const int nb_ = t % p.n_tiles_n, r_ = t / p.n_tiles_n;
        const int sp_ = p.tiles_w * p.tiles_h, s_ = r_ % sp_, gp_ = r_ / sp_;
        return ((2 * gp_ + (int)cta_rank) * sp_ + s_) * p.n_tiles_n + nb_;
    };
    const int STAGES = p.stages;
    constexpr int MAXST = 6;  // barrier slots
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned operand ring (required by the 128B swizzle atoms)
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // (EPI = 2, fused stem: the operand region is ROLL_R rolling activation rows + the resident filter, followed by the 768-byte
    //  cross-warp edge buffer of the pooling epilogue instead of store-staging buffers)
    constexpr int ROLL_R = SPLIT ? 5 : 6;
    constexpr uint32_t ROLL_SLOT = Cfg::PLANES * TC_A_BYTES, ROLL_B = Cfg::PLANES * Cfg::B_BYTES, ROLL_EDGE = 768;
    uint8_t* s_out = smem_gen + (EPI == 2 ? ROLL_R * ROLL_SLOT + 4 * ROLL_B : STAGES * Cfg::STAGE_BYTES);  // 1024-aligned
    float* s_scale = reinterpret_cast<float*>(s_out + (EPI == 2 ? ROLL_EDGE : (Cfg::TMA_OUT ? TC_EPI_GROUPS * p.nbuf * Cfg::BUF_BYTES : 0)));
    float* s_shift = s_scale + BN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + BN);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * MAXST, bar_tfull = bar_empty + 8 * MAXST,
                   bar_tempty = bar_tfull + 16;
    const uint32_t bar_res = bar_tempty + 16;  // [TC_EPI_GROUPS][TC_RES_BUFS] residual chunk landed in staging buffer b of group g
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAXST + 4 + TC_RES_BUFS * TC_EPI_GROUPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nK = p.taps * p.cin_chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b[0]);
        tma_prefetch_desc(&p.a[0][0]);
        if (SPLIT) { tma_prefetch_desc(&p.b[1]); tma_prefetch_desc(&p.a[1][0]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < (EPI == 2 ? ROLL_R : STAGES); ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, (PAIR ? 2 : 1) * 4 * TC_EPI_GROUPS); }   // (pair: the epilogue warps of BOTH CTAs release the leader's accumulator)
        for (int s = 0; s < TC_RES_BUFS * TC_EPI_GROUPS; ++s) mbar_init(bar_res + 8 * s, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) {   // (the same warp of both CTAs allocates; each CTA gets its own 128 lanes x TMEM_COLS columns at the same base)
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA completion targets them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (p.pdl) {
        // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no tensor of
        // the previous layer, so it may overlap that layer's tail.  Let OUR successor start its own prologue now, then wait until
        // every grid we depend on has completed and flushed before the first activation / residual byte is read or written.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }

    // Fused stem (EPI = 2): the operand region is NOT a ring of (A, B) stages.  The whole filter (4 K-chunks) stays resident, and the A
    // side is a rolling window of space-to-depth rows: conv row oy reads s2d rows oy..oy+3 (K-chunk ky = row oy+ky), and the CTA walks
    // the rows of a band top to bottom, so every row is loaded once per band instead of four times -- six times less L2 -> SM traffic than
    // re-fetching (A, B) per K-step, which is what bounded the first version of this kernel (ncu: 11.7 TB/s of TMA reads, tensor pipe 39 %).
    // (a window of exactly four rows would leave ONE load in flight per CTA, issued only after chunk 0 of a conv row has retired: measured,
    //  the per-row load latency (~1.7 us for the two 16 KiB boxes) then paced the whole kernel; ROLL_R - 4 extra slots keep loads ahead.)
    //   slot(L) = L % ROLL_R for the L-th row load of this CTA; a slot is free again once the MMAs of its LAST reader are done: chunk ky = 0 of
    //   the conv row that starts at it (or, for the three bottom rows of a band, chunks ky = 1..3 of the band's last conv row).
    const uint32_t roll_b_base = smem_base + ROLL_R * ROLL_SLOT;
    if (EPI == 2 && warp == 0) {
        {   // (all 32 lanes walk the loop; the single-thread instructions are predicated on elect_one())
            const uint32_t a_box_bytes = (uint32_t)p.bw * TC_BK * 2;
            if (elect_one()) {
                mbar_expect_tx(bar_res, 4 * ROLL_B);
                for (int ky = 0; ky < 4; ++ky) {
                    tma_load_2d(roll_b_base + ky * ROLL_B, &p.b[0], bar_res, ky * TC_BK, 0);
                    if (SPLIT) tma_load_2d(roll_b_base + ky * ROLL_B + Cfg::B_BYTES, &p.b[1], bar_res, ky * TC_BK, 0);
                }
            }
            uint32_t L = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileSeq sq = decode_seq(p, tile);
                for (int r = 0; r < sq.count + 3; ++r, ++L) {
                    const uint32_t slot = L % ROLL_R, full = bar_full + 8 * slot;
                    mbar_wait(bar_empty + 8 * slot, ((L / ROLL_R) & 1) ^ 1);
                    const uint32_t sa = smem_base + slot * ROLL_SLOT;
                    if (elect_one()) {
                        mbar_expect_tx(full, Cfg::PLANES * a_box_bytes);
                        if (p.a_rank5) {
                            tma_load_5d(sa, &p.a[0][0], full, 0, 0, 0, sq.th + r, sq.img);
                            if (SPLIT) tma_load_5d(sa + TC_A_BYTES, &p.a[1][0], full, 0, 0, 0, sq.th + r, sq.img);
                        } else {
                            tma_load_4d(sa, &p.a[0][0], full, 0, 0, sq.th + r, sq.img);
                            if (SPLIT) tma_load_4d(sa + TC_A_BYTES, &p.a[1][0], full, 0, 0, sq.th + r, sq.img);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (EPI == 2 && warp == 1) {
        {
            const uint32_t idesc = make_idesc(TC_BM, BN), idesc2 = make_idesc(TC_BM, 2 * BN);
            const bool fuse = Cfg::XACC && p.fuse_cross;
            mbar_wait(bar_res, 0);   // the filter is resident
            tc_fence_after();
            uint32_t base = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int count = decode_seq(p, tile).count;
                for (int q = 0; q < count; ++q, ++it) {
                    const int as = it & 1;
                    mbar_wait(bar_tempty + 8 * as, ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * Cfg::ACC_COLS;
                    const uint32_t tmem_x = Cfg::XACC ? tmem_d + BN : tmem_d;
                    for (int ky = 0; ky < 4; ++ky) {
                        const uint32_t L = base + q + ky, slot = L % ROLL_R;
                        mbar_wait(bar_full + 8 * slot, (L / ROLL_R) & 1);   // (returns at once for the rows earlier conv rows already waited for)
                        tc_fence_after();
                        const uint32_t sa = smem_base + slot * ROLL_SLOT, sb = roll_b_base + ky * ROLL_B;
                        if (elect_one()) {
                            // (descriptors of the four k16 slices / the lo planes are the base descriptor plus a constant: the start-address
                            //  field counts 16-byte units and shared-memory addresses stay far below its 14 bits)
                            const uint64_t a0 = make_smem_desc(sa), b0 = make_smem_desc(sb);
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; ++k) {
                                const uint64_t a_hi = a0 + 2 * k, b_hi = b0 + 2 * k;
                                if (SPLIT && fuse) {
                                    umma_f16(tmem_d, a_hi, b_hi, idesc2, (ky | k) ? 1u : 0u);
                                    umma_f16(tmem_x, a_hi + (TC_A_BYTES >> 4), b_hi, idesc, 1u);
                                    continue;
                                }
                                umma_f16(tmem_d, a_hi, b_hi, idesc, (ky | k) ? 1u : 0u);
                                if (SPLIT) {
                                    umma_f16(tmem_x, a_hi, b_hi + (Cfg::B_BYTES >> 4), idesc, Cfg::XACC ? ((ky | k) ? 1u : 0u) : 1u);
                                    umma_f16(tmem_x, a_hi + (TC_A_BYTES >> 4), b_hi, idesc, 1u);
                                }
                            }
                            if (ky == 0 || q == count - 1) umma_commit(bar_empty + 8 * slot);   // last reader of this row is done
                            if (ky == 3) umma_commit(bar_tfull + 8 * as);
                        }
                        __syncwarp();
                    }
                }
                base += count + 3;
            }
        }
    } else if (warp == 0) {
        // ================================ TMA producer ================================
        {   // (all 32 lanes walk the loop; the single-thread instructions are predicated on elect_one(), see tc_ptx.cuh)
            const uint32_t a_box_bytes = (uint32_t)p.bw * p.bh * p.bimg * TC_BK * 2;
            const uint32_t tx = (PAIR ? 2u : 1u) * (Cfg::PLANES * a_box_bytes + Cfg::B_REGION);   // (pair: the leader's barrier counts both CTAs' bytes)
            int stage = 0;
            uint32_t phase = 0;
            for (int tidx = t_first; tidx < t_count; tidx += t_step) {
                const int tile = tile_of(tidx);
                const TileSeq sq = decode_seq(p, tile);
                const int nb = sq.nb, w0 = sq.tw * p.bw;
                for (int q = 0; q < sq.count; ++q) {   // (count == 1 except in the fused Conf_Fusion / stem + max-pool launches)
                const int img = sq.img + q * sq.dimg;
                const int h0 = (sq.th + q * sq.dth) * p.bh;
                for (int ks = 0; ks < nK; ++ks) {
                    const int tap = ks / p.cin_chunks, cc = ks - tap * p.cin_chunks;
                    const int kh = tap / p.kw, kwi = tap - kh * p.kw;
                    int offh = kh * p.dh - p.ph, offw = kwi * p.dw - p.pw, par = 0;
                    // One-row tiles (bh == 1): a filter row that falls into the zero padding contributes exactly zero to every pixel of
                    // the tile -- neither loaded nor multiplied (the MMA issuer skips the same K-steps; adding 0 is exact, so results
                    // are unchanged).  2 of 31 rows x 1/3 of the taps for a pad-1 3x3 layer, 4 of 31 for the dilation-2 layers.
                    if (p.skip_pad_rows && (unsigned)(h0 + offh) >= (unsigned)p.hin) continue;
                    if (p.stride == 2) {
                        const int py = offh & 1, px = offw & 1;
                        par = py * 2 + px;
                        offh = (offh - py) >> 1;
                        offw = (offw - px) >> 1;
                    }
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::PLANES * TC_A_BYTES;
                    if constexpr (PAIR) {
                        // own activation boxes + own half of the weight rows -> own shared memory; the bytes complete on the LEADER's barrier
                        const uint32_t lfull = mapa_cluster(full, 0);
                        const int brow = nb * BN + (int)cta_rank * Cfg::B_ROWS;
                        if (elect_one()) {
                            if (cta_rank == 0) mbar_expect_tx(full, tx);
                            tma_load_4d_pair(sa, &p.a[0][par], lfull, cc * TC_BK, w0 + offw, h0 + offh, img);
                            tma_load_2d_pair(sb, &p.b[0], lfull, ks * TC_BK, brow);
                            if (SPLIT) {
                                tma_load_4d_pair(sa + TC_A_BYTES, &p.a[1][par], lfull, cc * TC_BK, w0 + offw, h0 + offh, img);
                                tma_load_2d_pair(sb + Cfg::B_BYTES, &p.b[1], lfull, ks * TC_BK, brow);
                            }
                        }
                    } else if (elect_one()) {
                        mbar_expect_tx(full, tx);
                        tma_load_4d(sa, &p.a[0][par], full, cc * TC_BK, w0 + offw, h0 + offh, img);
                        tma_load_2d(sb, &p.b[0], full, ks * TC_BK, nb * BN);
                        if (SPLIT) {
                            tma_load_4d(sa + TC_A_BYTES, &p.a[1][par], full, cc * TC_BK, w0 + offw, h0 + offh, img);
                            tma_load_2d(sb + Cfg::B_BYTES, &p.b[1], full, ks * TC_BK, nb * BN);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                }
                // DRAM-bound 1x1 layers (K <= 1024, ring only 2-4 stages deep): once every load of THIS tile is queued, pull the NEXT
                // tile's activation boxes into L2 (hints queue behind the demand loads), so its ring loads are L2 hits; every byte
                // is still fetched from HBM once.  (3x3 layers re-read A from L2 per tap anyway.)
                if (!PAIR && p.l2_prefetch && p.group == 1 && p.pool_bands == 0 && p.taps == 1 && p.stride == 1 && nb == p.n_tiles_n - 1) {
                    const int nt = tile + gridDim.x;  // same M tile is shared by the n_tiles_n consecutive tiles: prefetch once
                    if (nt < p.num_tiles) {
                        int mt2 = nt / p.n_tiles_n;
                        const int tw2 = mt2 % p.tiles_w; mt2 /= p.tiles_w;
                        const int th2 = mt2 % p.tiles_h;
                        const int img2 = (mt2 / p.tiles_h) * p.bimg;
                        if (elect_one())
                            for (int cc2 = 0; cc2 < p.cin_chunks; ++cc2) {
                                tma_prefetch_l2_4d(&p.a[0][0], cc2 * TC_BK, tw2 * p.bw - p.pw, th2 * p.bh - p.ph, img2);
                                if (SPLIT) tma_prefetch_l2_4d(&p.a[1][0], cc2 * TC_BK, tw2 * p.bw - p.pw, th2 * p.bh - p.ph, img2);
                            }
                        __syncwarp();
                    }
                }
            }
            if constexpr (PAIR) {
                // tail: the leader's commits arrive on THIS CTA's slot barriers too; wait for the last one of every slot so that no
                // multicast arrival is still in flight towards a CTA that has already exited
                for (int s2 = 0; s2 < STAGES; ++s2) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (!PAIR || cta_rank == 0) {   // (pair: the leader issues for both CTAs; all 32 lanes walk the loop and wait on the barriers; one elected lane issues the MMAs and commits of a K-step)
            const uint32_t idesc = make_idesc(PAIR ? 2 * TC_BM : TC_BM, BN);
            // Fused cross term: the lo weight tile sits right behind the hi tile in the stage and `cross` right behind `main` in
            // TMEM, so ONE MMA with N = 2*BN computes a_hi*w_hi -> main and a_hi*w_lo -> cross while reading a_hi from shared
            // memory once (the operand reads of three N=128 MMAs per k-step saturate the 128 B/clk shared-memory port).
            const uint32_t idesc2 = make_idesc(TC_BM, 2 * BN);
            // (pair: [w_hi | w_lo] of one CTA are not the two halves of an N = 2*BN operand -- the hardware takes rows [0, N/2) from the
            //  leader and [N/2, N) from the peer -- so the three products are issued as three N = BN MMAs, the tc_fuse_cross = 0 sequence.
            //  A fused pair layout was built and measured: region X = w_hi in the leader / w_lo in the peer for ONE N = 2*BN MMA, region Y =
            //  the w_hi halves for a_lo*w_hi, 56 KiB stages, three of them.  Bit-identical, but 2-3 % MORE cycles on the MMA-bound layers than
            //  this version with its four 48 KiB stages (layer3.0.downsample 6282 k vs 6175 k cycles; profiles/r02c_pair_*): removed.)
            const bool fuse = !PAIR && Cfg::XACC && p.fuse_cross;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tidx = t_first; tidx < t_count; tidx += t_step) {
            const TileSeq sq = decode_seq(p, tile_of(tidx));   // (pair: the peer's tile has the same patch row, hence the same skipped K-steps)
            for (int q = 0; q < sq.count; ++q, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * Cfg::ACC_COLS;
                const uint32_t tmem_x = Cfg::XACC ? tmem_d + BN : tmem_d;
                const int h0 = (sq.th + q * sq.dth) * p.bh;
                uint32_t started = 0;   // 0 until the first K-step of this tile has been issued (it overwrites the accumulator)
                const int per_row = p.kw * p.cin_chunks, n_rows = nK / per_row;   // K-steps per filter row (no divisions in the loop)
                for (int kh = 0; kh < n_rows; ++kh) {
                    // (same predicate as the producer: the K-steps of a filter row inside the zero padding were never loaded)
                    if (p.skip_pad_rows && (unsigned)(h0 + kh * p.dh - p.ph) >= (unsigned)p.hin) continue;
                for (int kj = 0; kj < per_row; ++kj) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::PLANES * TC_A_BYTES;
                    if (elect_one()) {
                        // (descriptors of the four k16 slices / the lo planes are the base descriptor plus a constant: the start-address field
                        //  counts 16-byte units and shared-memory addresses stay far below its 14 bits -- one shift + mask per K-step, not per MMA)
                        const uint64_t a0 = make_smem_desc(sa), b0 = make_smem_desc(sb);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {
                            const uint64_t a_hi = a0 + 2 * k, b_hi = b0 + 2 * k;
                            const uint32_t acc = (started | k) ? 1u : 0u;
                            if constexpr (PAIR) {
                                umma_f16_pair(tmem_d, a_hi, b_hi, idesc, acc);
                                if (SPLIT) {
                                    umma_f16_pair(tmem_x, a_hi, b_hi + (Cfg::B_BYTES >> 4), idesc, Cfg::XACC ? acc : 1u);
                                    umma_f16_pair(tmem_x, a_hi + (TC_A_BYTES >> 4), b_hi, idesc, 1u);
                                }
                                continue;
                            }
                            if (SPLIT && fuse) {
                                umma_f16(tmem_d, a_hi, b_hi, idesc2, acc);
                                umma_f16(tmem_x, a_hi + (TC_A_BYTES >> 4), b_hi, idesc, 1u);
                                continue;
                            }
                            umma_f16(tmem_d, a_hi, b_hi, idesc, acc);
                            if (SPLIT) {
                                umma_f16(tmem_x, a_hi, b_hi + (Cfg::B_BYTES >> 4), idesc, Cfg::XACC ? acc : 1u);
                                umma_f16(tmem_x, a_hi + (TC_A_BYTES >> 4), b_hi, idesc, 1u);
                            }
                        }
                        if constexpr (PAIR) umma_commit_pair(bar_empty + 8 * stage);   // frees the slot in BOTH CTAs
                        else umma_commit(bar_empty + 8 * stage);  // frees the smem slot once these MMAs have read it
                    }
                    __syncwarp();
                    started = 1;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                }
                if (elect_one()) { if constexpr (PAIR) umma_commit_pair(bar_tfull + 8 * as); else umma_commit(bar_tfull + 8 * as); }  // accumulator complete (tracks every MMA issued above)
                __syncwarp();
            }
            }
        }
    } else if (warp >= 4) {
      if constexpr (EPI == 1) {
        // ================================ epilogue: fused Conf_Fusion ================================
        // Columns [0, BN/2) of the tile are conf_gen channels nb*BN/2 + j, columns [BN/2, BN) the same channels of value_gen.  Group g
        // of 4 warps owns conf / value columns [32g, 32g + 32); thread = (tile row = pixel, group).  Per memory map q:
        //   e = exp(min(max(relu(conf), -6), 4)),  den += e,  num = fma(e, relu(value), num)      (connect.py:128-142)
        // and after the last map of the sample  out = num / den  goes to the fused map (split-fp16 planes and / or fp32).
        static_assert(EPI == 0 || BN == 128, "fused Conf_Fusion epilogue is written for the 64 + 64 column tile");
        constexpr int HALF = BN / 2;
        const int ew = warp & 3, eg = (warp - 4) >> 2, row = ew * 32 + lane, eall = threadIdx.x - 128;
        float num[32], den[32];
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int nb = tile % p.n_tiles_n;
            int mt = tile / p.n_tiles_n;
            const int tw = mt % p.tiles_w; mt /= p.tiles_w;
            const int th = mt % p.tiles_h;
            const int img0 = mt / p.tiles_h;
            const int hl = row / p.bw, wl = row - hl * p.bw;
            const int oh = th * p.bh + hl, ow = tw * p.bw + wl;
            const bool valid = hl < p.bh && oh < p.ho && ow < p.wo;
            for (int i = eall; i < BN; i += 128 * TC_EPI_GROUPS) { s_scale[i] = __ldg(p.scale + nb * BN + i); s_shift[i] = __ldg(p.shift + nb * BN + i); }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_GROUPS) : "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) { num[j] = 0.f; den[j] = 0.f; }
#pragma unroll 1
            for (int q = 0; q < p.group; ++q, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(bar_tfull + 8 * as, aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * Cfg::ACC_COLS;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = eg * 32 + h * 16;   // conf column; the matching value column is HALF + c
                    uint32_t u[16], w[16];
                    tmem_ld16(taddr + c, u);
                    tmem_ld16(taddr + HALF + c, w);
                    if (Cfg::XACC) {
                        uint32_t xu[16], xw[16];
                        tmem_ld16(taddr + BN + c, xu);
                        tmem_ld16(taddr + BN + HALF + c, xw);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            u[j] = __float_as_uint(__uint_as_float(u[j]) + __uint_as_float(xu[j]));
                            w[j] = __float_as_uint(__uint_as_float(w[j]) + __uint_as_float(xw[j]));
                        }
                    } else {
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float cf = fmaf(__uint_as_float(u[j]), s_scale[c + j], s_shift[c + j]);
                        float v = fmaf(__uint_as_float(w[j]), s_scale[HALF + c + j], s_shift[HALF + c + j]);
                        if (p.relu) { cf = fmaxf(cf, 0.f); v = fmaxf(v, 0.f); }
                        const float e = expf(fminf(fmaxf(cf, -6.f), 4.f));
                        den[h * 16 + j] += e;
                        num[h * 16 + j] = fmaf(e, v, num[h * 16 + j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
            }
            if (valid) {
                const size_t off = (((size_t)img0 * p.ho + oh) * p.wo + ow) * p.fuse_cout + nb * HALF + eg * 32;
                float y[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = num[j] / den[j];
                if (p.out_hi) {
                    uint4* oh4 = reinterpret_cast<uint4*>(p.out_hi + off);
                    uint4* ol4 = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        uint4 a, b;
                        __half2* ah = reinterpret_cast<__half2*>(&a);
                        __half2* bl = reinterpret_cast<__half2*>(&b);
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const float f0 = y[g4 * 8 + 2 * e2], f1 = y[g4 * 8 + 2 * e2 + 1];
                            const __half2 hh = __floats2half2_rn(f0, f1);
                            const float2 hf = __half22float2(hh);
                            ah[e2] = hh;
                            bl[e2] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
                        }
                        oh4[g4] = a;
                        if (SPLIT) ol4[g4] = b;
                    }
                }
                if (p.out_f32) {
                    float4* o = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) o[g4] = make_float4(y[4 * g4], y[4 * g4 + 1], y[4 * g4 + 2], y[4 * g4 + 3]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_GROUPS) : "memory");  // scale/shift staging may be overwritten next iteration
        }
      } else if constexpr (EPI == 2) {
        // ================================ epilogue: fused stem + max-pool ================================
        // thread = (pixel = TMEM lane, group g = channels [32g, 32g + 32)).  Vertical 3-tap maximum: a running maximum in registers.
        // Horizontal 3-tap / stride-2 maximum: warp shuffles (the even pixel 2px collects 2px - 1 and 2px + 1); the one neighbour that
        // lives in another warp (pixel 32w - 1, lane 31 of warp w - 1) crosses through a 128-byte shared-memory slot per (group, warp).
        static_assert(EPI != 2 || BN == 64, "the stem has 64 output channels");
        const int ew = warp & 3, eg = (warp - 4) >> 2, row = ew * 32 + lane, eall = threadIdx.x - 128;
        for (int i = eall; i < BN; i += 128 * TC_EPI_GROUPS) { s_scale[i] = __ldg(p.scale + i); s_shift[i] = __ldg(p.shift + i); }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_GROUPS) : "memory");
        float* edge = reinterpret_cast<float*>(s_out);   // [group][warp 0..2][32 channels]
        float vmax[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) vmax[j] = -FLT_MAX;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileSeq sq = decode_seq(p, tile);
            const int count = sq.count, th0 = sq.th, img = sq.img;
            for (int q = 0; q < count; ++q, ++it) {
                const int oy = th0 + q;
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(bar_tfull + 8 * as, aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * Cfg::ACC_COLS + eg * 32;
                uint32_t v[32];
                tmem_ld32(taddr, v);
                if (Cfg::XACC) {
                    uint32_t x[32];
                    tmem_ld32(taddr + BN, x);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(x[j]));
                } else {
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * as);   // the accumulator is in registers: the next row's MMAs may start
                const bool first = q == 0, odd = (oy & 1) != 0;
                const bool emit = (odd && !first) || (!odd && oy == p.ho - 1);   // conv row 2py+1 (or the last row of the map) completes pooled row py
                float y[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    y[j] = fmaf(__uint_as_float(v[j]), s_scale[eg * 32 + j], s_shift[eg * 32 + j]);
                    if (p.relu) y[j] = fmaxf(y[j], 0.f);
                    vmax[j] = first ? fmaxf(-FLT_MAX, y[j]) : fmaxf(vmax[j], y[j]);
                }
                if (emit) {   // (uniform over the CTA)
                    const int py = oy >> 1;   // odd row 2py+1 -> py ; even last row 2py -> py
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");   // the previous emit's edge values have been read
                    if (lane == 31 && ew < 3) {
                        float4* e4 = reinterpret_cast<float4*>(edge + (eg * 3 + ew) * 32);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) e4[g4] = make_float4(vmax[4 * g4], vmax[4 * g4 + 1], vmax[4 * g4 + 2], vmax[4 * g4 + 3]);
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");
                    const bool writer = (row & 1) == 0 && row < p.wo;   // even pixel 2px writes pooled pixel px
                    const bool has_right = row + 1 < p.wo;
                    const float* left_edge = edge + (eg * 3 + (ew > 0 ? ew - 1 : 0)) * 32;
                    const size_t o8 = ((((size_t)img * p.pool_po + py) * p.pool_po + (row >> 1)) * 64 + eg * 32) / 8;   // in uint4 (8 halves)
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        float m[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int j = g4 * 8 + e;
                            float left = __shfl_up_sync(0xffffffffu, vmax[j], 1);
                            float right = __shfl_down_sync(0xffffffffu, vmax[j], 1);
                            if (lane == 0) left = ew > 0 ? left_edge[j] : -FLT_MAX;
                            if (!has_right) right = -FLT_MAX;
                            m[e] = fmaxf(fmaxf(fmaxf(-FLT_MAX, left), vmax[j]), right);
                        }
                        if (writer) {
                            uint4 a, c;
                            __half2* ah = reinterpret_cast<__half2*>(&a);
                            __half2* cl = reinterpret_cast<__half2*>(&c);
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                const __half2 hh = __floats2half2_rn(m[2 * e2], m[2 * e2 + 1]);
                                const float2 hf = __half22float2(hh);
                                ah[e2] = hh;
                                cl[e2] = __floats2half2_rn(m[2 * e2] - hf.x, m[2 * e2 + 1] - hf.y);
                            }
                            reinterpret_cast<uint4*>(p.out_hi)[o8 + g4] = a;
                            if (SPLIT) reinterpret_cast<uint4*>(p.out_lo)[o8 + g4] = c;
                        }
                    }
                    if (odd) {   // conv row 2py+1 is also the first row of pooled row py+1
#pragma unroll
                        for (int j = 0; j < 32; ++j) vmax[j] = fmaxf(-FLT_MAX, y[j]);
                    }
                }
            }
        }
      } else {
        // ================================ epilogue ================================
        // Two groups of 4 warps; warp w reads TMEM lane quarter w % 4; group g owns the 32-column chunks with (chunk % 2 == g)
        // and one store-staging buffer, so the groups never synchronise with each other inside a tile.
        const int ew = warp & 3;              // TMEM lane quarter == warp % 4
        const int eg = (warp - 4) >> 2;       // epilogue group
        const int row = ew * 32 + lane;       // tile row == TMEM lane
        const int et = (threadIdx.x - 128) & 127;  // 0..127 inside the group
        const int eall = threadIdx.x - 128;        // 0..255 over both groups
        const uint32_t s_out_u32 = smem_u32(s_out);
        uint32_t gc = 0;                      // running chunk counter: selects the store-staging buffer
        // (single-fp16 mode has no lo planes: nothing is loaded, staged or stored for them)
        if (eall == 0 && p.tma_store) { tma_prefetch_desc(&p.o[0]); if (SPLIT) tma_prefetch_desc(&p.o[1]); }
        // (pair kernels never take the TMA-residual path -- the launch rule keeps residual layers on the one-CTA kernel, a forced pair reads its
        //  residual per thread -- so that machinery, and its registers, are compiled out of them)
#define tma_res_k (!PAIR && p.tma_res)
        if (eall == 32 && tma_res_k) { tma_prefetch_desc(&p.r[0]); if (SPLIT) tma_prefetch_desc(&p.r[1]); }
        // TMA-prefetched residual: with NB = p.nbuf staging buffers per group and a look-ahead of D = p.res_ahead chunks, the group's
        // leader requests chunk j+D into buffer (j+D) % NB when chunk j starts; that buffer was last used by chunk j+D-NB, whose bulk
        // store has finished reading shared memory once at most NB-D-1 stores are still pending (wait_group.read NB-D-1).
        // (NB = 3, D = 1 is the default.  D = NB-1 -- wait for the store issued one chunk ago instead, double the distance -- was measured
        //  to change nothing on the 1x1 + residual layers: they are not bound by the residual's latency.  Kept as tunable tc_res_ahead = 2.)
        const int NB = p.nbuf, RD = p.res_ahead;
        // (the chunk to request next and the chunk to consume next are tracked incrementally: no division by a run-time value sits on the
        //  issuing lane's path -- a first version that computed tile / column / buffer from the chunk index cost the residual layers 8-14 %)
        int rq_tile = t_first, rq_c0 = eg * 32;   // (pair mode: counted in pair tiles like the tile loop)
        uint32_t rq_buf = 0, use_buf = 0, use_phase = 0;
        // (the tile coordinates of the request are decoded once per tile, AFTER the request for the previous tile's last chunk has gone out --
        //  the four divisions used to sit in front of every chunk's TMA request on the warp the rest of the group waits for at the chunk barrier)
        int rq_n0 = 0, rq_x0 = 0, rq_y0 = 0, rq_i0 = 0;
        auto rq_decode = [&]() {
            const int t = tile_of(rq_tile);
            int mt_ = t / p.n_tiles_n;
            rq_n0 = (t - mt_ * p.n_tiles_n) * BN;
            const int tw_ = mt_ % p.tiles_w; mt_ /= p.tiles_w;
            rq_x0 = tw_ * p.bw;
            const int im_ = mt_ / p.tiles_h;
            rq_y0 = (mt_ - im_ * p.tiles_h) * p.bh;
            rq_i0 = im_ * p.bimg;
        };
        auto issue_res = [&](int c0, uint32_t b) {   // (called by the elected lane)
            const uint32_t dst = s_out_u32 + (eg * p.nbuf + b) * Cfg::BUF_BYTES, rb = bar_res + 8 * (eg * TC_RES_BUFS + b);
            mbar_expect_tx(rb, (SPLIT ? 2u : 1u) * p.bw * p.bh * p.bimg * 64u);
            tma_load_4d(dst, &p.r[0], rb, rq_n0 + c0, rq_x0, rq_y0, rq_i0);
            if (SPLIT) tma_load_4d(dst + TC_BM * 64, &p.r[1], rb, rq_n0 + c0, rq_x0, rq_y0, rq_i0);
        };
        // (single-thread TMA work of a group: one ELECTED lane of its first warp -- elect.sync picks the same lane every time, so the bulk
        //  async-groups it commits are the ones it later waits for; see elect_one() in tc_ptx.cuh for why not `if (et == 0)`)
        // every lane of the group's first warp keeps the same request state; only the elected lane issues
        auto rq_advance = [&]() {
            rq_c0 += 32 * TC_EPI_GROUPS;
            if (rq_c0 >= BN) { rq_c0 = eg * 32; rq_tile += t_step; if (rq_tile < t_count) rq_decode(); }
            if (++rq_buf == (uint32_t)NB) rq_buf = 0;
        };
        if (tma_res_k && et < 32 && eg * 32 < BN) {
            if (rq_tile < t_count) rq_decode();
            for (int k = 0; k < RD; ++k) {
                if (rq_tile < t_count && elect_one()) issue_res(rq_c0, rq_buf);
                __syncwarp();
                rq_advance();
            }
        }
        int it = 0;
        const uint32_t bar_tempty_arrive = PAIR ? mapa_cluster(bar_tempty, 0) : bar_tempty;   // (pair: the LEADER's MMA issuer waits for both epilogues)
        for (int tidx = t_first; tidx < t_count; tidx += t_step, ++it) {
            const int tile = tile_of(tidx);
            const int nb = tile % p.n_tiles_n;
            int mt = tile / p.n_tiles_n;
            const int tw = mt % p.tiles_w; mt /= p.tiles_w;
            const int th = mt % p.tiles_h;
            const int img = (mt / p.tiles_h) * p.bimg;   // first image of the tile: tile row = (image il, patch row hl, column wl)
            const int il = row / (p.bw * p.bh), r2 = row - il * p.bw * p.bh;
            const int hl = r2 / p.bw, wl = r2 - hl * p.bw;
            const int oh = th * p.bh + hl, ow = tw * p.bw + wl;
            const bool valid = il < p.bimg && img + il < p.n_img && oh < p.ho && ow < p.wo;
            const size_t pix = ((size_t)(img + il) * p.ho + oh) * p.wo + ow;
            const int n0 = nb * BN;
            // stage this tile's scale/shift (previous tile's readers are past the barrier at the end of the loop body)
            for (int i = eall; i < BN; i += 128 * TC_EPI_GROUPS) { s_scale[i] = __ldg(p.scale + n0 + i); s_shift[i] = __ldg(p.shift + n0 + i); }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_GROUPS) : "memory");
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const size_t off0 = pix * p.cout + n0;
            const int cfirst = eg * 32, cstep = 32 * TC_EPI_GROUPS;
            // Residual loads are software-pipelined one 32-column chunk ahead and the first chunk is requested BEFORE waiting
            // for the accumulator, so their DRAM latency hides behind the MMAs / the previous chunk's math.
            uint4 rh[4], rl[4];
            const bool has_res = p.res_hi != nullptr && valid && !tma_res_k;
            if (has_res) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    rh[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + off0 + cfirst) + q);
                    rl[q] = SPLIT ? __ldg(reinterpret_cast<const uint4*>(p.res_lo + off0 + cfirst) + q) : make_uint4(0, 0, 0, 0);
                }
            }
            mbar_wait(bar_tfull + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * Cfg::ACC_COLS;
#pragma unroll 1
            for (int c0 = cfirst; c0 < BN; c0 += cstep) {
                if (tma_res_k && et < 32) {
                    int nc0 = c0 + cstep, ntile = tile;
                    const bool wrap = nc0 >= BN;
                    if (wrap) { nc0 = cfirst; ntile = tile + gridDim.x; }   // (l2_prefetch is never set in pair mode)
                    if (elect_one()) {
                        const int pending = NB - RD - 1;   // stores that may still be reading their staging buffer
                        if (pending <= 0) bulk_wait_read<0>(); else if (pending == 1) bulk_wait_read<1>(); else bulk_wait_read<2>();
                        if (rq_tile < t_count) issue_res(rq_c0, rq_buf);
                        if (wrap && p.l2_prefetch && ntile < p.num_tiles) {
                            // the first chunk of the next tile is on its way; its REMAINING chunks start their trip from HBM to L2
                            // now (behind that demand load), so the per-chunk loads one chunk ahead no longer pay DRAM latency each
                            const int nb2 = ntile % p.n_tiles_n;
                            int mt2 = ntile / p.n_tiles_n;
                            const int tw2 = mt2 % p.tiles_w; mt2 /= p.tiles_w;
                            const int th2 = mt2 % p.tiles_h;
                            const int img2 = (mt2 / p.tiles_h) * p.bimg;
                            for (int c2 = cfirst + cstep; c2 < BN; c2 += cstep) {
                                tma_prefetch_l2_4d(&p.r[0], nb2 * BN + c2, tw2 * p.bw, th2 * p.bh, img2);
                                if (SPLIT) tma_prefetch_l2_4d(&p.r[1], nb2 * BN + c2, tw2 * p.bw, th2 * p.bh, img2);
                            }
                        }
                    }
                    __syncwarp();
                    rq_advance();
                }
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                uint4 ch[4], cl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { ch[q] = rh[q]; cl[q] = rl[q]; }
                if (has_res && c0 + cstep < BN) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        rh[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + off0 + c0 + cstep) + q);
                        rl[q] = SPLIT ? __ldg(reinterpret_cast<const uint4*>(p.res_lo + off0 + c0 + cstep) + q) : make_uint4(0, 0, 0, 0);
                    }
                }
                if (Cfg::XACC) {
                    uint32_t x[32];
                    tmem_ld32(taddr + BN + c0, x);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(x[j]));
                } else {
                    tmem_ld_wait();
                }
                if (valid || p.tma_store) {
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) y[j] = fmaf(__uint_as_float(v[j]), s_scale[c0 + j], s_shift[c0 + j]);
                    const size_t off = off0 + c0;
                    if (has_res) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __half2* ah = reinterpret_cast<const __half2*>(&ch[q]);
                            const __half2* bl = reinterpret_cast<const __half2*>(&cl[q]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 fa = __half22float2(ah[e]), fb = __half22float2(bl[e]);
                                y[q * 8 + 2 * e] += fa.x + fb.x;
                                y[q * 8 + 2 * e + 1] += fa.y + fb.y;
                            }
                        }
                    }
                    if (p.relu && !tma_res_k) {  // (with a TMA residual the add + ReLU happen after the chunk has landed, below)
#pragma unroll
                        for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
                    }
                    if (p.tma_store) {
                        // Stage this 32-column chunk (128 rows x 64 B per plane) in the 64B-swizzled layout of the output tensor
                        // map and let ONE thread issue the two bulk tensor stores: no per-thread global stores, rows beyond the
                        // image / patch are clipped by the TMA unit.  Two staging buffers alternate; a buffer is reused only after
                        // the stores issued from it have finished reading shared memory.
                        const uint32_t buf = eg * p.nbuf + (tma_res_k ? use_buf : 0);
                        uint8_t* rp = s_out + buf * Cfg::BUF_BYTES + row * 64;
                        const int sw = (row >> 1) & 3;
                        if (p.tma_f32) {
                            // fp32-only outputs (encoders -> xcorr, last tower conv -> pred): same staging buffer, rows of 128 B
                            // (32 floats) in the 128B-swizzled layout of the fp32 output map, ONE bulk tensor store per chunk
                            if (et < 32) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
                            asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");
                            uint8_t* rp32 = s_out + buf * Cfg::BUF_BYTES + row * 128;
                            const int sw7 = row & 7;
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                *reinterpret_cast<float4*>(rp32 + ((q ^ sw7) << 4)) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
                            fence_proxy_async_smem();
                            asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");
                            if (et < 32) {
                                if (elect_one()) {
                                    tma_store_4d(&p.o[0], s_out_u32 + buf * Cfg::BUF_BYTES, n0 + c0, tw * p.bw, th * p.bh, img);
                                    bulk_commit();
                                }
                                __syncwarp();
                            }
                            ++gc;
                            continue;
                        }
                        if (tma_res_k) {
                            mbar_wait(bar_res + 8 * (eg * TC_RES_BUFS + use_buf), use_phase);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 a = *reinterpret_cast<const uint4*>(rp + ((q ^ sw) << 4));
                                const uint4 b = SPLIT ? *reinterpret_cast<const uint4*>(rp + TC_BM * 64 + ((q ^ sw) << 4)) : make_uint4(0, 0, 0, 0);
                                const __half2* ah = reinterpret_cast<const __half2*>(&a);
                                const __half2* bl = reinterpret_cast<const __half2*>(&b);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 fa = __half22float2(ah[e]), fb = __half22float2(bl[e]);
                                    y[q * 8 + 2 * e] += fa.x + fb.x;
                                    y[q * 8 + 2 * e + 1] += fa.y + fb.y;
                                }
                            }
                            if (p.relu) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
                            }
                        } else {
                            if (et < 32) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
                            asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 a, b;
                            __half2* ah = reinterpret_cast<__half2*>(&a);
                            __half2* bl = reinterpret_cast<__half2*>(&b);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float f0 = y[q * 8 + 2 * e], f1 = y[q * 8 + 2 * e + 1];
                                const __half2 h = __floats2half2_rn(f0, f1);
                                const float2 hf = __half22float2(h);
                                ah[e] = h;
                                bl[e] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
                            }
                            *reinterpret_cast<uint4*>(rp + ((q ^ sw) << 4)) = a;
                            if (SPLIT) *reinterpret_cast<uint4*>(rp + TC_BM * 64 + ((q ^ sw) << 4)) = b;
                        }
                        fence_proxy_async_smem();
                        asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");
                        if (et < 32) {
                            const uint32_t src = s_out_u32 + buf * Cfg::BUF_BYTES;
                            if (elect_one()) {
                                tma_store_4d(&p.o[0], src, n0 + c0, tw * p.bw, th * p.bh, img);
                                if (SPLIT) tma_store_4d(&p.o[1], src + TC_BM * 64, n0 + c0, tw * p.bw, th * p.bh, img);
                                bulk_commit();
                            }
                            __syncwarp();
                        }
                        ++gc;
                        if (tma_res_k && ++use_buf == (uint32_t)NB) { use_buf = 0; use_phase ^= 1; }
                    } else if (p.out_hi) {
                        uint4* oh4 = reinterpret_cast<uint4*>(p.out_hi + off);
                        uint4* ol4 = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 a, b;
                            __half2* ah = reinterpret_cast<__half2*>(&a);
                            __half2* bl = reinterpret_cast<__half2*>(&b);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float f0 = y[q * 8 + 2 * e], f1 = y[q * 8 + 2 * e + 1];
                                const __half2 h = __floats2half2_rn(f0, f1);
                                const float2 hf = __half22float2(h);
                                ah[e] = h;
                                bl[e] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
                            }
                            oh4[q] = a;
                            if (SPLIT) ol4[q] = b;
                        }
                    }
                    if (p.out_f32 && valid) {
                        float4* o = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
                        for (int q = 0; q < 8; ++q) o[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(bar_tempty_arrive + 8 * as); else mbar_arrive(bar_tempty + 8 * as); }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_GROUPS) : "memory");  // scale/shift staging may be overwritten next iteration
        }
      }
    }
    if (threadIdx.x >= 128 && ((threadIdx.x - 128) & 127) < 32 && p.tma_store) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }  // staging must outlive the stores' reads
    if constexpr (PAIR) { tc_fence_before(); cluster_sync_all(); }   // neither CTA exits (or frees TMEM) while the pair's MMAs / remote arrivals may still touch it
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}
#undef tma_res_k
#undef cta_rank
#undef t_first
#undef t_step
#undef t_count

// =============================================================================================
// Host side: tensor maps, tiling, launch
// =============================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_tmap(CUtensorMap* m, int dtype, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                const cuuint32_t* box, int swizzle) {
    EncodeTiledFn fn = encode_fn();
    USOT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, (CUtensorMapDataType)dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("usot_b200: cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return 1;
    }
    return 0;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box, (int)swizzle);
}


// choose (tiles_w, bw, bh, bimg) maximising the fraction of the 128 MMA rows that are real output pixels.  A tile is a (bh x bw) patch of
// bimg consecutive images (one 4-D TMA box {64 ch, bw, bh, bimg}).  With one image per tile a 31-wide map fills 124 of 128 rows and its
// last row tile is 3/4 full (93.8 % overall; the 29 x 29 encoder maps: 82 %); ONE row of FOUR images fills the same 124 rows with no ragged
// last tile (96.9 %; 29 x 29: 90.6 %).  Every output element keeps its own accumulation order, so results do not depend on the choice.
Tunable g_tc_res_ahead = 1;          // tunable "tc_res_ahead": 1 = residual chunks requested one chunk ahead (3 staging buffers); 2 = NB-1 chunks ahead (measured: no gain; same results)
Tunable g_tc_skip_pad_rows = 1;       // tunable "tc_skip_pad_rows": one-row tiles skip the K-steps of filter rows that lie in the zero padding (A/B switch; same results)
Tunable g_tc_multi_image_tiles = 1;   // tunable "tc_multi_image_tiles": 0 = one image per tile (A/B switch; same arithmetic)
static void choose_tiling(int n, int ho, int wo, bool multi, double skip_frac_one_row, int* tiles_w, int* bw, int* bh, int* bimg) {
    // skip_frac_one_row: fraction of the (row, filter-row) pairs that fall into the zero padding, which ONE-ROW tiles skip entirely
    double best = -1;
    for (int tw = 1; tw <= 8; ++tw)   // full-width patches first: a narrower patch (shorter TMA rows) must buy at least 2 % more useful rows
        for (int bi = 1; bi <= (multi ? 8 : 1) && bi <= n; ++bi) {
            const int w = (wo + tw - 1) / tw;
            if (w * bi > 128) continue;
            const int h = std::min(128 / (w * bi), ho);
            const int th = (ho + h - 1) / h;
            const long tiles = (long)tw * th * ((n + bi - 1) / bi);
            const double eff = (double)n * ho * wo / ((double)tiles * 128.0) / (h == 1 ? 1.0 - skip_frac_one_row : 1.0);   // useful rows per executed MMA row
            const double need = best < 0 ? 0 : (tw > *tiles_w ? 0.02 : 1e-3);   // (near-ties keep the earlier = simpler plan)
            if (eff > best + need) { best = eff; *tiles_w = tw; *bw = w; *bh = h; *bimg = bi; }
        }
}

template <int BN, bool SPLIT, int EPI = 0, int PAIR = 0>
static int launch_cfg(TcParams& p, int grid, cudaStream_t st) {
    using Cfg = TcCfg<BN, SPLIT, PAIR>;
    static_assert(Cfg::STAGES >= 2, "not enough shared memory for a 2-stage pipeline");
    static SmemAttrCache attr;
    if (int rc = attr.ensure(conv_tc_kernel<BN, SPLIT, EPI, PAIR>, 227 * 1024)) return rc;
    if (!Cfg::TMA_OUT || EPI != 0 || PAIR) { if (!Cfg::TMA_OUT || EPI != 0) p.tma_store = 0; p.tma_res = 0; }
    if (p.tma_res && Cfg::STAGES_RES < 2) p.tma_res = 0;
    // residual pipeline: 3 buffers / look-ahead 1 (tc_res_ahead = 1, the first version) or look-ahead NB-1 with as many buffers (<= 4) as
    // still leave a two-stage operand ring (split mode: 3 buffers; single-fp16, whose stages are smaller: 4)
    int res_bufs = 3;
    if (p.tma_res && g_tc_res_ahead >= 2) {
        if (Cfg::stages_with_bufs(4) >= 2) res_bufs = 4;
        p.res_ahead = res_bufs - 1;
    } else {
        p.res_ahead = 1;
    }
    p.nbuf = EPI == 1 ? 0 : (p.tma_res ? res_bufs : 1);   // (the fused Conf_Fusion epilogue stores from registers: no staging buffers; the
                                                   //  stem + max-pool epilogue uses the two 16 KiB buffers of nbuf = 1 as its row buffer)
    p.stages = p.tma_res ? std::min(Cfg::stages_with_bufs(p.nbuf), Cfg::STAGES) : Cfg::STAGES;
    int smem = Cfg::smem_bytes(p.stages, p.nbuf);
    if (EPI == 2) {   // rolling activation rows + resident filter + edge buffer (must match the kernel's carve-up)
        constexpr int ROLL_R = SPLIT ? 5 : 6;
        p.stages = ROLL_R;
        p.nbuf = 0;
        smem = ROLL_R * Cfg::PLANES * TC_A_BYTES + 4 * Cfg::PLANES * Cfg::B_BYTES + 768 + Cfg::MISC_BYTES;
    }
    USOT_REQUIRE(smem <= 227 * 1024, "conv_tc: shared memory plan exceeds 227 KiB");
    if (PAIR) {   // clusters of two CTAs = the two SMs of a TPC
        USOT_REQUIRE(grid % 2 == 0 && p.num_pair_tiles > 0 && !p.pdl, "conv_tc: malformed CTA-pair launch");
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cudaLaunchAttribute cattr;
        memset(&cattr, 0, sizeof(cattr));
        cattr.id = cudaLaunchAttributeClusterDimension;
        cattr.val.clusterDim.x = 2; cattr.val.clusterDim.y = 1; cattr.val.clusterDim.z = 1;
        cfg.attrs = &cattr;
        cfg.numAttrs = 1;
        USOT_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, SPLIT, EPI, PAIR>, p));
        return 0;
    }
    if (p.pdl) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cudaLaunchAttribute attr;
        memset(&attr, 0, sizeof(attr));
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        USOT_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, SPLIT, EPI, PAIR>, p));
        return 0;
    }
    conv_tc_kernel<BN, SPLIT, EPI, PAIR><<<grid, TC_THREADS, smem, st>>>(p);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

Tunable g_tc_bn_max = 256;        // tunable: largest N tile (usot_set_tunable("tc_bn_max", 64|128|256))
Tunable g_tc_tma_res = 1;         // 1: residual via TMA into the staging buffer (needs tc_tma_store); 0: per-thread ld.global
Tunable g_tc_tma_store = 1;       // 1: TMA-store epilogue for split-fp16 outputs; 0: per-thread st.global (A/B switch)
Tunable g_tc_pdl = 1;             // programmatic dependent launch of the conv kernels on small grids (prologue overlaps the previous layer's tail)
Tunable g_tc_latency_split = 1;   // small grids: halve the N tile until at least half of the SMs have a CTA (batch-1 latency; same arithmetic)
Tunable g_tc_l2_prefetch = 0;     // TMA L2-prefetch hints for the next tile's residual chunks / 1x1 activation boxes (off: measured slower)
Tunable g_tc_tma_f32 = 1;         // fp32-only outputs leave through smem staging + one bulk tensor store per chunk (needs tc_tma_store)
Tunable g_tc_fuse_cross = 1;       // split mode with two accumulators: a_hi*[w_hi|w_lo] as ONE N = 2*BN MMA (A/B switch; same arithmetic)
Tunable g_tc_split_bn_max = 128;  // split mode: N <= 128 keeps the separate cross-term accumulator (accuracy); 256 trades it for reuse
// tunable "tc_cta_pair": large grids run as clusters of two CTAs that execute every MMA together (cta_group::2, M = 256; see conv_tc_kernel).
//   bit 0: single-fp16 launches with the 256- / 128-wide N tile, bit 1: split (fp16x3) launches with the 128-wide N tile, bit 2: also the
//   layers where the pair does not pay (short-K 1x1, residual; used by the tests).  Same results bit for bit.
Tunable g_tc_cta_pair = 3;

// Can this device keep one CTA pair per SM pair resident?  (cudaOccupancyMaxActiveClusters of the split pair kernel with its full shared-memory
// footprint; a GPU whose floor-swept GPCs strand SMs, or a context that refuses cluster launches, keeps the one-CTA kernels.)
static bool pair_launch_supported(int num_sms) {
    static std::atomic<int> cached[64];   // 0 = unknown, 1 = yes, 2 = no; per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    int c = cached[dev].load();
    if (c == 0) {
        int clusters = 0;
        auto kern = conv_tc_kernel<128, true, 0, 1>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)(num_sms & ~1));
            cfg.blockDim = dim3(TC_THREADS);
            cfg.dynamicSmemBytes = (size_t)TcCfg<128, true, 1>::smem_bytes(TcCfg<128, true, 1>::STAGES, 1);
            cudaLaunchAttribute cattr;
            memset(&cattr, 0, sizeof(cattr));
            cattr.id = cudaLaunchAttributeClusterDimension;
            cattr.val.clusterDim.x = 2; cattr.val.clusterDim.y = 1; cattr.val.clusterDim.z = 1;
            cfg.attrs = &cattr;
            cfg.numAttrs = 1;
            e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
        }
        if (e != cudaSuccess) { cudaGetLastError(); clusters = 0; }
        c = clusters >= num_sms / 2 ? 1 : 2;
        if (getenv("USOT_B200_DEBUG")) fprintf(stderr, "usot_b200: device %d keeps %d CTA pairs resident (%d SMs): pair kernels %s\n", dev, clusters, num_sms, c == 1 ? "on" : "off");
        cached[dev].store(c);
    }
    return c == 1;
}

struct PlanKey {  // plain words only (no padding: the key is hashed and compared as raw bytes)
    const void* ptr[11];
    int geom[14];
    int K, relu, split, device, knobs[9], group;
};
struct Plan { TcParams p; int bn, grid; };
struct PlanKeyHash {
    size_t operator()(const PlanKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 0xcbf29ce484222325ull;
        for (size_t i = 0; i < sizeof(PlanKey) / 8; ++i) { h = (h ^ w[i]) * 0x100000001b3ull; h ^= h >> 29; }
        return (size_t)h;
    }
};
struct PlanKeyEq { bool operator()(const PlanKey& a, const PlanKey& b) const { return memcmp(&a, &b, sizeof(PlanKey)) == 0; } };
static_assert(sizeof(PlanKey) % 8 == 0, "PlanKey is hashed as 64-bit words");
static std::mutex g_plan_mu;
static std::unordered_map<PlanKey, Plan, PlanKeyHash, PlanKeyEq> g_plans;

// (The key holds every pointer AND every shape field, so an address reused by another tensor of a different shape cannot alias.)
static bool plan_lookup(const PlanKey& k, Plan* out) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(k);
    if (it == g_plans.end()) return false;
    *out = it->second;
    return true;
}
static void plan_store(const PlanKey& k, const Plan& pl) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (g_plans.size() >= 4096) g_plans.clear();  // bounded: a sweep over many shapes simply re-encodes
    g_plans[k] = pl;
}

static int launch_plan(Plan& pl, bool split, cudaStream_t st) {
    if (pl.p.pool_bands > 0) return split ? launch_cfg<64, true, 2>(pl.p, pl.grid, st) : launch_cfg<64, false, 2>(pl.p, pl.grid, st);
    if (pl.p.group > 1 || pl.p.fuse_cout) return split ? launch_cfg<128, true, 1>(pl.p, pl.grid, st) : launch_cfg<128, false, 1>(pl.p, pl.grid, st);
    if (pl.p.num_pair_tiles > 0) {
        if (split) return launch_cfg<128, true, 0, 1>(pl.p, pl.grid, st);
        if (pl.bn == 256) return launch_cfg<256, false, 0, 1>(pl.p, pl.grid, st);
        return launch_cfg<128, false, 0, 1>(pl.p, pl.grid, st);
    }
    if (split) {
        if (pl.bn == 256) return launch_cfg<256, true>(pl.p, pl.grid, st);
        if (pl.bn == 128) return launch_cfg<128, true>(pl.p, pl.grid, st);
        return launch_cfg<64, true>(pl.p, pl.grid, st);
    }
    if (pl.bn == 256) return launch_cfg<256, false>(pl.p, pl.grid, st);
    if (pl.bn == 128) return launch_cfg<128, false>(pl.p, pl.grid, st);
    return launch_cfg<64, false>(pl.p, pl.grid, st);
}

int launch_conv_tc(const TcTensor& in, const ConvGeom& g, const TcWeights& w, const TcEpilogue& ep, bool split, cudaStream_t st, int fuse_group) {
    USOT_REQUIRE(g.cin % TC_BK == 0, "conv_tc needs Cin % 64 == 0");
    USOT_REQUIRE(g.cout % 64 == 0, "conv_tc needs Cout % 64 == 0");
    USOT_REQUIRE(g.stride == 1 || g.stride == 2, "conv_tc supports stride 1 and 2");
    USOT_REQUIRE(in.hi && (!split || in.lo), "conv_tc: missing input plane");
    USOT_REQUIRE(w.hi && (!split || w.lo), "conv_tc: missing weight plane");
    USOT_REQUIRE(ep.out_hi || ep.out_f32, "conv_tc: no output requested");
    if ((size_t)g.n * g.ho * g.wo == 0) return 0;
    if (fuse_group > 0) {
        USOT_REQUIRE(g.cout % 128 == 0 && g.n % fuse_group == 0, "fused Conf_Fusion: cout must be a multiple of 128 and n a multiple of the group");
        USOT_REQUIRE(!ep.res_hi && g.stride == 1, "fused Conf_Fusion: no residual, stride 1");
    }

    // Launch-plan cache: the tensor maps (up to 14 cuTensorMapEncodeTiled calls), tiling and kernel variant of a launch depend only on
    // the pointers, the geometry and the knobs.  The engine's arena hands out the same addresses for the same call shape, so in steady
    // state every layer hits (the eager medium-batch path was host-bound on the encodes).
    PlanKey key;
    memset(&key, 0, sizeof(key));
    const void* kp[11] = {in.hi, in.lo, w.hi, w.lo, w.scale, ep.shift, ep.res_hi, ep.res_lo, ep.out_hi, ep.out_lo, ep.out_f32};
    const int kg[14] = {g.n, g.h, g.w, g.cin, g.cout, g.kh, g.kw, g.stride, g.ph, g.pw, g.dh, g.dw, g.ho, g.wo};
    memcpy(key.ptr, kp, sizeof(kp));
    memcpy(key.geom, kg, sizeof(kg));
    key.K = w.K; key.relu = ep.relu; key.split = split ? 1 : 0; key.group = fuse_group;
    key.knobs[0] = g_tc_bn_max; key.knobs[1] = g_tc_split_bn_max; key.knobs[2] = g_tc_tma_store; key.knobs[3] = g_tc_tma_res;
    key.knobs[4] = g_tc_fuse_cross; key.knobs[5] = g_tc_tma_f32; key.knobs[6] = g_tc_l2_prefetch | (g_tc_multi_image_tiles << 1) | (g_tc_skip_pad_rows << 2) | (g_tc_res_ahead << 3); key.knobs[7] = g_tc_latency_split;
    key.knobs[8] = g_tc_pdl | (g_tc_cta_pair << 1);
    USOT_CUDA_OK(cudaGetDevice(&key.device));
    {
        Plan hit;
        if (plan_lookup(key, &hit)) return launch_plan(hit, split, st);
    }

    TcParams p;
    memset(&p, 0, sizeof(p));
    int bn = 64;
    const int bn_cap = split ? std::min(g_tc_bn_max.load(), g_tc_split_bn_max.load()) : g_tc_bn_max.load();
    if (g.cout % 256 == 0 && bn_cap >= 256) bn = 256;
    else if (g.cout % 128 == 0 && bn_cap >= 128) bn = 128;
    p.bimg = 1;
    double skip_frac = 0;   // stride-1 layers with vertical padding: share of (output row, filter row) pairs whose input row lies outside the map
    const bool multi = g_tc_multi_image_tiles && fuse_group == 0;
    if (multi && g_tc_skip_pad_rows && g.stride == 1 && g.ph > 0 && g.kh > 1) {
        int out = 0;
        for (int oy = 0; oy < g.ho; ++oy)
            for (int k = 0; k < g.kh; ++k) out += (oy + k * g.dh - g.ph < 0 || oy + k * g.dh - g.ph >= g.h) ? 1 : 0;
        skip_frac = (double)out / ((double)g.ho * g.kh);
    }
    choose_tiling(g.n, g.ho, g.wo, multi, skip_frac, &p.tiles_w, &p.bw, &p.bh, &p.bimg);
    p.hin = g.h;
    p.skip_pad_rows = (skip_frac > 0 && p.bh == 1) ? 1 : 0;
    p.tiles_h = (g.ho + p.bh - 1) / p.bh;
    const int img_tiles = (g.n / (fuse_group > 0 ? fuse_group : 1) + p.bimg - 1) / p.bimg;   // tiles along the image axis
    const int num_sms = device_sm_count();
    // Latency mode (small batches): while the grid would leave at least half of the SMs idle, halve the N tile -- twice as many CTAs,
    // and every MMA of a tile's K loop is half as wide (half as long).  The accumulation order of each output element does not
    // change, so results stay bit-identical to the wide-tile launch of a large batch (tests: batch independence, graph replay).
    // (Not across the 256 -> 128 step of split mode, which would switch to the two-accumulator arithmetic.)
    if (fuse_group > 0) bn = 128;   // 64 conf + 64 value columns per tile
    else if (g_tc_latency_split) {
        const int m_tiles = img_tiles * p.tiles_h * p.tiles_w;
        while (bn > 64 && !(split && bn == 256) && m_tiles * (g.cout / bn) * 2 <= num_sms) bn /= 2;
    }
    p.n_img = g.n; p.ho = g.ho; p.wo = g.wo; p.cout = g.cout;
    p.n_tiles_n = g.cout / bn;
    p.group = fuse_group > 0 ? fuse_group : 1;
    p.fuse_cout = fuse_group > 0 ? g.cout / 2 : 0;
    p.num_tiles = img_tiles * p.tiles_h * p.tiles_w * p.n_tiles_n;   // fused launch: one tile index = the `group` maps of a sample
    // CTA pairs (cta_group::2): large grids only (every pair of SMs gets a pair tile; small grids keep the latency-mode plan), ordinary epilogue,
    // the N tiles whose pair kernels exist.  The pair = image groups 2*gp and 2*gp + 1 at the same (patch, N block).
    bool pair = false;
    if (fuse_group == 0) {
        const int want = split ? (g_tc_cta_pair & 2) : (g_tc_cta_pair & 1);
        const bool bn_ok = split ? bn == 128 : (bn == 256 || bn == 128);   // (64-wide N tiles as pairs: measured, no gain -- layer1's 3x3 469 k vs 456 k cycles)
        // Where it pays (ncu launch lists of one batch-256 step, profiles/r02c_pair_*): the MMA-bound layers -- every 3x3 and the K = 1024
        // 1x1 layers: tensor pipe 84-91 % -> 91-96 % in fp16x3, -3 ... -8 % time -- but NOT the short-K 1x1 (+ residual) layers, whose
        // pair runs 30-38 % SLOWER (also with a third ring stage bought with one staging buffer less: measured, no change; these layers are
        // epilogue-bound, and -- presumably the reason -- the two epilogues of a pair are coupled through the leader's accumulator hand-over).  Bit 2 of the tunable overrides the rule (tests).
        const int k_steps = g.kh * g.kw * (g.cin / TC_BK);
        const bool pays = split ? (k_steps >= 12 && !ep.res_hi) : (k_steps >= 16 && !ep.res_hi && bn == 256 && !(ep.out_hi && ep.out_f32));
        const int pair_tiles = ((img_tiles + 1) / 2) * p.tiles_h * p.tiles_w * p.n_tiles_n;
        // (single-fp16 mode gains least -- its 3x3 layers are not feed-bound after all: tensor pipe 66-72 % with or without the pair -- and
        //  at batch 64 the pairs measured 4 % slower: there only grids of at least eight waves run as pairs)
        const int min_tiles = (split || (g_tc_cta_pair & 4)) ? num_sms / 2 : 8 * (num_sms / 2);
        if (want && bn_ok && (pays || (g_tc_cta_pair & 4)) && img_tiles >= 2 && pair_tiles >= min_tiles && pair_launch_supported(num_sms)) { pair = true; p.num_pair_tiles = pair_tiles; }
    }
    p.taps = g.kh * g.kw; p.kw = g.kw; p.cin_chunks = g.cin / TC_BK;
    p.stride = g.stride; p.ph = g.ph; p.pw = g.pw; p.dh = g.dh; p.dw = g.dw;
    p.scale = w.scale; p.shift = ep.shift;
    p.res_hi = ep.res_hi; p.res_lo = ep.res_lo;
    p.out_hi = ep.out_hi; p.out_lo = ep.out_lo; p.out_f32 = ep.out_f32;
    p.relu = ep.relu;
    USOT_REQUIRE(!split || !p.res_hi || p.res_lo, "conv_tc: split-mode residual needs both planes");
    USOT_REQUIRE(!split || !p.out_hi || p.out_lo, "conv_tc: split-mode output needs both planes");

    // ---- activation maps: dims {C, W', H', N}, one per (plane, parity) ----
    const int planes = split ? 2 : 1;
    const int npar = g.stride == 2 ? 4 : 1;
    for (int pl = 0; pl < planes; ++pl) {
        const __half* base = pl == 0 ? in.hi : in.lo;
        for (int par = 0; par < npar; ++par) {
            const int py = par >> 1, px = par & 1;
            cuuint64_t dims[4], strides[3];
            cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bimg};
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
            if (int rc = encode_map(&p.a[pl][par], b, 4, dims, strides, box)) return rc;
        }
        // ---- weight map: dims {K, Cout} ----
        cuuint64_t wd[2] = {(cuuint64_t)w.K, (cuuint64_t)g.cout}, ws[1] = {(cuuint64_t)w.K * 2};
        cuuint32_t wb[2] = {(cuuint32_t)TC_BK, (cuuint32_t)(pair ? bn / 2 : bn)};   // (pair: each CTA loads its half of the rows)
        if (int rc = encode_map(&p.b[pl], pl == 0 ? w.hi : w.lo, 2, wd, ws, wb)) return rc;
    }

    p.fuse_cross = g_tc_fuse_cross;
    p.l2_prefetch = pair ? 0 : g_tc_l2_prefetch.load();
    p.pdl = 0;  // decided below, once the tile count is known
    p.tma_store = (g_tc_tma_store && p.out_hi && !(split && bn == 256) && fuse_group == 0) ? 1 : 0;
    p.tma_f32 = 0;
    if (fuse_group == 0 && g_tc_tma_store && g_tc_tma_f32 && !p.out_hi && p.out_f32 && !p.res_hi && !(split && bn == 256)) {
        // fp32-only output: one fp32 map, box {32 floats = 128 B, bw, bh, 1}, 128B swizzle; uses the split path's staging buffers
        cuuint64_t od[4] = {(cuuint64_t)g.cout, (cuuint64_t)g.wo, (cuuint64_t)g.ho, (cuuint64_t)g.n};
        cuuint64_t os[3] = {(cuuint64_t)g.cout * 4, (cuuint64_t)g.wo * g.cout * 4, (cuuint64_t)g.ho * g.wo * g.cout * 4};
        cuuint32_t ob[4] = {32, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bimg};
        if (int rc = encode_tmap(&p.o[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.out_f32, 4, od, os, ob, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
        p.o[1] = p.o[0];
        p.tma_store = 1;
        p.tma_f32 = 1;
    } else if (p.tma_store) {
        cuuint64_t od[4] = {(cuuint64_t)g.cout, (cuuint64_t)g.wo, (cuuint64_t)g.ho, (cuuint64_t)g.n};
        cuuint64_t os[3] = {(cuuint64_t)g.cout * 2, (cuuint64_t)g.wo * g.cout * 2, (cuuint64_t)g.ho * g.wo * g.cout * 2};
        cuuint32_t ob[4] = {32, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bimg};
        if (int rc = encode_map(&p.o[0], p.out_hi, 4, od, os, ob, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
        if (split) { if (int rc = encode_map(&p.o[1], p.out_lo, 4, od, os, ob, CU_TENSOR_MAP_SWIZZLE_64B)) return rc; }
        else p.o[1] = p.o[0];
        if (p.res_hi && g_tc_tma_res) {
            p.tma_res = 1;
            if (int rc = encode_map(&p.r[0], p.res_hi, 4, od, os, ob, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
            if (split) { if (int rc = encode_map(&p.r[1], p.res_lo, 4, od, os, ob, CU_TENSOR_MAP_SWIZZLE_64B)) return rc; }
            else p.r[1] = p.r[0];
        }
    }

    Plan plan;
    plan.p = p;
    plan.bn = bn;
    plan.grid = pair ? 2 * std::min(p.num_pair_tiles, num_sms / 2) : std::min(p.num_tiles, num_sms);
    // Programmatic dependent launch pays in the latency regime (every tile has its own SM, idle SMs host the successor's prologue):
    // batch 1: -7 % per track() call.  At batch 256 it measured -1.6 % (within clock noise, no possible gain): not used there.
    plan.p.pdl = (!pair && g_tc_pdl && p.num_tiles * p.group <= num_sms) ? 1 : 0;
    plan_store(key, plan);
    return launch_plan(plan, split, st);
}

// =============================================================================================
// Stem (conv1 7x7 / stride 2 / pad 0, Cin = 3) + max-pool as a TMA-fed implicit GEMM            lib/models/modules.py:70-74,138-141
// =============================================================================================
// Space-to-depth turns the stride-2 7x7 conv over 3 channels into a stride-1 4x4 conv over 12 channels (filter zero-padded to 8x8):
//     s2d[n][Y][X][c*4 + a*2 + b] = x[n][c][2Y + a][2X + b]          (16 channels per pixel, 12 used)
//     out[oy][ox][co] = sum_{ky,kx<4} sum_{ch<16} s2d[oy + ky][ox + kx][ch] * w2[co][ky][kx][ch],   w2 = w[co][c][2ky + a][2kx + b]
// For a fixed ky the 64 values (kx, ch) of output pixel ox are the 128 contiguous bytes that START at s2d pixel (oy + ky, ox): the A
// operand of K-chunk ky is a view of the s2d plane whose "pixel" stride is 32 bytes and whose row length is 128 bytes -- rows overlap.
// A tensor map with dims {64, HO, HP, n} and strides {32 B, WP*32 B, HP*WP*32 B} describes exactly that, so ONE TMA box {64, HO, 1, 1}
// per (ky, plane) lands the im2col rows of a whole output row in shared memory, already 128B-swizzled: no software im2col, no
// per-tap conversion (the image is split into fp16 hi / lo planes once, by stem_s2d_kernel).  K = 4 x 64 = 256 (147 useful).
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* __restrict__ x, int n, int S, int HP, uint4* __restrict__ hi, uint4* __restrict__ lo) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * HP * HP;
    if (idx >= total) return;
    const int X = (int)(idx % HP);
    const int Y = (int)((idx / HP) % HP);
    const size_t b = idx / ((size_t)HP * HP);
    float v[16];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
                const int yy = 2 * Y + a, xx = 2 * X + bb;
                v[c * 4 + a * 2 + bb] = (yy < S && xx < S) ? __ldg(x + ((b * 3 + c) * S + yy) * S + xx) : 0.f;
            }
#pragma unroll
    for (int j = 12; j < 16; ++j) v[j] = 0.f;
    uint4 h[2], l[2];
    __half2* hh = reinterpret_cast<__half2*>(h);
    __half2* ll = reinterpret_cast<__half2*>(l);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const __half2 t = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 tf = __half22float2(t);
        hh[e] = t;
        ll[e] = __floats2half2_rn(v[2 * e] - tf.x, v[2 * e + 1] - tf.y);
    }
    hi[2 * idx] = h[0]; hi[2 * idx + 1] = h[1];
    if (lo) { lo[2 * idx] = l[0]; lo[2 * idx + 1] = l[1]; }
}

// stem_w: [147][64] fp32 with k = (c*7 + kh)*7 + kw  ->  w_kn2: [256][64] fp32 with k2 = ky*64 + kx*16 + c*4 + a*2 + b (kh = 2ky+a, kw = 2kx+b)
__global__ void stem_s2d_weight_kernel(const float* __restrict__ stem_w, float* __restrict__ w_kn2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 256 * 64) return;
    const int co = i & 63, k2 = i >> 6;
    const int ky = k2 >> 6, kx = (k2 >> 4) & 3, ch = k2 & 15;
    const int c = ch >> 2, a = (ch >> 1) & 1, b = ch & 1;
    const int kh = 2 * ky + a, kw = 2 * kx + b;
    w_kn2[i] = (ch < 12 && kh < 7 && kw < 7) ? __ldg(stem_w + ((c * 7 + kh) * 7 + kw) * 64 + co) : 0.f;
}

int launch_stem_s2d_weights(const float* stem_w, const float* stem_scale, float* w_kn2_scratch, __half* w_hi, __half* w_lo, float* scale_out,
                            cudaStream_t st) {
    stem_s2d_weight_kernel<<<64, 256, 0, st>>>(stem_w, w_kn2_scratch);
    USOT_CUDA_OK(cudaGetLastError());
    return launch_pack_tc_weights(w_kn2_scratch, 256, 64, stem_scale, w_hi, w_lo, scale_out, st);
}

// Bands per image for the fused stem + max-pool launch (0: not applicable, the map is wider than one 128-pixel tile, S > 261).
// A band walks its 2*band_rows + 1 conv rows one after the other on ONE SM: pick the count that minimises waves x rows per band.
int stem_pool_bands(int n, int s) {
    const int HO = (s - 7) / 2 + 1, PO = (HO + 2 - 3) / 2 + 1;
    if (HO > 128 || HO < 3 || n <= 0) return 0;
    const int G = device_sm_count();
    long best = -1;
    int best_nb = 0;
    for (int nb = 1; nb <= PO && nb <= 64; ++nb) {
        const int band_rows = (PO + nb - 1) / nb;
        if ((nb - 1) * band_rows >= PO) continue;   // the last band would be empty
        const long cost = (((long)n * nb + G - 1) / G) * (2L * band_rows + 1);
        if (best < 0 || cost < best) { best = cost; best_nb = nb; }
    }
    return best_nb;
}

size_t stem_s2d_plane_elems(int n, int s) { const int HP = (s - 7) / 2 + 1 + 3; return (size_t)n * HP * HP * 16; }

int launch_stem_s2d_pool(const float* x, int n, int s, const __half* w_hi, const __half* w_lo, const float* scale_tc, const float* shift,
                         __half* s2d_hi, __half* s2d_lo, __half* pool_hi, __half* pool_lo, bool split, cudaStream_t st) {
    const int nb = stem_pool_bands(n, s);
    USOT_REQUIRE(nb > 0, "fused stem + max-pool is not applicable to this shape (ask stem_pool_bands first)");
    USOT_REQUIRE(s2d_hi && pool_hi && (!split || (s2d_lo && pool_lo && w_lo)), "fused stem + max-pool: missing plane");
    const int HO = (s - 7) / 2 + 1, HP = HO + 3, PO = (HO + 2 - 3) / 2 + 1;
    {   // image -> space-to-depth fp16 planes
        const size_t total = (size_t)n * HP * HP;
        stem_s2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, n, s, HP, reinterpret_cast<uint4*>(s2d_hi),
                                                                          split ? reinterpret_cast<uint4*>(s2d_lo) : nullptr);
        USOT_CUDA_OK(cudaGetLastError());
    }
    PlanKey key;
    memset(&key, 0, sizeof(key));
    const void* kp[11] = {s2d_hi, s2d_lo, w_hi, w_lo, scale_tc, shift, nullptr, nullptr, pool_hi, pool_lo, nullptr};
    const int kg[14] = {n, HP, HP, 16, 64, 4, 1, 1, 0, 0, 1, 1, HO, HO};
    memcpy(key.ptr, kp, sizeof(kp));
    memcpy(key.geom, kg, sizeof(kg));
    key.K = 256; key.relu = 1; key.split = split ? 1 : 0; key.group = 1000 + nb;
    key.knobs[4] = g_tc_fuse_cross;
    USOT_CUDA_OK(cudaGetDevice(&key.device));
    Plan plan;
    if (plan_lookup(key, &plan)) return launch_plan(plan, split, st);

    TcParams p;
    memset(&p, 0, sizeof(p));
    p.n_img = n; p.ho = HO; p.wo = HO; p.cout = 64;
    p.bw = HO; p.bh = 1; p.bimg = 1; p.tiles_w = 1; p.tiles_h = HO;
    p.n_tiles_n = 1; p.group = 1;
    p.pool_bands = nb; p.pool_band_rows = (PO + nb - 1) / nb; p.pool_po = PO;
    p.num_tiles = n * nb;
    p.taps = 4; p.kw = 1; p.cin_chunks = 1;
    p.stride = 1; p.ph = 0; p.pw = 0; p.dh = 1; p.dw = 1;
    p.scale = scale_tc; p.shift = shift;
    p.out_hi = pool_hi; p.out_lo = pool_lo;
    p.relu = 1;
    p.fuse_cross = g_tc_fuse_cross;
    for (int pl = 0; pl < (split ? 2 : 1); ++pl) {
        // overlapping-row view of the s2d plane: "pixel" ox = the 64 halves that start at s2d pixel ox (stride 16 halves = 32 B)
        const __half* base = pl == 0 ? s2d_hi : s2d_lo;
        cuuint64_t dims[4] = {64, (cuuint64_t)HO, (cuuint64_t)HP, (cuuint64_t)n};
        cuuint64_t strides[3] = {32, (cuuint64_t)HP * 32, (cuuint64_t)HP * HP * 32};
        cuuint32_t box[4] = {64, (cuuint32_t)HO, 1, 1};
        if (!p.a_rank5 && encode_map(&p.a[pl][0], base, 4, dims, strides, box) != 0) {
            USOT_REQUIRE(pl == 0, "stem: tensor-map encode failed for the lo plane only");
            p.a_rank5 = 1;   // the driver refuses rows that overlap: describe the same bytes as {16 ch, 4 kx, ox, row, image}
        }
        if (p.a_rank5) {
            cuuint64_t d5[5] = {16, 4, (cuuint64_t)HO, (cuuint64_t)HP, (cuuint64_t)n};
            cuuint64_t s5[4] = {32, 32, (cuuint64_t)HP * 32, (cuuint64_t)HP * HP * 32};
            cuuint32_t b5[5] = {16, 4, (cuuint32_t)HO, 1, 1};
            if (int rc = encode_map(&p.a[pl][0], base, 5, d5, s5, b5)) return rc;
        }
        cuuint64_t wd[2] = {256, 64}, ws[1] = {256 * 2};
        cuuint32_t wb[2] = {64, 64};
        if (int rc = encode_map(&p.b[pl], pl == 0 ? w_hi : w_lo, 2, wd, ws, wb)) return rc;
    }
    plan.p = p;
    plan.bn = 64;
    plan.grid = std::min(p.num_tiles, device_sm_count());
    plan_store(key, plan);
    return launch_plan(plan, split, st);
}

// =============================================================================================
// fp32 -> split-fp16 conversion (activations) and host-side weight packing
// =============================================================================================
__global__ void f32_to_split_kernel(const float4* __restrict__ in, size_t n4, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                    const float* __restrict__ mul /*device scalar (a power of two) or null*/) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = __ldg(in + i);
    if (mul) { const float f = __ldg(mul); v.x *= f; v.y *= f; v.z *= f; v.w *= f; }
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    uint2 a, b;
    a.x = *reinterpret_cast<const uint32_t*>(&h0); a.y = *reinterpret_cast<const uint32_t*>(&h1);
    b.x = *reinterpret_cast<const uint32_t*>(&l0); b.y = *reinterpret_cast<const uint32_t*>(&l1);
    hi[i] = a;
    if (lo) lo[i] = b;
}

int launch_f32_to_split(const float* in, size_t n, __half* hi, __half* lo, cudaStream_t st, const float* mul) {
    USOT_REQUIRE(n % 4 == 0, "f32_to_split: element count must be a multiple of 4");
    if (n == 0) return 0;
    const size_t n4 = n / 4;
    f32_to_split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in), n4, reinterpret_cast<uint2*>(hi),
                                                                       reinterpret_cast<uint2*>(lo), mul);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void split_to_f32_kernel(const __half2* __restrict__ hi, const __half2* __restrict__ lo, size_t n2, float2* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const float2 a = __half22float2(hi[i]);
    const float2 b = lo ? __half22float2(lo[i]) : make_float2(0.f, 0.f);
    out[i] = make_float2(a.x + b.x, a.y + b.y);
}

int launch_split_to_f32(const __half* hi, const __half* lo, size_t n, float* out, cudaStream_t st) {
    USOT_REQUIRE(n % 2 == 0, "split_to_f32: element count must be even");
    if (n == 0) return 0;
    split_to_f32_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __half2*>(hi), reinterpret_cast<const __half2*>(lo),
                                                                         n / 2, reinterpret_cast<float2*>(out));
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// Device-side twin of pack_tc_weights_host (bit-identical results): one block per output channel.  Used where the weights change
// between calls (training, stand-alone op): w_kn [K][cout] fp32 -> [cout][K] fp16 hi / lo planes of w * 2^e[co], scale_out = scale_in * 2^-e.
__global__ void __launch_bounds__(256) pack_tc_weights_kernel(const float* __restrict__ w_kn, int K, int cout, const float* __restrict__ scale_in,
                                                              __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ scale_out) {
    __shared__ float red[8];
    const int co = blockIdx.x;
    float mx = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, fabsf(__ldg(w_kn + (size_t)k * cout + co)));
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    int e = 0;
    if (mx > 0.f && isfinite(mx)) {
        int ex;
        frexpf(mx, &ex);
        e = 8 - ex;
    }
    const float s = ldexpf(1.0f, e);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float v = __ldg(w_kn + (size_t)k * cout + co) * s;
        const __half h = __float2half_rn(v);
        hi[(size_t)co * K + k] = h;
        if (lo) lo[(size_t)co * K + k] = __float2half_rn(v - __half2float(h));
    }
    if (threadIdx.x == 0) scale_out[co] = __ldg(scale_in + co) * ldexpf(1.0f, -e);
}

int launch_pack_tc_weights(const float* w_kn, int K, int cout, const float* scale_in, __half* hi, __half* lo, float* scale_out, cudaStream_t st) {
    pack_tc_weights_kernel<<<cout, 256, 0, st>>>(w_kn, K, cout, scale_in, hi, lo, scale_out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

// w_kn: [K][cout] fp32 (k = tap*cin + c).  Produces [cout][K] fp16 hi/lo planes of w * 2^e[co] and scale_out = scale_in * 2^-e.
void pack_tc_weights_host(const float* w_kn, int K, int cout, const float* scale_in, std::vector<__half>& hi, std::vector<__half>& lo,
                          std::vector<float>& scale_out) {
    hi.resize((size_t)cout * K);
    lo.resize((size_t)cout * K);
    scale_out.resize(cout);
    for (int co = 0; co < cout; ++co) {
        float mx = 0.f;
        for (int k = 0; k < K; ++k) mx = std::max(mx, std::fabs(w_kn[(size_t)k * cout + co]));
        int e = 0;
        if (mx > 0.f && std::isfinite(mx)) {
            int ex;
            std::frexp(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)
            e = 8 - ex;           // scaled max lands in [128, 256): hi and lo planes both far from fp16 under/overflow
        }
        const float s = std::ldexp(1.0f, e);
        for (int k = 0; k < K; ++k) {
            const float v = w_kn[(size_t)k * cout + co] * s;
            const __half h = __float2half_rn(v);
            hi[(size_t)co * K + k] = h;
            lo[(size_t)co * K + k] = __float2half_rn(v - __half2float(h));
        }
        scale_out[co] = scale_in[co] * std::ldexp(1.0f, -e);
    }
}

}  // namespace usot
