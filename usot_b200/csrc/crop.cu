// SiamFC-style crop + average-colour padding + bilinear resize on the device, bit-exact with the reference's host path:
//   get_subwindow_tracking   lib/utils/track_utils.py:30-119  (context window :41-55, uint8 canvas padded with the truncated
//                            channel means :58-70, crop :71, cv2.resize :77-78, HWC uint8 -> CHW float32 `im_to_torch` :24-27)
// cv2.resize(uint8, INTER_LINEAR) is OpenCV's 11-bit fixed-point bilinear (modules/imgproc/src/resize.cpp, HResizeLinear +
// VResizeLinear<uchar,int,short>); an exact 2x down-scale is routed to INTER_AREA there.  The per-index coefficients are
// recomputed here with non-contracting IEEE intrinsics in the same double/float sequence OpenCV uses, so every output byte
// equals the host result (tests/test_gpu_crop.py; oracle/crop_oracle.py is pinned against the live cv2 and the live reference).
// Integer/byte gather work: one thread per output pixel (3 channels), reads hit L2 (a 480x640 frame is 0.9 MB), the fp32 CHW
// patch is written in 128-byte row segments; batched over (frame, window) pairs (grid.z) so several videos / scales share one launch.
#include "common.cuh"

namespace usot {

static __device__ __forceinline__ void lin_coef(int d, int ssize, int dsize, bool vertical, int& s, int& c0, int& c1) {
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dsize, (double)ssize));
    float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
    s = __float2int_rd(f);
    f = __fsub_rn(f, (float)s);
    if (!vertical) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    }
    c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

struct Px { int v[3]; };

static __device__ __forceinline__ Px fetch(const uint8_t* __restrict__ frame, int H, int W, int y, int x, const uint8_t* __restrict__ fill) {
    Px p;
    if (y >= 0 && y < H && x >= 0 && x < W) {
        const uint8_t* q = frame + ((size_t)y * W + x) * 3;
        p.v[0] = q[0]; p.v[1] = q[1]; p.v[2] = q[2];
    } else {
        p.v[0] = fill[0]; p.v[1] = fill[1]; p.v[2] = fill[2];
    }
    return p;
}

// Block = 32 x 8 output pixels of one window.  The fixed-point coefficients depend only on the output column (horizontal) or row
// (vertical), so each block computes its 32 column and 8 row coefficient sets once into shared memory (40 double-precision
// divisions per 256 pixels instead of 512): the first version recomputed both per pixel and was instruction-issue bound (81 %).
__global__ void __launch_bounds__(256) crop_resize_kernel(const uint8_t* __restrict__ frames, int n_frames, int H, int W,
                                                          const int* __restrict__ crops, int4 one, const uint8_t* __restrict__ fills, int msz,
                                                          float* __restrict__ out) {
    __shared__ int s_col[32][3], s_row[8][3];
    const int i = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int dx = blockIdx.x * 32 + tx, dy = blockIdx.y * 8 + ty;
    // windows come from a device table, or (crops == nullptr, one window) by value so that a per-frame caller uploads nothing
    const int4 cw = crops ? make_int4(crops[i * 4 + 0], crops[i * 4 + 1], crops[i * 4 + 2], crops[i * 4 + 3]) : one;
    const int fi = min(max(cw.x, 0), n_frames - 1), xmin = cw.y, ymin = cw.z, osz = cw.w;
    const bool bilinear = osz != msz && osz != 2 * msz;
    if (bilinear) {
        if (ty == 0 && dx < msz) lin_coef(dx, osz, msz, false, s_col[tx][0], s_col[tx][1], s_col[tx][2]);
        if (ty == 1 && tx < 8 && blockIdx.y * 8 + tx < msz) lin_coef(blockIdx.y * 8 + tx, osz, msz, true, s_row[tx][0], s_row[tx][1], s_row[tx][2]);
        __syncthreads();
    }
    if (dx >= msz || dy >= msz) return;
    const int pix = dy * msz + dx;
    const uint8_t* frame = frames + (size_t)fi * H * W * 3;
    const uint8_t* fill = fills + i * 3;
    int r[3];
    if (osz == msz) {  // track_utils.py:77-80: no resize
        const Px p = fetch(frame, H, W, ymin + dy, xmin + dx, fill);
        r[0] = p.v[0]; r[1] = p.v[1]; r[2] = p.v[2];
    } else if (osz == 2 * msz) {  // resize.cpp: INTER_LINEAR with an exact 2x2 decimation runs the INTER_AREA fast path
        const Px a = fetch(frame, H, W, ymin + 2 * dy, xmin + 2 * dx, fill), b = fetch(frame, H, W, ymin + 2 * dy, xmin + 2 * dx + 1, fill);
        const Px c = fetch(frame, H, W, ymin + 2 * dy + 1, xmin + 2 * dx, fill), d = fetch(frame, H, W, ymin + 2 * dy + 1, xmin + 2 * dx + 1, fill);
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = (a.v[k] + b.v[k] + c.v[k] + d.v[k] + 2) >> 2;
    } else {
        const int sx = s_col[tx][0], a0 = s_col[tx][1], a1 = s_col[tx][2];
        const int sy = s_row[ty][0], b0 = s_row[ty][1], b1 = s_row[ty][2];
        const int x0 = sx, x1 = min(sx + 1, osz - 1);
        const int y0 = min(max(sy, 0), osz - 1), y1 = min(max(sy + 1, 0), osz - 1);
        const Px p00 = fetch(frame, H, W, ymin + y0, xmin + x0, fill), p01 = fetch(frame, H, W, ymin + y0, xmin + x1, fill);
        const Px p10 = fetch(frame, H, W, ymin + y1, xmin + x0, fill), p11 = fetch(frame, H, W, ymin + y1, xmin + x1, fill);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int h0 = p00.v[k] * a0 + p01.v[k] * a1, h1 = p10.v[k] * a0 + p11.v[k] * a1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            r[k] = min(max(v, 0), 255);
        }
    }
    const size_t plane = (size_t)msz * msz;
    float* o = out + (size_t)i * 3 * plane + pix;
    o[0] = (float)r[0]; o[plane] = (float)r[1]; o[2 * plane] = (float)r[2];
}

int launch_crop_resize(const uint8_t* frames, int n_frames, int H, int W, const int* crops, const uint8_t* fills, int n, int msz, float* out,
                       cudaStream_t st) {
    if (n == 0) return 0;
    dim3 grid((unsigned)((msz + 31) / 32), (unsigned)((msz + 7) / 8), (unsigned)n);
    crop_resize_kernel<<<grid, 256, 0, st>>>(frames, n_frames, H, W, crops, make_int4(0, 0, 0, 0), fills, msz, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_crop_resize_one(const uint8_t* frame, int H, int W, int xmin, int ymin, int osz, const uint8_t* fill3, int msz, float* out,
                           cudaStream_t st) {
    dim3 grid((unsigned)((msz + 31) / 32), (unsigned)((msz + 7) / 8), 1u);
    crop_resize_kernel<<<grid, 256, 0, st>>>(frame, 1, H, W, nullptr, make_int4(0, xmin, ymin, osz), fill3, msz, out);
    USOT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace usot
