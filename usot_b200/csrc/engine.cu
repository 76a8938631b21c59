// Engine + C ABI of the usot_b200 library.
//
// The engine owns the packed weights (BN folded on the host at finalize()) and a grow-only device arena for
// activations, and runs the reference's forward graphs as sequences of the kernels in kernels_simt.cu /
// conv_tc.cu on the caller's stream:
//   backbone_neck   ResNet_plus2.forward + AdjustLayer      lib/models/modules.py:137-151, connect.py:294-296
//   template        USOT_.template                          lib/models/models.py:173-177
//   track           USOT_.track -> box_tower_reg.forward    lib/models/models.py:179-198, connect.py:221-281
//   extract_memory_feature                                  lib/models/models.py:200-206
#include "common.cuh"
#include "conv_tc.cuh"
#include "../../include/usot_b200.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

namespace usot {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

// ---- optional per-kernel-family profiling (CUDA events on the launching stream) + launch counting ----
enum { FAM_CONV = 0, FAM_STEM, FAM_POOL, FAM_XCORR, FAM_PRED, FAM_FUSION, FAM_PRROI, FAM_OTHER, FAM_WGRAD, FAM_TRAIN, FAM_COUNT };
static const char* kFamNames[FAM_COUNT] = {"conv", "stem", "maxpool", "groupdw_xcorr", "pred_conv", "conf_fusion", "prroi_pool", "other",
                                           "conv_wgrad", "train_other"};
static_assert(FAM_COUNT <= 16, "launch-count arrays hold 16 families");
struct Profiler {
    bool on = false;
    long long launches[FAM_COUNT] = {0};
    double flops[FAM_COUNT] = {0};   // algorithmic FLOPs (dense convs) issued since reset
    double bytes[FAM_COUNT] = {0};   // algorithmic bytes (bandwidth kernels) issued since reset
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev;
};
static Profiler g_prof;
static thread_local long long tl_launches[16] = {0};  // this host thread's launches (graph capture takes its per-family counts from here)
static std::mutex g_prof_mu;  // launches of several engines (DataParallel-style host threads) update the counters concurrently
static void count_launches(int fam, long long n) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.launches[fam] += n;
}
// launch accounting for the operators that live in other translation units (ops_abi.cu, train_kernels.cu)
void count_op_launch(int family, int n) { if (family >= 0 && family < FAM_COUNT) count_launches(family, n); }
static Tunable g_conf_fusion_fused{1};  // tunable "conf_fusion_fused": 1 = conf_gen || value_gen as ONE conv with the Conf_Fusion reduction in its epilogue (tcgen05 modes)
static Tunable g_stem_pool_fused{1};  // tunable "stem_pool_fused": 1 = stem + max-pool as ONE TMA-fed implicit GEMM over the space-to-depth image (conv_tc.cu, EPI = 2)
static Tunable g_stem_tc{1};  // tunable "stem_tc": 1 = tensor-core stem in the tcgen05 precision modes, 0 = CUDA-core stem
struct Scope {
    int fam; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    Scope(int fam_, cudaStream_t st_, double flops = 0, double bytes = 0) : fam(fam_), st(st_) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.launches[fam]++;
        tl_launches[fam]++;
        g_prof.flops[fam] += flops;
        g_prof.bytes[fam] += bytes;
        if (g_prof.on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~Scope() {
        if (!a) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.ev.push_back({fam, {a, b}});
    }
};

// ---------------------------------------------------------------------------------------------
// Architecture table (restates the reference constructors, lib/models/modules.py:61-135, connect.py:160-219)
// ---------------------------------------------------------------------------------------------
struct ConvSpec {
    std::string name, bn;
    int cin, cout, k, stride, ph, pw, dh, dw;
    bool bias, relu;
};

static std::vector<ConvSpec> build_specs() {
    std::vector<ConvSpec> v;
    const std::string bk = "features.features.";
    auto add = [&](const std::string& name, const std::string& bn, int cin, int cout, int k, int stride, int ph, int pw, int dh,
                   int dw, bool bias, bool relu) { v.push_back({name, bn, cin, cout, k, stride, ph, pw, dh, dw, bias, relu}); };
    struct L { const char* name; int planes, blocks, stride, dil; };
    const L layers[3] = {{"layer1", 64, 3, 1, 1}, {"layer2", 128, 4, 2, 1}, {"layer3", 256, 6, 1, 2}};
    int inplanes = 64;
    for (const L& l : layers) {
        for (int i = 0; i < l.blocks; ++i) {
            std::string p = bk + l.name + "." + std::to_string(i) + ".";
            const bool down = i == 0;
            const int s = down ? l.stride : 1;
            int pad = 2 - s, dil = l.dil;  // Bottleneck.__init__, modules.py:18-29
            if (down && dil > 1) { dil /= 2; pad = dil; }
            if (dil > 1) pad = dil;
            add(p + "conv1", p + "bn1", inplanes, l.planes, 1, 1, 0, 0, 1, 1, false, true);
            add(p + "conv2", p + "bn2", l.planes, l.planes, 3, s, pad, pad, dil, dil, false, true);
            add(p + "conv3", p + "bn3", l.planes, l.planes * 4, 1, 1, 0, 0, 1, 1, false, true /* after the residual add */);
            if (down) {
                if (l.stride == 1 && l.dil == 1)  // modules.py:110-115
                    add(p + "downsample.0", p + "downsample.1", inplanes, l.planes * 4, 1, 1, 0, 0, 1, 1, false, false);
                else {  // modules.py:116-126 (3x3 shortcut; dilation is not passed to the conv)
                    const int dpad = l.dil > 1 ? l.dil / 2 : 0;
                    add(p + "downsample.0", p + "downsample.1", inplanes, l.planes * 4, 3, l.stride, dpad, dpad, 1, 1, false, false);
                }
                inplanes = l.planes * 4;
            }
        }
    }
    add("neck.downsample.0", "neck.downsample.1", 1024, 256, 1, 1, 0, 0, 1, 1, false, false);  // connect.py:287-290
    for (const char* enc : {"cls_encode", "reg_encode"})
        for (int m = 0; m < 3; ++m)
            for (const char* br : {"k", "s"}) {
                static const char* mn[3] = {"matrix11", "matrix12", "matrix21"};
                static const int dh[3] = {1, 2, 1}, dw[3] = {1, 1, 2};  // connect.py:20-53
                std::string p = std::string("connect_model.") + enc + "." + mn[m] + "_" + br + ".";
                add(p + "0", p + "1", 256, 256, 3, 1, 0, 0, dh[m], dw[m], false, true);
            }
    for (const char* g : {"conf_gen", "value_gen"}) {  // connect.py:112-121
        std::string p = std::string("connect_model.conf_fusion.") + g + ".";
        add(p + "0", p + "1", 256, 256, 3, 1, 1, 1, 1, 1, true, true);
    }
    for (const char* t : {"bbox_tower", "cls_tower", "cls_memory_tower"})  // connect.py:178-209
        for (int i = 0; i < 4; ++i) {
            std::string p = std::string("connect_model.") + t + ".";
            add(p + std::to_string(3 * i), p + std::to_string(3 * i + 1), 256, 256, 3, 1, 1, 1, 1, 1, true, true);
        }
    return v;
}

struct ConvW {
    ConvSpec s;
    float* w_kn = nullptr;   // [k*k*cin][cout] fp32 (SIMT path)
    float* scale = nullptr;  // folded BN
    float* shift = nullptr;
    // tcgen05 path: [cout][K] fp16 planes of w*2^e, scale_tc = scale*2^-e
    __half* w_hi = nullptr;
    __half* w_lo = nullptr;
    float* scale_tc = nullptr;
};

// One NHWC activation tensor; may exist as fp32, as split fp16 planes, or both.
struct T {
    float* f = nullptr;
    __half* hi = nullptr;
    __half* lo = nullptr;
    int n = 0, h = 0, w = 0, c = 0;
    size_t numel() const { return (size_t)n * h * w * c; }
};

struct PredW {
    float* w = nullptr;   // [9][cout][256]
    float* w4 = nullptr;  // [64][9*cout][4]: same weights, channel groups outermost (small-batch kernel)
    float* b = nullptr;
    int cout = 0;
};

struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    bool plan = false;  // planning pass: count bytes, launch nothing
    void* alloc(size_t bytes) {
        size_t a = (off + 255) & ~size_t(255);
        off = a + bytes;
        return plan ? reinterpret_cast<void*>(uintptr_t(0x1000) + a) : (void*)(base + a);
    }
    float* f(size_t n) { return static_cast<float*>(alloc(n * sizeof(float))); }
    __half* h(size_t n) { return static_cast<__half*>(alloc(n * sizeof(__half))); }
};

}  // namespace usot

using namespace usot;

struct usot_engine {
    int device = 0;
    int precision = USOT_PREC_FP32_SIMT;
    bool finalized = false;
    std::map<std::string, std::vector<float>> host;
    std::map<std::string, ConvW> convs;
    std::vector<void*> owned;  // device allocations holding weights
    int64_t weight_bytes = 0;
    // packed-weight image (usot_engine_export_packed / usot_engine_import_packed): every buffer finalize() uploads, in upload
    // order.  `records` remembers (device pointer or host copy, bytes); `replay` is the cursor over an imported image.
    struct Record { const void* dev; std::vector<uint8_t> host; size_t bytes; };
    std::vector<Record> records;
    const uint8_t* replay = nullptr;
    const uint8_t* replay_end = nullptr;
    float *stem_w = nullptr, *stem_scale = nullptr, *stem_shift = nullptr;
    void* stem_tc_img = nullptr;      // tensor-core stem: packed weight tile + scale*2^-e
    float* stem_tc_scale = nullptr;
    __half *stem2_hi = nullptr, *stem2_lo = nullptr;  // fused stem + max-pool: [64][256] K-major planes of the space-to-depth filter
    float* stem2_scale = nullptr;                     // (derived on the device from stem_w; not part of the packed image)
    PredW bbox_pred, cls_pred, cls_memory_pred;
    float dw_cls[3] = {0, 0, 0}, dw_reg[3] = {0, 0, 0};  // softmax(GroupDW.weight)
    float *adjust = nullptr, *bias4 = nullptr;
    Arena arena;
    // CUDA-graph cache of track() for small batches: key = (n, size, nz, nq); valid while the arena has not moved
    struct GraphEntry { cudaGraphExec_t exec = nullptr; uint64_t arena_gen = 0, weights_gen = 0; int seen = 0; long long launches[16] = {0}; };
    std::map<std::tuple<int, int, int, int>, GraphEntry> graphs;
    uint64_t arena_gen = 0;
    uint64_t weights_gen = 0;         // bumped by every (re)pack: captured graphs hold weight pointers and weight tensor maps
    cudaEvent_t ev_done = nullptr;    // end of the last call that used the arena / frame workspace: the next call's stream waits on it
    cudaStream_t gstream = nullptr;   // graphs are captured and replayed on an engine-owned stream (the caller's may be the
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;  // legacy default stream, which cannot be captured); ordered by events
    std::mutex mu;  // one forward at a time per engine (DataParallel replicas own separate engines)
    // scratch of usot_engine_track_frame (crop, gathered memory templates, score / box maps, xf, pool box); grow-only, outside the arena
    char* frame_ws = nullptr;
    size_t frame_ws_cap = 0;
    std::mutex frame_mu;  // one usot_engine_track_frame at a time per engine (the workspace is shared; use one stream per engine)

    ~usot_engine() {
        cudaSetDevice(device);
        for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        if (gstream) cudaStreamDestroy(gstream);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_out) cudaEventDestroy(ev_out);
        if (ev_done) cudaEventDestroy(ev_done);
        for (void* p : owned) cudaFree(p);
        if (arena.base) cudaFree(arena.base);
        if (frame_ws) cudaFree(frame_ws);
    }
};

namespace usot {

#define RUN(expr)                        \
    do {                                 \
        if (!ar.plan) {                  \
            int _rc = (expr);            \
            if (_rc) return _rc;         \
        }                                \
    } while (0)

// Next record of an imported packed image: [u64 bytes][payload, padded to 8 bytes]
static int replay_next(usot_engine* e, const uint8_t** data, size_t* bytes) {
    USOT_REQUIRE(e->replay + 8 <= e->replay_end, "packed weight image is truncated");
    uint64_t n = 0;
    memcpy(&n, e->replay, 8);
    const uint8_t* p = e->replay + 8;
    const size_t padded = (size_t)((n + 7) & ~uint64_t(7));
    USOT_REQUIRE(padded <= (size_t)(e->replay_end - p), "packed weight image is truncated");
    *data = p;
    *bytes = (size_t)n;
    e->replay = p + padded;
    return 0;
}

// Upload one packed buffer (or, when replaying an imported image, the next record instead of `host`; `expect` = 0 skips the size check).
static int upload_bytes(usot_engine* e, const void* host, size_t bytes, size_t expect, void** out) {
    if (e->replay) {
        const uint8_t* p = nullptr;
        if (int rc = replay_next(e, &p, &bytes)) return rc;
        USOT_REQUIRE(expect == 0 || bytes == expect, "packed weight image does not match this architecture / precision");
        host = p;
    }
    void* d = nullptr;
    USOT_CUDA_OK(cudaMalloc(&d, bytes ? bytes : 1));
    USOT_CUDA_OK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
    e->owned.push_back(d);
    e->weight_bytes += (int64_t)bytes;
    e->records.push_back({d, {}, bytes});
    *out = d;
    return 0;
}
static int upload(usot_engine* e, const std::vector<float>& h, float** out, size_t expect_count = 0) {
    return upload_bytes(e, h.data(), h.size() * sizeof(float), expect_count * sizeof(float), reinterpret_cast<void**>(out));
}
static int upload_half(usot_engine* e, const std::vector<__half>& h, __half** out, size_t expect_count = 0) {
    return upload_bytes(e, h.data(), h.size() * sizeof(__half), expect_count * sizeof(__half), reinterpret_cast<void**>(out));
}
// Small host-side constants that finalize() derives from the state_dict (softmaxed GroupDW weights) travel in the image too.
static int host_record(usot_engine* e, float* data, size_t count) {
    if (e->replay) {
        const uint8_t* p = nullptr;
        size_t bytes = 0;
        if (int rc = replay_next(e, &p, &bytes)) return rc;
        USOT_REQUIRE(bytes == count * sizeof(float), "packed weight image does not match this architecture / precision");
        memcpy(data, p, bytes);
    }
    usot_engine::Record r{nullptr, {}, count * sizeof(float)};
    r.host.assign(reinterpret_cast<uint8_t*>(data), reinterpret_cast<uint8_t*>(data) + count * sizeof(float));
    e->records.push_back(std::move(r));
    return 0;
}

static const std::vector<float>* find(usot_engine* e, const std::string& k, size_t numel) {
    auto it = e->host.find(k);
    if (it == e->host.end()) { set_error("usot_b200: missing state_dict tensor '" + k + "'"); return nullptr; }
    if (it->second.size() != numel) {
        set_error("usot_b200: tensor '" + k + "' has " + std::to_string(it->second.size()) + " elements, expected " + std::to_string(numel));
        return nullptr;
    }
    return &it->second;
}

// scale/shift of conv(+bias)+BN folded in double precision (eps = 1e-5, nn.BatchNorm2d default)
static int fold_bn(usot_engine* e, const std::string& conv, const std::string& bn, int cout, bool bias, std::vector<float>& scale,
                   std::vector<float>& shift) {
    scale.assign(cout, 1.f);
    shift.assign(cout, 0.f);
    const std::vector<float>* b = nullptr;
    if (bias && !(b = find(e, conv + ".bias", cout))) return 3;
    if (bn.empty()) {
        if (b) shift = *b;
        return 0;
    }
    const auto *g = find(e, bn + ".weight", cout), *be = find(e, bn + ".bias", cout), *mu = find(e, bn + ".running_mean", cout),
               *var = find(e, bn + ".running_var", cout);
    if (!g || !be || !mu || !var) return 3;
    for (int c = 0; c < cout; ++c) {
        double sc = (double)(*g)[c] / std::sqrt((double)(*var)[c] + 1e-5);
        double sh = (double)(*be)[c] - (double)(*mu)[c] * sc + (b ? (double)(*b)[c] * sc : 0.0);
        scale[c] = (float)sc;
        shift[c] = (float)sh;
    }
    return 0;
}

// Packs every weight (BN folding in double, layout changes, fp16 hi/lo split) and uploads the result.  With e->replay set
// (usot_engine_import_packed) nothing is computed and no state_dict tensor is needed: each upload takes the next record of the
// imported image instead, in the same order.
static int finalize_impl(usot_engine* e) {
    USOT_CUDA_OK(cudaSetDevice(e->device));
    // Captured track() graphs carry the old weight pointers / weight tensor maps as kernel parameters: make every one of them
    // stale BEFORE the buffers are freed (replay is gated on weights_gen), and drop the executables.
    ++e->weights_gen;
    USOT_CUDA_OK(cudaDeviceSynchronize());  // nothing in flight may still read the buffers freed below
    for (auto& g : e->graphs)
        if (g.second.exec) { cudaGraphExecDestroy(g.second.exec); g.second.exec = nullptr; }
    for (void* p : e->owned) cudaFree(p);
    e->owned.clear();
    e->convs.clear();
    e->records.clear();
    e->weight_bytes = 0;
    e->finalized = false;
    const bool rp = e->replay != nullptr;
    const bool tc = e->precision != USOT_PREC_FP32_SIMT;
    std::vector<float> scale, shift, packed;
    {   // stem: OIHW (64,3,7,7) -> [(c*7+kh)*7+kw][co]
        std::vector<uint8_t> img;
        std::vector<float> scale_tc, as_f;
        if (!rp) {
            const auto* w = find(e, "features.features.conv1.weight", 64 * 147);
            if (!w) return 3;
            packed.assign(147 * 64, 0.f);
            for (int co = 0; co < 64; ++co)
                for (int k = 0; k < 147; ++k) packed[(size_t)k * 64 + co] = (*w)[(size_t)co * 147 + k];
            if (int rc = fold_bn(e, "features.features.conv1", "features.features.bn1", 64, false, scale, shift)) return rc;
            if (tc) {
                pack_stem_tc_host(w->data(), scale.data(), img, scale_tc);
                as_f.resize(img.size() / 4);
                memcpy(as_f.data(), img.data(), img.size());
            }
        }
        if (upload(e, packed, &e->stem_w, 147 * 64) || upload(e, scale, &e->stem_scale, 64) || upload(e, shift, &e->stem_shift, 64)) return 1;
        if (tc) {
            float* d = nullptr;
            if (upload(e, as_f, &d, stem_tc_image_bytes() / sizeof(float)) || upload(e, scale_tc, &e->stem_tc_scale, 64)) return 1;
            e->stem_tc_img = d;
        }
    }
    for (const ConvSpec& s : build_specs()) {
        const int K = s.k * s.k * s.cin;
        std::vector<__half> hi, lo;
        std::vector<float> scale_tc;
        if (!rp) {
            const auto* w = find(e, s.name + ".weight", (size_t)s.cout * K);
            if (!w) return 3;
            packed.assign((size_t)K * s.cout, 0.f);
            for (int co = 0; co < s.cout; ++co)
                for (int c = 0; c < s.cin; ++c)
                    for (int t = 0; t < s.k * s.k; ++t)
                        packed[((size_t)t * s.cin + c) * s.cout + co] = (*w)[((size_t)co * s.cin + c) * s.k * s.k + t];
            if (int rc = fold_bn(e, s.name, s.bn, s.cout, s.bias, scale, shift)) return rc;
            if (tc) pack_tc_weights_host(packed.data(), K, s.cout, scale.data(), hi, lo, scale_tc);
        }
        ConvW cw;
        cw.s = s;
        const size_t nw = (size_t)K * s.cout;
        if (upload(e, packed, &cw.w_kn, nw) || upload(e, scale, &cw.scale, s.cout) || upload(e, shift, &cw.shift, s.cout)) return 1;
        if (tc)
            if (upload_half(e, hi, &cw.w_hi, nw) || upload_half(e, lo, &cw.w_lo, nw) || upload(e, scale_tc, &cw.scale_tc, s.cout)) return 1;
        e->convs[s.name] = cw;
    }
    auto pack_pred = [&](const std::string& name, int cout, PredW& pw) -> int {
        std::vector<float> packed4, bias;
        const size_t nw = (size_t)9 * cout * 256;
        if (!rp) {
            const auto* w = find(e, name + ".weight", nw);
            const auto* b = find(e, name + ".bias", cout);
            if (!w || !b) return 3;
            packed.assign(nw, 0.f);
            for (int co = 0; co < cout; ++co)
                for (int c = 0; c < 256; ++c)
                    for (int t = 0; t < 9; ++t) packed[((size_t)t * cout + co) * 256 + c] = (*w)[((size_t)co * 256 + c) * 9 + t];
            packed4.resize(nw);
            for (int k = 0; k < 9 * cout; ++k)
                for (int c = 0; c < 256; ++c) packed4[((size_t)(c / 4) * 9 * cout + k) * 4 + (c & 3)] = packed[(size_t)k * 256 + c];
            bias = *b;
        }
        pw.cout = cout;
        if (upload(e, packed, &pw.w, nw) || upload(e, packed4, &pw.w4, nw) || upload(e, bias, &pw.b, cout)) return 1;
        return 0;
    };
    if (int rc = pack_pred("connect_model.bbox_pred", 4, e->bbox_pred)) return rc;
    if (int rc = pack_pred("connect_model.cls_pred", 1, e->cls_pred)) return rc;
    if (int rc = pack_pred("connect_model.cls_memory_pred", 1, e->cls_memory_pred)) return rc;
    auto softmax3 = [&](const std::string& name, float* out) -> int {
        if (!rp) {
            const auto* w = find(e, name, 3);
            if (!w) return 3;
            double m = std::fmax((*w)[0], std::fmax((*w)[1], (*w)[2]));
            double ex[3], sum = 0;
            for (int i = 0; i < 3; ++i) { ex[i] = std::exp((double)(*w)[i] - m); sum += ex[i]; }
            for (int i = 0; i < 3; ++i) out[i] = (float)(ex[i] / sum);
        }
        return host_record(e, out, 3);
    };
    if (int rc = softmax3("connect_model.cls_dw.weight", e->dw_cls)) return rc;
    if (int rc = softmax3("connect_model.reg_dw.weight", e->dw_reg)) return rc;
    std::vector<float> adj, b4;
    if (!rp) {
        const auto* a1 = find(e, "connect_model.adjust", 1);
        const auto* a4 = find(e, "connect_model.bias", 4);
        if (!a1 || !a4) return 3;
        adj = *a1;
        b4 = *a4;
    }
    if (upload(e, adj, &e->adjust, 1) || upload(e, b4, &e->bias4, 4)) return 1;
    if (rp) USOT_REQUIRE(e->replay == e->replay_end, "packed weight image has trailing data (built for another architecture?)");
    if (tc) {
        // Fused stem + max-pool: the 7x7/2 filter re-indexed for the space-to-depth image (4x4 taps x 16 channels, K = 256)
        void *h2 = nullptr, *l2 = nullptr, *s2 = nullptr, *scratch = nullptr;
        USOT_CUDA_OK(cudaMalloc(&h2, 64 * 256 * sizeof(__half))); e->owned.push_back(h2);
        USOT_CUDA_OK(cudaMalloc(&l2, 64 * 256 * sizeof(__half))); e->owned.push_back(l2);
        USOT_CUDA_OK(cudaMalloc(&s2, 64 * sizeof(float))); e->owned.push_back(s2);
        USOT_CUDA_OK(cudaMalloc(&scratch, 256 * 64 * sizeof(float))); e->owned.push_back(scratch);
        e->stem2_hi = static_cast<__half*>(h2); e->stem2_lo = static_cast<__half*>(l2); e->stem2_scale = static_cast<float*>(s2);
        if (int rc = launch_stem_s2d_weights(e->stem_w, e->stem_scale, static_cast<float*>(scratch), e->stem2_hi, e->stem2_lo, e->stem2_scale, 0)) return rc;
    }
    if (tc) {
        // Fused Conf_Fusion (connect.py:104-144): conf_gen and value_gen read the same input, so they run as ONE conv whose 128-column
        // weight tiles hold 64 conf channels followed by the same 64 value channels (conv_tc.cu, EPI = 1).  The interleaved planes are
        // DERIVED on the device from the two packed layers (not part of the packed-weight image).
        const ConvW& cg = e->convs["connect_model.conf_fusion.conf_gen.0"];
        const ConvW& vg = e->convs["connect_model.conf_fusion.value_gen.0"];
        ConvW f;
        f.s = cg.s;
        f.s.name = "connect_model.conf_fusion.fused";
        f.s.cout = 512;
        const size_t K = (size_t)9 * 256;
        void *whi = nullptr, *wlo = nullptr, *sc = nullptr, *sh = nullptr;
        USOT_CUDA_OK(cudaMalloc(&whi, 512 * K * sizeof(__half))); e->owned.push_back(whi);
        USOT_CUDA_OK(cudaMalloc(&wlo, 512 * K * sizeof(__half))); e->owned.push_back(wlo);
        USOT_CUDA_OK(cudaMalloc(&sc, 512 * sizeof(float))); e->owned.push_back(sc);
        USOT_CUDA_OK(cudaMalloc(&sh, 512 * sizeof(float))); e->owned.push_back(sh);
        f.w_hi = static_cast<__half*>(whi); f.w_lo = static_cast<__half*>(wlo);
        f.scale_tc = static_cast<float*>(sc); f.shift = static_cast<float*>(sh);
        for (int b = 0; b < 4; ++b)
            for (int half = 0; half < 2; ++half) {
                const ConvW& src = half ? vg : cg;
                const size_t drow = (size_t)128 * b + 64 * half, srow = (size_t)64 * b;
                USOT_CUDA_OK(cudaMemcpy(f.w_hi + drow * K, src.w_hi + srow * K, 64 * K * sizeof(__half), cudaMemcpyDeviceToDevice));
                USOT_CUDA_OK(cudaMemcpy(f.w_lo + drow * K, src.w_lo + srow * K, 64 * K * sizeof(__half), cudaMemcpyDeviceToDevice));
                USOT_CUDA_OK(cudaMemcpy(f.scale_tc + drow, src.scale_tc + srow, 64 * sizeof(float), cudaMemcpyDeviceToDevice));
                USOT_CUDA_OK(cudaMemcpy(f.shift + drow, src.shift + srow, 64 * sizeof(float), cudaMemcpyDeviceToDevice));
            }
        e->convs[f.s.name] = f;
    }
    USOT_CUDA_OK(cudaDeviceSynchronize());
    e->finalized = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// forward graphs
// ---------------------------------------------------------------------------------------------
struct Ctx {
    usot_engine* e;
    Arena& ar;
    cudaStream_t st;
    bool tc() const { return e->precision != USOT_PREC_FP32_SIMT; }
    bool split() const { return e->precision == USOT_PREC_FP16X3_TC; }
};

// make sure `t` has the fp16 planes the tensor-core path reads (hi always, lo in split mode)
static int ensure_split(Ctx& c, T& t) {
    Arena& ar = c.ar;
    if (t.hi) return 0;
    USOT_REQUIRE(t.f != nullptr, "tensor has no storage");
    t.hi = ar.h(t.numel());
    t.lo = c.split() ? ar.h(t.numel()) : nullptr;
    if (!ar.plan) {
        Scope sc(FAM_OTHER, c.st, 0, (double)t.numel() * (4 + (c.split() ? 4 : 2)));
        RUN(launch_f32_to_split(t.f, t.numel(), t.hi, t.lo, c.st));
    }
    return 0;
}

enum { OUT_F32 = 1, OUT_SPLIT = 2 };

// One dense conv (+folded BN, +residual, +ReLU).  `want` selects the storage of the result (tensor-core mode only; the SIMT
// path always produces fp32).  `f32_dst` optionally places the fp32 result in caller memory.
static int run_conv(Ctx& c, const std::string& name, T& in, T* residual, int want, T* out, float* f32_dst = nullptr) {
    Arena& ar = c.ar;
    auto it = c.e->convs.find(name);
    USOT_REQUIRE(it != c.e->convs.end(), "unknown conv layer");
    const ConvW& cw = it->second;
    USOT_REQUIRE(in.c == cw.s.cin, "conv input channel mismatch");
    ConvGeom g;
    g.n = in.n; g.h = in.h; g.w = in.w; g.cin = cw.s.cin; g.cout = cw.s.cout; g.kh = g.kw = cw.s.k;
    g.stride = cw.s.stride; g.ph = cw.s.ph; g.pw = cw.s.pw; g.dh = cw.s.dh; g.dw = cw.s.dw;
    g.ho = conv_out(in.h, g.kh, g.stride, g.ph, g.dh);
    g.wo = conv_out(in.w, g.kw, g.stride, g.pw, g.dw);
    USOT_REQUIRE(g.ho > 0 && g.wo > 0, "conv output is empty");
    T o;
    o.n = in.n; o.h = g.ho; o.w = g.wo; o.c = g.cout;
    const double flops = 2.0 * o.n * g.ho * g.wo * (double)g.cout * g.kh * g.kw * g.cin;
    if (!c.tc()) {
        o.f = f32_dst ? f32_dst : ar.f(o.numel());
        USOT_REQUIRE(!residual || residual->f, "SIMT conv needs an fp32 residual");
        Epilogue ep{cw.scale, cw.shift, residual ? residual->f : nullptr, cw.s.relu ? 1 : 0};
        if (!ar.plan) {
            Scope sc(FAM_CONV, c.st, flops);
            RUN(launch_conv_simt(in.f, g, cw.w_kn, ep, o.f, c.st));
        }
    } else {
        if (int rc = ensure_split(c, in)) return rc;
        if (residual)
            if (int rc = ensure_split(c, *residual)) return rc;
        if (f32_dst) want |= OUT_F32;
        if (want & OUT_F32) o.f = f32_dst ? f32_dst : ar.f(o.numel());
        if (want & OUT_SPLIT) {
            o.hi = ar.h(o.numel());
            o.lo = c.split() ? ar.h(o.numel()) : nullptr;
        }
        // (in single-fp16 mode the lo planes do not exist: the kernel neither reads nor writes them)
        TcTensor ti{in.hi, in.lo};
        TcWeights tw{cw.w_hi, cw.w_lo, cw.scale_tc, g.kh * g.kw * g.cin};
        TcEpilogue ep{cw.shift, residual ? residual->hi : nullptr, residual ? residual->lo : nullptr, o.hi, o.lo, o.f, cw.s.relu ? 1 : 0};
        if (!ar.plan) {
            Scope sc(FAM_CONV, c.st, flops);
            RUN(launch_conv_tc(ti, g, tw, ep, c.split(), c.st));
        }
    }
    *out = o;
    return 0;
}

// x (n,3,S,S) nchw -> xf (n,F,F,256) nhwc; fp32 copy written to xf_dst if given; split planes kept for the encoders
static int backbone_neck(Ctx& c, const float* x, int n, int S, float* xf_dst, T* xf_out) {
    Arena& ar = c.ar;
    usot_engine* e = c.e;
    const int h1 = (S - 7) / 2 + 1;
    const int h2 = (h1 + 2 - 3) / 2 + 1;
    // stem + max-pool as ONE tensor-core kernel over the space-to-depth image (no (n, h1, h1, 64) fp32 map in HBM, no software im2col);
    // maps wider than one 128-pixel tile (271-pixel crops) keep the two kernels
    const bool fused_pool = c.tc() && g_stem_tc && g_stem_pool_fused && stem_pool_bands(n, S) > 0;
    __half *s2d_hi = nullptr, *s2d_lo = nullptr;
    if (fused_pool) {
        s2d_hi = ar.h(stem_s2d_plane_elems(n, S));
        s2d_lo = c.split() ? ar.h(stem_s2d_plane_elems(n, S)) : nullptr;
    }
    float* a0 = fused_pool ? nullptr : ar.f((size_t)n * h1 * h1 * 64);
    if (!ar.plan && !fused_pool) {
        Scope sc(FAM_STEM, c.st, 2.0 * n * h1 * h1 * 64.0 * 147);
        if (c.tc() && g_stem_tc) RUN(launch_stem_tc(x, n, S, e->stem_tc_img, e->stem_tc_scale, e->stem_shift, a0, c.split(), c.st));
        else RUN(launch_stem(x, n, S, e->stem_w, e->stem_scale, e->stem_shift, a0, c.st));
    }
    T cur;
    cur.n = n; cur.h = cur.w = h2; cur.c = 64;
    if (c.tc()) {  // only the tensor-core convs read the pooled map: write their operand format directly
        cur.hi = ar.h(cur.numel());
        cur.lo = c.split() ? ar.h(cur.numel()) : nullptr;
    } else {
        cur.f = ar.f(cur.numel());
    }
    if (!ar.plan && fused_pool) {
        Scope sc(FAM_STEM, c.st, 2.0 * n * h1 * h1 * 64.0 * 147);
        count_launches(FAM_STEM, 1); tl_launches[FAM_STEM]++;   // (two kernels: image -> space-to-depth planes, then the GEMM + pooling epilogue)
        RUN(launch_stem_s2d_pool(x, n, S, e->stem2_hi, e->stem2_lo, e->stem2_scale, e->stem_shift, s2d_hi, s2d_lo, cur.hi, cur.lo, c.split(), c.st));
    }
    if (!ar.plan && !fused_pool) {
        Scope sc(FAM_POOL, c.st, 0, 4.0 * n * 64 * ((double)h1 * h1 + (double)h2 * h2));
        RUN(launch_maxpool3x3s2p1(a0, n, h1, h1, 64, cur.f, cur.hi, cur.lo, c.st));
    }
    struct L { const char* name; int blocks; };
    const L layers[3] = {{"layer1", 3}, {"layer2", 4}, {"layer3", 6}};
    for (const L& l : layers) {
        for (int i = 0; i < l.blocks; ++i) {
            const std::string p = std::string("features.features.") + l.name + "." + std::to_string(i) + ".";
            T res, t1, t2, t3;
            if (i == 0)
                if (int rc = run_conv(c, p + "downsample.0", cur, nullptr, OUT_SPLIT, &res)) return rc;
            if (int rc = run_conv(c, p + "conv1", cur, nullptr, OUT_SPLIT, &t1)) return rc;
            if (i > 0) res = cur;  // identity shortcut (taken after conv1 so fp16 planes materialised for it are shared)
            if (int rc = run_conv(c, p + "conv2", t1, nullptr, OUT_SPLIT, &t2)) return rc;
            if (int rc = run_conv(c, p + "conv3", t2, &res, OUT_SPLIT, &t3)) return rc;
            cur = t3;
        }
    }
    if (int rc = run_conv(c, "neck.downsample.0", cur, nullptr, OUT_F32 | OUT_SPLIT, xf_out, xf_dst)) return rc;
    return 0;
}

struct Enc3 {
    T m[3];
};

// matrix.forward for one branch: three parallel 3x3 convs on the same input (connect.py:55-74); fp32 results (xcorr input)
static int encode(Ctx& c, const char* enc, char branch, T& in, Enc3* out) {
    static const char* mn[3] = {"matrix11", "matrix12", "matrix21"};
    for (int i = 0; i < 3; ++i) {
        const std::string name = std::string("connect_model.") + enc + "." + mn[i] + "_" + branch + ".0";
        if (int rc = run_conv(c, name, in, nullptr, OUT_F32, &out->m[i])) return rc;
    }
    return 0;
}

static int tower(Ctx& c, const char* name, T in, T* out) {
    T cur = in;
    for (int i = 0; i < 4; ++i) {
        T o;
        if (int rc = run_conv(c, std::string("connect_model.") + name + "." + std::to_string(3 * i), cur, nullptr,
                              i == 3 ? OUT_F32 : OUT_SPLIT, &o))
            return rc;
        cur = o;
    }
    *out = cur;
    return 0;
}

static int groupdw(Ctx& c, const Enc3& x, const Enc3& z, int n_out, int F, const float* w3, T* out) {
    Arena& ar = c.ar;
    const int nx = x.m[0].n, nz = z.m[0].n, R = F - 6;
    T o;
    o.n = n_out; o.h = o.w = R; o.c = 256;
    // every consumer of the correlation map is a tensor-core conv (towers, conf/value generators): in the tcgen05 modes the FFMA2
    // kernel writes their split-fp16 operand planes directly (same 4 B/element as fp32, no separate conversion pass)
    const bool split_out = c.tc() && groupdw_split_output_supported(F);
    if (split_out) {
        o.hi = ar.h(o.numel());
        o.lo = c.split() ? ar.h(o.numel()) : nullptr;
    } else {
        o.f = ar.f(o.numel());
    }
    GroupDWArgs a;
    a.x11 = x.m[0].f; a.x12 = x.m[1].f; a.x21 = x.m[2].f;
    a.z11 = z.m[0].f; a.z12 = z.m[1].f; a.z21 = z.m[2].f;
    a.dw_weight = nullptr; a.out = o.f; a.nx = nx; a.nz = nz; a.n_out = n_out; a.C = 256; a.F = F;
    a.out_hi = o.hi; a.out_lo = o.lo;
    if (!ar.plan) {
        // algorithmic bytes: every search map read once, every output written once, taps once
        const double per_x = 256.0 * ((F - 2.0) * (F - 2) + 2.0 * (F - 4) * (F - 2));
        Scope sc(FAM_XCORR, c.st, 0, 4.0 * (nx * per_x + n_out * 256.0 * R * R + nz * 256.0 * 55));
        RUN(launch_groupdw_w(a, w3[0], w3[1], w3[2], c.st));
    }
    *out = o;
    return 0;
}

static int pred(Ctx& c, const PredW& pw, const T& in, int mode, float* out) {
    Arena& ar = c.ar;
    usot_engine* e = c.e;
    if (!ar.plan) {
        // algorithmic bytes: the tower output read once, the (n,cout,R,R) map written once, the weights once
        Scope sc(FAM_PRED, c.st, 2.0 * in.n * in.h * in.w * 256.0 * 9 * pw.cout,
                 4.0 * ((double)in.n * in.h * in.w * (256.0 + pw.cout) + 9.0 * pw.cout * 256));
        RUN(launch_pred_conv(in.f, in.n, in.h, 256, pw.w, pw.w4, pw.b, pw.cout, mode, mode == 0 ? 0.1f : 1.f, e->adjust, e->bias4, out, c.st));
    }
    return 0;
}

static T wrap_f32(const float* p, int n, int h, int w, int ch) {
    T t;
    t.f = const_cast<float*>(p); t.n = n; t.h = h; t.w = w; t.c = ch;
    return t;
}

// box_tower_reg.forward, offline branch (connect.py:224-247). Returns the encoded cls search maps for the memory branch.
static int head_offline(Ctx& c, T& xf, const float* zf, int nz, float* cls, float* bbox, Enc3* cls_x) {
    usot_engine* e = c.e;
    const int n = xf.n, F = xf.h;
    Enc3 cls_z, reg_z, reg_x;
    T z = wrap_f32(zf, nz, 7, 7, 256);
    if (int rc = encode(c, "cls_encode", 'k', z, &cls_z)) return rc;
    if (int rc = encode(c, "cls_encode", 's', xf, cls_x)) return rc;
    if (int rc = encode(c, "reg_encode", 'k', z, &reg_z)) return rc;
    if (int rc = encode(c, "reg_encode", 's', xf, &reg_x)) return rc;
    T cls_dw, reg_dw, x_reg, x_cls;
    if (int rc = groupdw(c, *cls_x, cls_z, n, F, e->dw_cls, &cls_dw)) return rc;
    if (int rc = groupdw(c, reg_x, reg_z, n, F, e->dw_reg, &reg_dw)) return rc;
    if (int rc = tower(c, "bbox_tower", reg_dw, &x_reg)) return rc;
    if (int rc = pred(c, e->bbox_pred, x_reg, 1, bbox)) return rc;
    if (int rc = tower(c, "cls_tower", cls_dw, &x_cls)) return rc;
    if (int rc = pred(c, e->cls_pred, x_cls, 0, cls)) return rc;
    return 0;
}

// box_tower_reg.forward, memory branch (connect.py:248-280). mem: (n*nq,7,7,256) nhwc.
// `mem_n` distinct kernels serve the n*nq (sample, slot) pairs in order (mem_n == n*nq at inference; in the cycle-memory
// forward pass one pooled feature per sample is shared by its m memory frames, models.py:240-244).
static int head_memory(Ctx& c, const Enc3& cls_x, int n, int F, const float* mem, int nq, int mem_n, float* cls_mem) {
    Arena& ar = c.ar;
    usot_engine* e = c.e;
    const int R = F - 6;
    Enc3 mz;
    T m = wrap_f32(mem, mem_n, 7, 7, 256);
    if (int rc = encode(c, "cls_encode", 'k', m, &mz)) return rc;
    T dw, conf, val, t;
    if (int rc = groupdw(c, cls_x, mz, n * nq, F, e->dw_cls, &dw)) return rc;
    const size_t per_map = (size_t)R * R * 256;
    T fused;
    fused.n = n; fused.h = fused.w = R; fused.c = 256;
    if (c.split() && g_conf_fusion_fused) {   // (single-fp16 mode: its two N = 256 tiles share more operand traffic than the 64 + 64 tile saves)
        // ONE conv for conf_gen || value_gen (every activation box is loaded once for both) with the clamp / exp / sum over the N_q maps /
        // divide in its epilogue: the two (n*nq, R, R, 256) fp32 maps are never written, and the result leaves as the split-fp16
        // planes the memory tower reads (no conversion pass).
        const ConvW& cw = e->convs.find("connect_model.conf_fusion.fused")->second;
        if (int rc = ensure_split(c, dw)) return rc;
        fused.hi = ar.h(fused.numel());
        fused.lo = c.split() ? ar.h(fused.numel()) : nullptr;
        ConvGeom g;
        g.n = n * nq; g.h = g.w = R; g.cin = 256; g.cout = 512; g.kh = g.kw = 3; g.stride = 1; g.ph = g.pw = 1; g.dh = g.dw = 1; g.ho = g.wo = R;
        TcTensor ti{dw.hi, dw.lo};
        TcWeights tw{cw.w_hi, cw.w_lo, cw.scale_tc, 9 * 256};
        TcEpilogue ep{cw.shift, nullptr, nullptr, fused.hi, fused.lo, nullptr, 1};
        if (!ar.plan) {
            Scope sc(FAM_CONV, c.st, 2.0 * g.n * R * R * 512.0 * 9 * 256);
            RUN(launch_conv_tc(ti, g, tw, ep, c.split(), c.st, nq));
        }
    } else {
        if (int rc = run_conv(c, "connect_model.conf_fusion.conf_gen.0", dw, nullptr, OUT_F32, &conf)) return rc;
        if (int rc = run_conv(c, "connect_model.conf_fusion.value_gen.0", dw, nullptr, OUT_F32, &val)) return rc;
        fused.f = ar.f(fused.numel());
        if (!ar.plan) {
            Scope sc(FAM_FUSION, c.st, 0, 4.0 * per_map * (2.0 * n * nq + n));
            RUN(launch_conf_fusion(conf.f, val.f, n, nq, per_map, fused.f, c.st));
        }
    }
    if (int rc = tower(c, "cls_memory_tower", fused, &t)) return rc;
    if (int rc = pred(c, e->cls_memory_pred, t, 0, cls_mem)) return rc;
    return 0;
}

static int prroi(Ctx& c, const T& feat, const float* boxes, int n_rois, float* out) {
    Arena& ar = c.ar;
    if (!ar.plan) {
        Scope sc(FAM_PRROI, c.st);
        RUN(launch_prroi_nhwc(feat.f, feat.n, feat.h, feat.w, feat.c, boxes, n_rois, out, c.st));
    }
    return 0;
}


// Two-pass driver: plan (count arena bytes) -> grow arena if needed -> execute.
// The arena (and the frame workspace) is reused from offset 0 by every call, so calls are ordered ON THE DEVICE as well as on the
// host: each call's stream first waits for the event recorded at the end of the previous call (a no-op on the same stream) --
// callers may therefore issue calls of one engine from different streams.
static int order_begin(usot_engine* e, cudaStream_t st) {
    if (!e->ev_done) USOT_CUDA_OK(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));
    else USOT_CUDA_OK(cudaStreamWaitEvent(st, e->ev_done, 0));
    return 0;
}
static int order_end(usot_engine* e, cudaStream_t st) {
    USOT_CUDA_OK(cudaEventRecord(e->ev_done, st));
    return 0;
}

template <typename Fn>
static int with_arena(usot_engine* e, cudaStream_t st, Fn&& body, bool ordered = true) {
    USOT_REQUIRE(e && e->finalized, "engine not finalized (load the state_dict and call usot_engine_finalize)");
    std::lock_guard<std::mutex> lk(e->mu);
    USOT_CUDA_OK(cudaSetDevice(e->device));
    if (ordered)
        if (int rc = order_begin(e, st)) return rc;
    Arena& ar = e->arena;
    ar.plan = true;
    ar.off = 0;
    if (int rc = body(ar)) { ar.plan = false; return rc; }
    const size_t need = ar.off + 256;
    ar.plan = false;
    if (need > ar.cap) {
        USOT_CUDA_OK(cudaDeviceSynchronize());
        if (ar.base) USOT_CUDA_OK(cudaFree(ar.base));
        ar.base = nullptr;
        ar.cap = 0;
        void* p = nullptr;
        cudaError_t err = cudaMalloc(&p, need);
        if (err != cudaSuccess) {
            set_error("usot_b200: cannot allocate a " + std::to_string(need >> 20) + " MiB activation arena: " + cudaGetErrorString(err));
            return 1;
        }
        ar.base = static_cast<char*>(p);
        ar.cap = need;
        ++e->arena_gen;  // captured graphs hold arena addresses
    }
    ar.off = 0;
    if (int rc = body(ar)) return rc;
    return ordered ? order_end(e, st) : 0;
}

static Tunable g_graph_max_batch{8};  // tunable "graph_max_batch": track() with n <= this replays a captured CUDA graph (0 = off)

}  // namespace usot

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* usot_last_error(void) { return g_err.c_str(); }
int usot_abi_version(void) { return 3; }

/* Profiling: kernel launches are always counted; with on=1 every launch is also bracketed by CUDA events on its stream. */
int usot_profile_reset(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& e : g_prof.ev) { cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second); }
    g_prof.ev.clear();
    for (int i = 0; i < FAM_COUNT; ++i) { g_prof.launches[i] = 0; g_prof.flops[i] = 0; g_prof.bytes[i] = 0; }
    g_prof.on = on != 0;
    return 0;
}
int usot_profile_family_count(void) { return FAM_COUNT; }
/* Account for launches that did not pass through the library's launchers: a replay of a CUDA graph captured by the CALLER (the graphed
 * training step) launches the kernels recorded at capture time; the caller adds those per-family counts after each replay. */
int usot_profile_count(int fam, int64_t launches) {
    USOT_REQUIRE(fam >= 0 && fam < FAM_COUNT, "bad family");
    count_launches(fam, (long long)launches);
    return 0;
}
const char* usot_profile_family_name(int fam) { return (fam >= 0 && fam < FAM_COUNT) ? kFamNames[fam] : ""; }
/* Synchronises the device.  out[4] = {launches, milliseconds (0 unless profiling was on), algorithmic FLOPs, algorithmic bytes}. */
int usot_profile_read(int fam, double* out) {
    USOT_REQUIRE(fam >= 0 && fam < FAM_COUNT && out, "bad argument");
    USOT_CUDA_OK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double ms = 0;
    for (auto& e : g_prof.ev)
        if (e.first == fam) { float t = 0; cudaEventElapsedTime(&t, e.second.first, e.second.second); ms += t; }
    out[0] = (double)g_prof.launches[fam]; out[1] = ms; out[2] = g_prof.flops[fam]; out[3] = g_prof.bytes[fam];
    return 0;
}

int usot_set_tunable(const char* name, int value) {
    USOT_REQUIRE(name, "null name");
    if (!strcmp(name, "tc_bn_max")) { USOT_REQUIRE(value == 64 || value == 128 || value == 256, "tc_bn_max must be 64, 128 or 256"); g_tc_bn_max = value; return 0; }
    if (!strcmp(name, "graph_max_batch")) { USOT_REQUIRE(value >= 0 && value <= 64, "graph_max_batch must be in [0, 64]"); g_graph_max_batch = value; return 0; }
    if (!strcmp(name, "groupdw_tma")) { USOT_REQUIRE(value >= 0 && value <= 2, "groupdw_tma must be 0 (register-staged), 1 (TMA ring, scalar FMA) or 2 (TMA ring, packed FFMA2)"); g_groupdw_tma = value; return 0; }
    if (!strcmp(name, "conf_fusion_fused")) { USOT_REQUIRE(value == 0 || value == 1, "conf_fusion_fused must be 0 or 1"); g_conf_fusion_fused = value; return 0; }
    if (!strcmp(name, "stem_pool_fused")) { USOT_REQUIRE(value == 0 || value == 1, "stem_pool_fused must be 0 or 1"); g_stem_pool_fused = value; return 0; }
    if (!strcmp(name, "stem_tc")) { USOT_REQUIRE(value == 0 || value == 1, "stem_tc must be 0 or 1"); g_stem_tc = value; return 0; }
    if (!strcmp(name, "tc_tma_res")) { USOT_REQUIRE(value == 0 || value == 1, "tc_tma_res must be 0 or 1"); g_tc_tma_res = value; return 0; }
    if (!strcmp(name, "tc_res_ahead")) { USOT_REQUIRE(value == 1 || value == 2, "tc_res_ahead must be 1 or 2"); g_tc_res_ahead = value; return 0; }
    if (!strcmp(name, "tc_skip_pad_rows")) { USOT_REQUIRE(value == 0 || value == 1, "tc_skip_pad_rows must be 0 or 1"); g_tc_skip_pad_rows = value; return 0; }
    if (!strcmp(name, "tc_multi_image_tiles")) { USOT_REQUIRE(value == 0 || value == 1, "tc_multi_image_tiles must be 0 or 1"); g_tc_multi_image_tiles = value; return 0; }
    if (!strcmp(name, "tc_cta_pair")) { USOT_REQUIRE(value >= 0 && value <= 7, "tc_cta_pair must be 0..7 (bit 0: single-fp16 launches, bit 1: fp16x3 launches run as CTA pairs, bit 2: also where it does not pay)"); g_tc_cta_pair = value; return 0; }
    if (!strcmp(name, "tc_pdl")) { USOT_REQUIRE(value == 0 || value == 1, "tc_pdl must be 0 or 1"); g_tc_pdl = value; return 0; }
    if (!strcmp(name, "tc_latency_split")) { USOT_REQUIRE(value == 0 || value == 1, "tc_latency_split must be 0 or 1"); g_tc_latency_split = value; return 0; }
    if (!strcmp(name, "tc_l2_prefetch")) { USOT_REQUIRE(value == 0 || value == 1, "tc_l2_prefetch must be 0 or 1"); g_tc_l2_prefetch = value; return 0; }
    if (!strcmp(name, "tc_tma_f32")) { USOT_REQUIRE(value == 0 || value == 1, "tc_tma_f32 must be 0 or 1"); g_tc_tma_f32 = value; return 0; }
    if (!strcmp(name, "tc_fuse_cross")) { USOT_REQUIRE(value == 0 || value == 1, "tc_fuse_cross must be 0 or 1"); g_tc_fuse_cross = value; return 0; }
    if (!strcmp(name, "tc_tma_store")) { USOT_REQUIRE(value == 0 || value == 1, "tc_tma_store must be 0 or 1"); g_tc_tma_store = value; return 0; }
    if (!strcmp(name, "tc_split_bn_max")) { USOT_REQUIRE(value == 64 || value == 128 || value == 256, "tc_split_bn_max must be 64, 128 or 256"); g_tc_split_bn_max = value; return 0; }
    if (!strcmp(name, "pred_tma_min_batch")) { USOT_REQUIRE(value >= 0, "pred_tma_min_batch must be >= 0 (0 = never use the TMA-streamed pred kernel)"); g_pred_tma_min_batch = value; return 0; }
    if (!strcmp(name, "groupdw_row_split")) { USOT_REQUIRE(value == 0 || value == 1, "groupdw_row_split must be 0 or 1"); g_groupdw_row_split = value; return 0; }
    if (!strcmp(name, "groupdw_warps4")) { USOT_REQUIRE(value == 0 || value == 1, "groupdw_warps4 must be 0 or 1"); g_groupdw_warps4 = value; return 0; }
    if (!strcmp(name, "groupdw_strips")) { USOT_REQUIRE(value == 2 || value == 3, "groupdw_strips must be 2 or 3"); g_groupdw_strips = value; return 0; }
    USOT_REQUIRE(false, "unknown tunable");
}

int usot_prroi_pool_forward(const float* features, const float* rois, float* output, int n_features, int n_rois, int channels,
                            int height, int width, int pooled_height, int pooled_width, float spatial_scale, void* stream) {
    USOT_REQUIRE(n_rois == 0 || (features && rois && output), "null pointer");
    USOT_REQUIRE(n_features >= 0 && n_rois >= 0 && channels > 0 && height > 0 && width > 0 && pooled_height > 0 && pooled_width > 0,
                 "bad shape");
    return launch_prroi_nchw(features, channels, height, width, rois, n_rois, pooled_height, pooled_width, spatial_scale, output,
                             (cudaStream_t)stream);
}

int usot_prroi_pool_backward(const float* rois, const float* output_diff, float* features_diff, int n_features, int n_rois, int channels,
                             int height, int width, int pooled_height, int pooled_width, float spatial_scale, void* stream) {
    USOT_REQUIRE(features_diff && (n_rois == 0 || (rois && output_diff)), "null pointer");
    USOT_REQUIRE(n_features >= 0 && n_rois >= 0 && channels > 0 && height > 0 && width > 0 && pooled_height > 0 && pooled_width > 0, "bad shape");
    return launch_prroi_backward(rois, output_diff, features_diff, n_features, n_rois, channels, height, width, pooled_height, pooled_width,
                                 spatial_scale, (cudaStream_t)stream);
}

int usot_prroi_pool_coor_backward(const float* features, const float* rois, const float* output, const float* output_diff, float* rois_diff,
                                  int n_rois, int channels, int height, int width, int pooled_height, int pooled_width, float spatial_scale,
                                  void* stream) {
    USOT_REQUIRE(n_rois == 0 || (features && rois && output && output_diff && rois_diff), "null pointer");
    USOT_REQUIRE(n_rois >= 0 && channels > 0 && height > 0 && width > 0 && pooled_height > 0 && pooled_width > 0, "bad shape");
    return launch_prroi_coor_backward(features, rois, output, output_diff, rois_diff, n_rois, channels, height, width, pooled_height,
                                      pooled_width, spatial_scale, (cudaStream_t)stream);
}

int usot_xcorr_depthwise(const float* x, const float* kernel, float* out, int bx, int bk, int channels, int hx, int wx, int hk, int wk,
                         void* stream) {
    USOT_REQUIRE(bx == 0 || (x && kernel && out), "null pointer");
    return launch_xcorr_nchw(x, kernel, out, bx, bk, channels, hx, wx, hk, wk, (cudaStream_t)stream);
}

int usot_xcorr_depthwise_backward(const float* x, const float* kernel, const float* grad_out, float* grad_x, float* grad_kernel, int bx, int bk,
                                  int channels, int hx, int wx, int hk, int wk, void* stream) {
    USOT_REQUIRE(bx == 0 || (x && kernel && grad_out), "null pointer");
    USOT_REQUIRE(grad_x || grad_kernel, "nothing to compute: both gradient outputs are NULL");
    return launch_xcorr_backward(x, kernel, grad_out, grad_x, grad_kernel, bx, bk, channels, hx, wx, hk, wk, (cudaStream_t)stream);
}

int usot_groupdw_xcorr(const float* x11, const float* x12, const float* x21, const float* z11, const float* z12, const float* z21,
                       const float* weight, float* out, int nx, int nz, int n_out, int channels, int feat_size, void* stream) {
    USOT_REQUIRE(x11 && x12 && x21 && z11 && z12 && z21 && weight && out, "null pointer");
    GroupDWArgs a{x11, x12, x21, z11, z12, z21, weight, out, nx, nz, n_out, channels, feat_size};
    return launch_groupdw(a, (cudaStream_t)stream);
}

static int conv2d_nhwc_impl(const float* in, int n, int h, int w, int cin, const float* weight_kn, int cout, int kh, int kw, int stride,
                            int pad_h, int pad_w, int dil_h, int dil_w, const float* scale, const float* shift, const float* residual,
                            int relu, float* out, int precision, void* stream, const float* in_mul);

int usot_conv2d_nhwc(const float* in, int n, int h, int w, int cin, const float* weight_kn, int cout, int kh, int kw, int stride,
                     int pad_h, int pad_w, int dil_h, int dil_w, const float* scale, const float* shift, const float* residual,
                     int relu, float* out, int precision, void* stream) {
    return conv2d_nhwc_impl(in, n, h, w, cin, weight_kn, cout, kh, kw, stride, pad_h, pad_w, dil_h, dil_w, scale, shift, residual, relu, out, precision,
                            stream, nullptr);
}

int usot_conv2d_nhwc_scaled(const float* in, const float* in_scale, int n, int h, int w, int cin, const float* weight_kn, int cout, int kh, int kw,
                            int stride, int pad_h, int pad_w, int dil_h, int dil_w, const float* scale, const float* shift, float* out,
                            int precision, void* stream) {
    USOT_REQUIRE(in_scale, "null in_scale");
    USOT_REQUIRE(precision != USOT_PREC_FP32_SIMT, "usot_conv2d_nhwc_scaled is the tensor-core route (fp32 FMA needs no input scaling)");
    return conv2d_nhwc_impl(in, n, h, w, cin, weight_kn, cout, kh, kw, stride, pad_h, pad_w, dil_h, dil_w, scale, shift, nullptr, 0, out, precision, stream,
                            in_scale);
}

static int conv2d_nhwc_impl(const float* in, int n, int h, int w, int cin, const float* weight_kn, int cout, int kh, int kw, int stride,
                            int pad_h, int pad_w, int dil_h, int dil_w, const float* scale, const float* shift, const float* residual,
                            int relu, float* out, int precision, void* stream, const float* in_mul) {
    USOT_REQUIRE(in && weight_kn && scale && shift && out, "null pointer");
    USOT_REQUIRE(precision >= USOT_PREC_FP32_SIMT && precision <= USOT_PREC_FP16_TC, "unknown precision mode");
    ConvGeom g{n, h, w, cin, cout, kh, kw, stride, pad_h, pad_w, dil_h, dil_w, conv_out(h, kh, stride, pad_h, dil_h),
               conv_out(w, kw, stride, pad_w, dil_w)};
    USOT_REQUIRE(g.ho > 0 && g.wo > 0, "conv output is empty");
    cudaStream_t st = (cudaStream_t)stream;
    Scope sc(FAM_CONV, st, 2.0 * n * g.ho * g.wo * (double)cout * kh * kw * cin);   // (the training path runs every dense conv through this entry)
    if (precision == USOT_PREC_FP32_SIMT) {
        Epilogue ep{scale, shift, residual, relu};
        return launch_conv_simt(in, g, weight_kn, ep, out, st);
    }
    // Tensor-core path of the stand-alone op: weights are split / packed ON THE DEVICE and the scratch planes come from the stream-ordered
    // allocator, so the call enqueues without any host synchronisation (the training path issues one such call per layer and step).
    const bool split = precision == USOT_PREC_FP16X3_TC;
    const int K = kh * kw * cin;
    const size_t n_in = (size_t)n * h * w * cin, n_out = (size_t)n * g.ho * g.wo * cout;
    const size_t n_w = (size_t)K * cout;
    const bool dbg_split_out = getenv("USOT_DEBUG_SPLIT_OUT") != nullptr;  // profiling aid: exercise the split-fp16 output path as the engine does
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t planes = split ? 2 : 1;
    size_t bytes = planes * al(n_w * 2) + al((size_t)cout * 4) + planes * al(n_in * 2) + (residual ? planes * al(n_out * 2) : 0) +
                   (dbg_split_out ? 2 * al(n_out * 2) : 0);
    ensure_async_pool();
    char* ws = nullptr;
    USOT_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), bytes, st));
    char* cur = ws;
    auto take = [&](size_t b) { char* p = cur; cur += al(b); return p; };
    __half* d_whi = reinterpret_cast<__half*>(take(n_w * 2));
    __half* d_wlo = split ? reinterpret_cast<__half*>(take(n_w * 2)) : nullptr;
    float* d_scale = reinterpret_cast<float*>(take((size_t)cout * 4));
    __half* d_ihi = reinterpret_cast<__half*>(take(n_in * 2));
    __half* d_ilo = split ? reinterpret_cast<__half*>(take(n_in * 2)) : nullptr;
    __half *d_rhi = nullptr, *d_rlo = nullptr, *d_ohi = nullptr, *d_olo = nullptr;
    int rc = 0;
    do {
        if ((rc = launch_pack_tc_weights(weight_kn, K, cout, scale, d_whi, d_wlo, d_scale, st))) break;
        if ((rc = launch_f32_to_split(in, n_in, d_ihi, d_ilo, st, in_mul))) break;
        if (residual) {
            d_rhi = reinterpret_cast<__half*>(take(n_out * 2));
            d_rlo = split ? reinterpret_cast<__half*>(take(n_out * 2)) : nullptr;
            if ((rc = launch_f32_to_split(residual, n_out, d_rhi, d_rlo, st))) break;
        }
        TcTensor ti{d_ihi, d_ilo};
        TcWeights tw{d_whi, d_wlo, d_scale, K};
        TcEpilogue ep{shift, d_rhi, d_rlo, nullptr, nullptr, out, relu};
        if (dbg_split_out) {
            d_ohi = reinterpret_cast<__half*>(take(n_out * 2));
            d_olo = reinterpret_cast<__half*>(take(n_out * 2));
            ep.out_hi = d_ohi; ep.out_lo = split ? d_olo : nullptr;
            if (getenv("USOT_DEBUG_SPLIT_OUT")[0] == '2') ep.out_f32 = nullptr;  // split output only
        }
        rc = launch_conv_tc(ti, g, tw, ep, split, st);
    } while (0);
    cudaFreeAsync(ws, st);
    return rc;
}

int usot_pred_conv(const float* in, int n, int r, int channels, const float* weight, const float* bias, int cout, int mode, float mul,
                   const float* adjust, const float* bias4, float* out, void* stream) {
    USOT_REQUIRE(n == 0 || (in && weight && bias && out), "null pointer");
    USOT_REQUIRE(n >= 0 && r > 0 && channels > 0, "bad shape");
    USOT_REQUIRE(mode == 0 || (mode == 1 && adjust && bias4), "mode must be 0, or 1 with adjust and bias4");
    count_launches(FAM_PRED, 1);
    return launch_pred_conv(in, n, r, channels, weight, nullptr, bias, cout, mode, mul, adjust, bias4, out, (cudaStream_t)stream);
}

int usot_crop_resize(const uint8_t* frames, int n_frames, int height, int width, const int32_t* crops, const uint8_t* fill, int n,
                     int model_sz, float* out, void* stream) {
    USOT_REQUIRE(n == 0 || (frames && crops && fill && out), "null pointer");
    USOT_REQUIRE(n >= 0 && n <= 65535 && n_frames > 0 && height > 0 && width > 0 && model_sz > 0 && model_sz <= 4096, "bad shape");
    count_launches(FAM_OTHER, 1);
    return launch_crop_resize(frames, n_frames, height, width, crops, fill, n, model_sz, out, (cudaStream_t)stream);
}

int usot_nchw_to_nhwc(const float* in, int n, int c, int h, int w, float* out, void* stream) {
    return launch_nchw_to_nhwc(in, n, c, h, w, out, (cudaStream_t)stream);
}
int usot_nhwc_to_nchw(const float* in, int n, int h, int w, int c, float* out, void* stream) {
    return launch_nhwc_to_nchw(in, n, h, w, c, out, (cudaStream_t)stream);
}

int usot_engine_create(usot_engine** out, int device, int precision) {
    USOT_REQUIRE(out, "null pointer");
    USOT_REQUIRE(precision >= USOT_PREC_FP32_SIMT && precision <= USOT_PREC_FP16_TC, "unknown precision mode");
    int count = 0;
    USOT_CUDA_OK(cudaGetDeviceCount(&count));
    USOT_REQUIRE(device >= 0 && device < count, "no such CUDA device (usot_b200 has no CPU fallback)");
    cudaDeviceProp prop;
    USOT_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    USOT_REQUIRE(prop.major == 10, "usot_b200 is built for sm_100a (Blackwell B200) only");
    usot_engine* e = new usot_engine();
    e->device = device;
    e->precision = precision;
    *out = e;
    return 0;
}

int usot_engine_destroy(usot_engine* e) {
    delete e;
    return 0;
}

int usot_engine_load_tensor(usot_engine* e, const char* name, const float* host_data, int64_t numel) {
    USOT_REQUIRE(e && name && (host_data || numel == 0) && numel >= 0, "bad argument");
    e->host[name].assign(host_data, host_data + numel);
    e->finalized = false;
    return 0;
}

int usot_engine_finalize(usot_engine* e) {
    USOT_REQUIRE(e, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    return finalize_impl(e);
}

/* Packed-weight image: header {magic, abi, precision, record count, checksum} + records [u64 bytes][payload padded to 8]. */
static const uint64_t kPackedMagic = 0x55534f5442323030ull;  // "USOTB200"
static const int kPackedHeader = 40;

// 64-bit word-wise multiply/xorshift hash of the record area (sizes + payloads; every record is padded to 8 bytes)
static uint64_t packed_checksum(const uint8_t* p, size_t bytes) {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    for (size_t i = 0; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
    }
    return h;
}

int64_t usot_engine_packed_size(const usot_engine* e) {
    if (!e || !e->finalized) return 0;
    int64_t n = kPackedHeader;
    for (const auto& r : e->records) n += 8 + (int64_t)((r.bytes + 7) & ~size_t(7));
    return n;
}

int usot_engine_export_packed(usot_engine* e, void* host_buf, int64_t capacity) {
    USOT_REQUIRE(e && e->finalized, "engine not finalized");
    USOT_REQUIRE(host_buf && capacity >= usot_engine_packed_size(e), "buffer too small (see usot_engine_packed_size)");
    std::lock_guard<std::mutex> lk(e->mu);
    USOT_CUDA_OK(cudaSetDevice(e->device));
    uint8_t* const base = static_cast<uint8_t*>(host_buf);
    uint8_t* p = base + kPackedHeader;
    for (const auto& r : e->records) {
        const uint64_t n = r.bytes;
        memcpy(p, &n, 8);
        p += 8;
        if (r.dev) USOT_CUDA_OK(cudaMemcpy(p, r.dev, r.bytes, cudaMemcpyDeviceToHost));
        else memcpy(p, r.host.data(), r.bytes);
        const size_t padded = (r.bytes + 7) & ~size_t(7);
        memset(p + r.bytes, 0, padded - r.bytes);
        p += padded;
    }
    const uint64_t hdr[5] = {kPackedMagic, (uint64_t)usot_abi_version(), (uint64_t)e->precision, (uint64_t)e->records.size(),
                             packed_checksum(base + kPackedHeader, (size_t)(p - base) - kPackedHeader)};
    memcpy(base, hdr, kPackedHeader);
    return 0;
}

int usot_engine_import_packed(usot_engine* e, const void* host_buf, int64_t size) {
    USOT_REQUIRE(e && host_buf && size >= kPackedHeader, "bad argument");
    std::lock_guard<std::mutex> lk(e->mu);
    const uint8_t* p = static_cast<const uint8_t*>(host_buf);
    uint64_t hdr[5];
    memcpy(hdr, p, kPackedHeader);
    USOT_REQUIRE(hdr[0] == kPackedMagic, "not a usot_b200 packed weight image");
    USOT_REQUIRE(hdr[1] == (uint64_t)usot_abi_version(), "packed weight image was written by another ABI version");
    USOT_REQUIRE(hdr[2] == (uint64_t)e->precision, "packed weight image was written for another precision mode");
    USOT_REQUIRE((size - kPackedHeader) % 8 == 0 && hdr[4] == packed_checksum(p + kPackedHeader, (size_t)size - kPackedHeader),
                 "packed weight image is corrupt (checksum mismatch)");
    e->replay = p + kPackedHeader;
    e->replay_end = p + size;
    int rc = finalize_impl(e);
    if (rc == 0 && e->records.size() != hdr[3]) { set_error("usot_b200: packed weight image has an unexpected record count"); rc = 2; }
    e->replay = e->replay_end = nullptr;
    if (rc) e->finalized = false;
    return rc;
}

int64_t usot_engine_device_bytes(const usot_engine* e) { return e ? e->weight_bytes + (int64_t)e->arena.cap : 0; }

int usot_feature_size(int s) {
    int h1 = (s - 7) / 2 + 1;          // conv1 7x7/2 p0
    int h2 = (h1 + 2 - 3) / 2 + 1;     // maxpool 3x3/2 p1
    return (h2 - 3) / 2 + 1;           // layer2.0 3x3/2 p0 ; layer3 keeps the size
}

int usot_engine_backbone_neck(usot_engine* e, const float* x, int n, int size, float* xf, void* stream) {
    USOT_REQUIRE(x && xf && n > 0 && size >= 63, "bad argument");
    return with_arena(e, (cudaStream_t)stream, [&](Arena& ar) {
        Ctx c{e, ar, (cudaStream_t)stream};
        T f;
        return backbone_neck(c, x, n, size, xf, &f);
    });
}

int usot_engine_template(usot_engine* e, const float* z, int n, int size, const float* template_bbox, float* zf, float* x_ori,
                         void* stream) {
    USOT_REQUIRE(z && zf && n > 0 && size >= 63, "bad argument");
    return with_arena(e, (cudaStream_t)stream, [&](Arena& ar) {
        Ctx c{e, ar, (cudaStream_t)stream};
        T f;
        if (int rc = backbone_neck(c, z, n, size, x_ori, &f)) return rc;
        if (template_bbox) return prroi(c, f, template_bbox, n, zf);
        USOT_REQUIRE(f.h - 8 == 7, "pr_pool=False template needs a 127x127 exemplar (15x15 feature)");
        RUN(launch_center_crop_nhwc(f.f, n, f.h, f.w, 256, 4, zf, c.st));
        return 0;
    });
}

int usot_engine_track(usot_engine* e, const float* x, int n, int size, const float* zf, int nz, const float* template_mem, int nq,
                      float* cls, float* bbox, float* cls_mem, float* xf, void* stream) {
    USOT_REQUIRE(x && zf && cls && bbox && n > 0, "bad argument");
    USOT_REQUIRE(nz == 1 || nz == n, "template batch must be 1 or equal to the search batch");
    USOT_REQUIRE(nq >= 0 && (nq == 0 || (template_mem && cls_mem)), "memory branch needs template_mem and cls_mem");
    USOT_REQUIRE(usot_feature_size(size) >= 9, "search crop too small");
    const cudaStream_t caller = (cudaStream_t)stream;
    cudaStream_t st = caller;
    const int F = usot_feature_size(size), R = F - 6;
    const size_t n_x = (size_t)n * 3 * size * size, n_zf = (size_t)nz * 49 * 256, n_mem = (size_t)n * nq * 49 * 256;
    const size_t n_cls = (size_t)n * R * R, n_xf = (size_t)n * F * F * 256;
    const bool want_graph = g_graph_max_batch > 0 && n <= g_graph_max_batch && !g_prof.on;
    // I/O staging inside the arena (allocated first => fixed addresses): lets a captured graph be replayed for any caller pointers
    struct IO { float *x, *zf, *mem, *cls, *bbox, *cls_mem, *xf; } io{};
    auto body = [&](Arena& ar, bool staged) -> int {
        Ctx c{e, ar, st};
        if (staged) {
            io.x = ar.f(n_x); io.zf = ar.f(n_zf); io.mem = nq ? ar.f(n_mem) : nullptr;
            io.cls = ar.f(n_cls); io.bbox = ar.f(4 * n_cls); io.cls_mem = nq ? ar.f(n_cls) : nullptr; io.xf = ar.f(n_xf);
        }
        const float* xi = staged ? io.x : x;
        const float* zi = staged ? io.zf : zf;
        const float* mi = staged ? io.mem : template_mem;
        T f;
        if (int rc = backbone_neck(c, xi, n, size, staged ? io.xf : xf, &f)) return rc;
        Enc3 cls_x;
        if (int rc = head_offline(c, f, zi, nz, staged ? io.cls : cls, staged ? io.bbox : bbox, &cls_x)) return rc;
        if (nq > 0)
            if (int rc = head_memory(c, cls_x, n, f.h, mi, nq, n * nq, staged ? io.cls_mem : cls_mem)) return rc;
        return 0;
    };
    if (!want_graph) return with_arena(e, st, [&](Arena& ar) { return body(ar, false); });

    usot_engine::GraphEntry* ge = nullptr;
    bool first = false;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        ge = &e->graphs[std::make_tuple(n, size, nz, nq)];  // (std::map: the entry's address is stable)
        first = ge->seen++ == 0;
    }
    if (first)  // first call with this shape runs eagerly (one-time kernel attribute / constant setup must not be captured)
        return with_arena(e, st, [&](Arena& ar) { return body(ar, false); });
    USOT_CUDA_OK(cudaSetDevice(e->device));
    if (!e->gstream) {
        USOT_CUDA_OK(cudaStreamCreateWithFlags(&e->gstream, cudaStreamNonBlocking));
        USOT_CUDA_OK(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
        USOT_CUDA_OK(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
    }
    st = e->gstream;
    USOT_CUDA_OK(cudaEventRecord(e->ev_in, caller));   // everything the caller enqueued (inputs!) precedes the graph
    USOT_CUDA_OK(cudaStreamWaitEvent(st, e->ev_in, 0));
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (int rc = order_begin(e, st)) return rc;   // (outside the capture: the event belongs to no graph)
    }
    if (!ge->exec || ge->arena_gen != e->arena_gen || ge->weights_gen != e->weights_gen) {
        // (re)capture: the planning pass of with_arena sizes / grows the arena, the execute pass is recorded into the graph
        if (ge->exec) { cudaGraphExecDestroy(ge->exec); ge->exec = nullptr; }
        long long before[FAM_COUNT];
        for (int i = 0; i < FAM_COUNT; ++i) before[i] = tl_launches[i];
        bool capturing = false;
        int rc = with_arena(e, st, [&](Arena& ar) -> int {
            if (!ar.plan) {
                if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { set_error("usot_b200: cudaStreamBeginCapture failed"); return 1; }
                capturing = true;
            }
            return body(ar, true);
        }, /*ordered=*/false);
        cudaGraph_t graph = nullptr;
        if (capturing) {
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc == 0 && ce != cudaSuccess) { set_error(std::string("usot_b200: cudaStreamEndCapture: ") + cudaGetErrorString(ce)); rc = 1; }
        }
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        cudaError_t ie = cudaGraphInstantiate(&ge->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { ge->exec = nullptr; set_error(std::string("usot_b200: cudaGraphInstantiate: ") + cudaGetErrorString(ie)); return 1; }
        ge->arena_gen = e->arena_gen;
        ge->weights_gen = e->weights_gen;
        for (int i = 0; i < FAM_COUNT; ++i) { ge->launches[i] = tl_launches[i] - before[i]; count_launches(i, -ge->launches[i]); }  // (capture launched nothing)
    } else {
        // replay: recompute the (deterministic) staging addresses without launching anything
        std::lock_guard<std::mutex> lk(e->mu);
        Arena& ar = e->arena;
        ar.plan = false; ar.off = 0;
        io.x = ar.f(n_x); io.zf = ar.f(n_zf); io.mem = nq ? ar.f(n_mem) : nullptr;
        io.cls = ar.f(n_cls); io.bbox = ar.f(4 * n_cls); io.cls_mem = nq ? ar.f(n_cls) : nullptr; io.xf = ar.f(n_xf);
    }
    USOT_CUDA_OK(cudaMemcpyAsync(io.x, x, n_x * 4, cudaMemcpyDeviceToDevice, st));
    USOT_CUDA_OK(cudaMemcpyAsync(io.zf, zf, n_zf * 4, cudaMemcpyDeviceToDevice, st));
    if (nq) USOT_CUDA_OK(cudaMemcpyAsync(io.mem, template_mem, n_mem * 4, cudaMemcpyDeviceToDevice, st));
    USOT_CUDA_OK(cudaGraphLaunch(ge->exec, st));
    for (int i = 0; i < FAM_COUNT; ++i) count_launches(i, ge->launches[i]);
    USOT_CUDA_OK(cudaMemcpyAsync(cls, io.cls, n_cls * 4, cudaMemcpyDeviceToDevice, st));
    USOT_CUDA_OK(cudaMemcpyAsync(bbox, io.bbox, 4 * n_cls * 4, cudaMemcpyDeviceToDevice, st));
    if (nq) USOT_CUDA_OK(cudaMemcpyAsync(cls_mem, io.cls_mem, n_cls * 4, cudaMemcpyDeviceToDevice, st));
    if (xf) USOT_CUDA_OK(cudaMemcpyAsync(xf, io.xf, n_xf * 4, cudaMemcpyDeviceToDevice, st));
    USOT_CUDA_OK(cudaEventRecord(e->ev_out, st));
    USOT_CUDA_OK(cudaStreamWaitEvent(caller, e->ev_out, 0));  // the caller's stream sees the outputs in order
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (int rc = order_end(e, st)) return rc;
    }
    return 0;
}

int usot_engine_track_frame(usot_engine* e, const uint8_t* frame, int height, int width, int context_xmin, int context_ymin,
                            int original_sz, const uint8_t* fill, int instance_size, const float* zf, const float* mem_buf,
                            const int32_t* mem_rows, int nq, const double* window, double target_w, double target_h, double ratio,
                            double penalty_k, double window_influence, int total_stride, double* result, float* feat_out, void* stream) {
    USOT_REQUIRE(e && e->finalized, "engine not finalized");
    USOT_REQUIRE(frame && fill && zf && mem_buf && mem_rows && window && result && feat_out, "null pointer");
    USOT_REQUIRE(height > 0 && width > 0 && original_sz > 0 && nq > 0 && nq <= 16, "bad argument");
    USOT_REQUIRE(total_stride == 8, "the response grid of this network has stride 8");
    USOT_REQUIRE(target_w > 0 && target_h > 0, "target size must be positive");
    const int F = usot_feature_size(instance_size), R = F - 6;
    USOT_REQUIRE(F >= 9, "search crop too small");
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> frame_lock(e->frame_mu);
    USOT_CUDA_OK(cudaSetDevice(e->device));
    // carve the frame workspace (256-byte aligned pieces)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t a = (off + 255) & ~size_t(255); off = a + bytes; return a; };
    const size_t o_x = take((size_t)3 * instance_size * instance_size * 4), o_mem = take((size_t)nq * 49 * 256 * 4);
    const size_t o_cls = take((size_t)R * R * 4), o_bbox = take((size_t)4 * R * R * 4), o_cm = take((size_t)R * R * 4);
    const size_t o_xf = take((size_t)F * F * 256 * 4), o_box = take(16);
    if (off > e->frame_ws_cap) {
        std::lock_guard<std::mutex> lk(e->mu);
        USOT_CUDA_OK(cudaDeviceSynchronize());
        if (e->frame_ws) USOT_CUDA_OK(cudaFree(e->frame_ws));
        e->frame_ws = nullptr;
        e->frame_ws_cap = 0;
        void* p = nullptr;
        USOT_CUDA_OK(cudaMalloc(&p, off));
        e->frame_ws = static_cast<char*>(p);
        e->frame_ws_cap = off;
    }
    {   // the previous call of this engine (possibly issued on another stream) may still be reading the workspace
        std::lock_guard<std::mutex> lk(e->mu);
        if (int rc = order_begin(e, st)) return rc;
    }
    char* ws = e->frame_ws;
    float *x = reinterpret_cast<float*>(ws + o_x), *mem = reinterpret_cast<float*>(ws + o_mem), *cls = reinterpret_cast<float*>(ws + o_cls);
    float *bbox = reinterpret_cast<float*>(ws + o_bbox), *cm = reinterpret_cast<float*>(ws + o_cm), *xf = reinterpret_cast<float*>(ws + o_xf);
    float* box = reinterpret_cast<float*>(ws + o_box);
    count_launches(FAM_OTHER, 4);
    if (int rc = launch_crop_resize_one(frame, height, width, context_xmin, context_ymin, original_sz, fill, instance_size, x, st)) return rc;
    if (int rc = launch_gather_rows(mem_buf, mem_rows, nq, (size_t)49 * 256, mem, st)) return rc;
    if (int rc = usot_engine_track(e, x, 1, instance_size, zf, 1, mem, nq, cls, bbox, cm, xf, stream)) return rc;
    if (int rc = launch_tracker_post(cls, cm, bbox, window, R, instance_size, target_w, target_h, (float)ratio, penalty_k, window_influence,
                                     result, st)) return rc;
    if (int rc = launch_pool_box_from_result(result, R, instance_size, total_stride, box, st)) return rc;
    return usot_engine_extract_memory_feature(e, nullptr, 1, 0, xf, F, box, feat_out, stream);
}

int usot_engine_extract_memory_feature(usot_engine* e, const float* ori_x, int n, int size, const float* xf, int feat,
                                       const float* search_bbox, float* out, void* stream) {
    USOT_REQUIRE((ori_x != nullptr) != (xf != nullptr), "exactly one of ori_x / xf must be given");
    USOT_REQUIRE(search_bbox && out && n > 0, "bad argument");
    return with_arena(e, (cudaStream_t)stream, [&](Arena& ar) {
        Ctx c{e, ar, (cudaStream_t)stream};
        T f = wrap_f32(xf, n, feat, feat, 256);
        if (ori_x)
            if (int rc = backbone_neck(c, ori_x, n, size, nullptr, &f)) return rc;
        return prroi(c, f, search_bbox, n, out);
    });
}

int usot_engine_forward_train(usot_engine* e, const float* zf, const float* xf, const float* xf_mem, int n, int m, int feat,
                              const float* label, const float* reg_target, const float* reg_weight, const float* search_bbox,
                              float cls_ratio, float* losses, float* backward_map, float* pool_box, void* stream) {
    USOT_REQUIRE(zf && xf && label && reg_target && reg_weight && losses && n > 0 && m >= 0, "bad argument");
    USOT_REQUIRE(m == 0 || (xf_mem && search_bbox), "cycle memory needs xf_mem and search_bbox");
    USOT_REQUIRE(feat >= 9, "feature map too small");
    return with_arena(e, (cudaStream_t)stream, [&](Arena& ar) {
        Ctx c{e, ar, (cudaStream_t)stream};
        const int R = feat - 6, cells = R * R;
        T xf_t = wrap_f32(xf, n, feat, feat, 256);
        float* cls1 = ar.f((size_t)n * cells);
        float* bbox1 = ar.f((size_t)n * 4 * cells);
        Enc3 cls_x;
        if (int rc = head_offline(c, xf_t, zf, n, cls1, bbox1, &cls_x)) return rc;                    // models.py:225
        RUN(launch_iou(bbox1, reg_target, reg_weight, n, cells, losses + 2, c.st));                  // :228
        RUN(launch_bce(cls1, label, n * cells, losses + 0, c.st));                                   // :230
        if (m == 0) {
            if (!ar.plan) USOT_CUDA_OK(cudaMemsetAsync(losses + 1, 0, sizeof(float), c.st));
            return 0;
        }
        T xfm_t = wrap_f32(xf_mem, n * m, feat, feat, 256);
        float* spf = ar.f((size_t)n * 49 * 256);
        if (int rc = prroi(c, xf_t, search_bbox, n, spf)) return rc;                                   // :240
        float* off_cls = ar.f((size_t)n * m * cells);
        float* off_bbox = ar.f((size_t)n * m * 4 * cells);
        Enc3 fwd_x;
        if (int rc = head_offline(c, xfm_t, zf, n, off_cls, off_bbox, &fwd_x)) return rc;             // :253 (zf shared by the m frames)
        float* mem_cls = ar.f((size_t)n * m * cells);
        if (int rc = head_memory(c, fwd_x, n * m, feat, spf, 1, n, mem_cls)) return rc;               // :256-259
        float* pbox = pool_box ? pool_box : ar.f((size_t)n * m * 4);
        RUN(launch_cycle_glue(off_cls, mem_cls, off_bbox, n * m, R, 255 + (feat - 31) * 8, R, cls_ratio, pbox, nullptr, nullptr, c.st));  // :265-274
        float* pooled = ar.f((size_t)n * m * 49 * 256);
        if (int rc = prroi(c, xfm_t, pbox, n * m, pooled)) return rc;                                  // :277
        float* back = backward_map ? backward_map : ar.f((size_t)n * cells);
        if (int rc = head_memory(c, cls_x, n, feat, pooled, m, n * m, back)) return rc;               // :279-281
        RUN(launch_bce(back, label, n * cells, losses + 1, c.st));                                   // :284
        return 0;
    });
}

int usot_tracker_postprocess(const float* cls, const float* cls_mem, const float* bbox, const double* window, int score_size,
                             int instance_size, double target_w, double target_h, double ratio, double penalty_k,
                             double window_influence, double* result, void* stream) {
    USOT_REQUIRE(cls && cls_mem && bbox && window && result, "null pointer");
    USOT_REQUIRE(target_w > 0 && target_h > 0, "target size must be positive");
    count_launches(FAM_OTHER, 1);
    return launch_tracker_post(cls, cls_mem, bbox, window, score_size, instance_size, target_w, target_h, (float)ratio, penalty_k,
                               window_influence, result, (cudaStream_t)stream);
}

}  // extern "C"
