/* Minimal C host of the usot_b200 boundary (no Python, no torch): build the engine from raw fp32 state_dict tensors, run one
 * template + track call, read the response maps back.  It shows that include/usot_b200.h is a plain C interface; the Python
 * package (usot_b200/_lib.py) binds exactly these symbols with ctypes.
 *
 *   gcc -std=c99 -I include examples/c_abi_demo.c -o c_abi_demo -L usot_b200 -lusot_b200 -L /usr/local/cuda/lib64 -lcudart
 *
 * Weights: a directory of <state_dict key>.f32 files (raw little-endian fp32, reference shapes), e.g. written with
 *   for k, v in net.state_dict().items(): v.float().numpy().tofile(f"{dir}/{k}.f32")
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "usot_b200.h"

/* the CUDA runtime entry points this demo needs (declared here so that the file compiles without the CUDA headers) */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
enum { H2D = 1, D2H = 2 };

#define CHECK(call)                                                              \
    do {                                                                         \
        if ((call) != 0) {                                                       \
            fprintf(stderr, "%s failed: %s\n", #call, usot_last_error());        \
            return 1;                                                            \
        }                                                                        \
    } while (0)

static int load_dir(usot_engine* e, const char* dir, const char* const* keys, const long* numel, int n) {
    for (int i = 0; i < n; ++i) {
        char path[1024];
        snprintf(path, sizeof(path), "%s/%s.f32", dir, keys[i]);
        FILE* f = fopen(path, "rb");
        if (!f) { fprintf(stderr, "missing %s\n", path); return 1; }
        float* buf = (float*)malloc((size_t)numel[i] * 4);
        if (fread(buf, 4, (size_t)numel[i], f) != (size_t)numel[i]) { fclose(f); free(buf); return 1; }
        fclose(f);
        CHECK(usot_engine_load_tensor(e, keys[i], buf, numel[i]));
        free(buf);
    }
    return 0;
}

int main(int argc, char** argv) {
    printf("usot_b200 ABI version %d, feature size of a 255 crop: %d\n", usot_abi_version(), usot_feature_size(255));
    if (argc < 2) {
        printf("usage: %s <weights dir> (see the header comment); without it only the ABI is probed\n", argv[0]);
        return 0;
    }
    usot_engine* e = NULL;
    CHECK(usot_engine_create(&e, 0, USOT_PREC_FP16X3_TC));
    /* a real host walks the 443 floating tensors of the state_dict; two are listed to keep the demo short */
    static const char* const keys[] = {"features.features.conv1.weight", "connect_model.adjust"};
    static const long numel[] = {64 * 3 * 7 * 7, 1};
    if (load_dir(e, argv[1], keys, numel, 2)) return 1;
    if (usot_engine_finalize(e) != 0) {  /* fails loudly while tensors are missing: there is no fallback */
        fprintf(stderr, "finalize: %s\n", usot_last_error());
        usot_engine_destroy(e);
        return 2;
    }
    const int S = 255, F = usot_feature_size(S), R = F - 6;
    float *z, *x, *zf, *cls, *bbox;
    cudaMalloc((void**)&z, 3 * 127 * 127 * 4);
    cudaMalloc((void**)&x, (size_t)3 * S * S * 4);
    cudaMalloc((void**)&zf, 49 * 256 * 4);
    cudaMalloc((void**)&cls, (size_t)R * R * 4);
    cudaMalloc((void**)&bbox, (size_t)4 * R * R * 4);
    CHECK(usot_engine_template(e, z, 1, 127, NULL, zf, NULL, NULL));
    CHECK(usot_engine_track(e, x, 1, S, zf, 1, NULL, 0, cls, bbox, NULL, NULL, NULL));
    cudaDeviceSynchronize();
    float first;
    cudaMemcpy(&first, cls, 4, D2H);
    printf("cls[0] = %g\n", first);
    cudaFree(z); cudaFree(x); cudaFree(zf); cudaFree(cls); cudaFree(bbox);
    usot_engine_destroy(e);
    return 0;
}
