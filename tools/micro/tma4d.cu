// probe: 4-D fp32 TMA box {4,2,C,1} with negative / out-of-bounds start coordinates (what k23_rc stages)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int C, int x, int y, int n, int mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* cells = (float*)sm;
    uint64_t* bar = (uint64_t*)(sm + 8192);
    const int lane = threadIdx.x & 31;
    if (mode == 0 ? (threadIdx.x == 0) : true) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(C * 32) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(s32(cells)), "l"(&tm), "r"(s32(bar)), "r"(x), "r"(y), "r"(0), "r"(n) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
    }
    for (int i = lane; i < C * 8; i += 32) out[i] = cells[i];
}
int main(int argc, char** argv) {
    int B = 2, C = argc > 1 ? atoi(argv[1]) : 150, h = 32, w = 32;
    int x = argc > 2 ? atoi(argv[2]) : -1, y = argc > 3 ? atoi(argv[3]) : -1;
    float* d; size_t n = (size_t)B * C * h * w;
    cudaMalloc(&d, n * 4);
    float* hbuf = (float*)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) hbuf[i] = (float)(i % 100003);
    cudaMemcpy(d, hbuf, n * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4, (cuuint64_t)w * h * C * 4};
    cuuint32_t box[4] = {4, 2, (cuuint32_t)C, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((PFN_enc)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    float* o; cudaMalloc(&o, C * 8 * 4);
    k<<<1, 32, 8192 + 64>>>(tm, o, C, x, y, 1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        float* ho = (float*)malloc(C * 8 * 4);
        cudaMemcpy(ho, o, C * 8 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < C; ++c) for (int r2 = 0; r2 < 2; ++r2) for (int cx = 0; cx < 4; ++cx) {
            int yy = y + r2, xx = x + cx;
            float ref = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? hbuf[(((size_t)1 * C + c) * h + yy) * w + xx] : 0.f;
            if (ho[c * 8 + r2 * 4 + cx] != ref) ++bad;
        }
        printf("mismatches %d\n", bad);
    }
    return 0;
}
