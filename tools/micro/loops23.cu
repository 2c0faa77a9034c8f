// Microbenchmark of the k23_rc inner loops in isolation (sm_100a, 16 warps per SM = 4 per SMSP): cycles per loop iteration
// per SMSP for the packed / scalar formulations, to pick the cheapest instruction forms.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../lc2is_b200/csrc/k2_strip.cuh"
using namespace lc2is;
__constant__ float2 J2c[8] = {{0.f, 1.f}, {2.f, 3.f}, {4.f, 5.f}, {6.f, 7.f}, {8.f, 9.f}, {10.f, 11.f}, {12.f, 13.f}, {14.f, 15.f}};
__device__ long long g_cyc;
__device__ __forceinline__ float fmax3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// MODE 0: pass A only  1: argmax only (FFMA2 + FMNMX3)  2: both  3: argmax with scalar FFMA + FMNMX3  4: argmax FFMA2 + FMNMX (2-input, one class)
//      5: pass A scalar
template <int MODE> __global__ void rowk(float* out, int iters, const float* in) {
    float2 S2[8]; float best[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) S2[k] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 16; ++k) best[k] = -1e30f;
    float Ea = in[threadIdx.x], ra = in[threadIdx.x + 1], Eb = in[threadIdx.x + 2], rb = in[threadIdx.x + 3];
    float da = in[threadIdx.x + 4], db = in[threadIdx.x + 5], v0a = in[threadIdx.x + 6], v0b = in[threadIdx.x + 7];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float2 e2a = make_float2(Ea, Ea * ra), e2b = make_float2(Eb, Eb * rb);
        const float2 r2a = bc2(ra * ra), r2b = bc2(rb * rb);
        const float2 da2 = bc2(da), db2 = bc2(db), va0 = bc2(v0a), vb0 = bc2(v0b);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            if (MODE == 0 || MODE == 2) {
                S2[jj] = fadd2(S2[jj], fadd2(e2a, e2b));
                if (jj < 7) { e2a = fmul2(e2a, r2a); e2b = fmul2(e2b, r2b); }
            }
            if (MODE == 5) {
                S2[jj].x += e2a.x + e2b.x; S2[jj].y += e2a.y + e2b.y;
                if (jj < 7) { e2a.x *= r2a.x; e2a.y *= r2a.x; e2b.x *= r2b.x; e2b.y *= r2b.x; }
            }
            if (MODE == 1 || MODE == 2) {
                const float2 va = ffma2(J2c[jj], da2, va0), vb = ffma2(J2c[jj], db2, vb0);
                best[2 * jj] = fmax3f(va.x, vb.x, best[2 * jj]);
                best[2 * jj + 1] = fmax3f(va.y, vb.y, best[2 * jj + 1]);
            }
            if (MODE == 3) {
                const float vax = fmaf((float)(2 * jj), da, v0a), vay = fmaf((float)(2 * jj + 1), da, v0a);
                const float vbx = fmaf((float)(2 * jj), db, v0b), vby = fmaf((float)(2 * jj + 1), db, v0b);
                best[2 * jj] = fmax3f(vax, vbx, best[2 * jj]);
                best[2 * jj + 1] = fmax3f(vay, vby, best[2 * jj + 1]);
            }
            if (MODE == 4) {
                const float2 va = ffma2(J2c[jj], da2, va0), vb = ffma2(J2c[jj], db2, vb0);
                best[2 * jj] = fmaxf(va.x, best[2 * jj]); best[2 * jj] = fmaxf(vb.x, best[2 * jj]);
                best[2 * jj + 1] = fmaxf(va.y, best[2 * jj + 1]); best[2 * jj + 1] = fmaxf(vb.y, best[2 * jj + 1]);
            }
        }
        Ea += 1e-7f; Eb -= 1e-7f; da += 1e-7f; db -= 1e-7f;      // keep the iterations from being hoisted
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += S2[k].x + S2[k].y + best[2 * k] + best[2 * k + 1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0);
    if (threadIdx.x == 0 && blockIdx.x == 0) g_cyc = t1 - t0;
}

// Horner sweep of the class phase: MODE 0 packed NB=1 (2 chains)  1 packed NB=2 (4 chains)  2 scalar NB=1  3 scalar NB=2
template <int MODE> __global__ void hornk(float* out, int iters, const float* in) {
    extern __shared__ float4 U4[];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) U4[i] = make_float4(in[i], in[i + 1], in[i + 2], in[i + 3]);
    __syncthreads();
    constexpr int NB = (MODE & 1) ? 2 : 1;
    float2 r2a[NB], r2b[NB], acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { r2a[b] = make_float2(in[threadIdx.x + b], in[threadIdx.x + 8 + b]); r2b[b] = make_float2(in[threadIdx.x + 16 + b], in[threadIdx.x + 24 + b]); acc[b] = make_float2(0.f, 0.f); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float2 ha[NB], hb[NB], dA[NB], dB[NB];
        float4 u = U4[15 * 4];
#pragma unroll
        for (int b = 0; b < NB; ++b) { ha[b] = make_float2(u.x, u.y); hb[b] = make_float2(u.z, u.w); dA[b] = ha[b]; dB[b] = hb[b]; }
#pragma unroll
        for (int j = 14; j >= 0; --j) {
            u = U4[j * 4];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (MODE < 2) {
                    dA[b] = ffma2(dA[b], r2a[b], ha[b]); dB[b] = ffma2(dB[b], r2b[b], hb[b]);
                    ha[b] = ffma2(ha[b], r2a[b], make_float2(u.x, u.y)); hb[b] = ffma2(hb[b], r2b[b], make_float2(u.z, u.w));
                } else if (MODE == 4) {
                    dA[b] = ffma2(dA[b], r2a[b], ha[b]); dB[b] = ffma2(dB[b], r2b[b], hb[b]);
                    ha[b].x = fmaf(ha[b].x, r2a[b].x, u.x); ha[b].y = fmaf(ha[b].y, r2a[b].y, u.y);
                    hb[b].x = fmaf(hb[b].x, r2b[b].x, u.z); hb[b].y = fmaf(hb[b].y, r2b[b].y, u.w);
                } else {
                    dA[b].x = fmaf(dA[b].x, r2a[b].x, ha[b].x); dA[b].y = fmaf(dA[b].y, r2a[b].y, ha[b].y);
                    dB[b].x = fmaf(dB[b].x, r2b[b].x, hb[b].x); dB[b].y = fmaf(dB[b].y, r2b[b].y, hb[b].y);
                    ha[b].x = fmaf(ha[b].x, r2a[b].x, u.x); ha[b].y = fmaf(ha[b].y, r2a[b].y, u.y);
                    hb[b].x = fmaf(hb[b].x, r2b[b].x, u.z); hb[b].y = fmaf(hb[b].y, r2b[b].y, u.w);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) { acc[b] = fadd2(acc[b], fadd2(fadd2(ha[b], hb[b]), fadd2(dA[b], dB[b]))); r2a[b].x += 1e-7f; r2a[b].y += 2e-7f; r2b[b].x -= 3e-7f; r2b[b].y -= 1e-7f; }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) s += acc[b].x + acc[b].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) g_cyc = t1 - t0;
}
template <class K> void run(const char* name, K kern, int wps, double work, size_t smem = 0) {
    float *d, *in; cudaMalloc(&d, 148 * 512 * 4 + 64); cudaMalloc(&in, 4096 * 4);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 0.9f + 1e-4f * (i % 97);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int iters = 4000;
    kern<<<148, wps * 128, smem>>>(d, 10, in); cudaDeviceSynchronize();
    kern<<<148, wps * 128, smem>>>(d, iters, in); cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpyFromSymbol(&c, g_cyc, 8);
    printf("%-46s warps/SMSP=%d: %7.1f cyc/iter/warp  %7.1f cyc/iter/SMSP  (%.2f cyc per unit/SMSP) %s\n", name, wps, (double)c / iters, (double)c / iters / wps,
           (double)c / iters / wps / work, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(in);
}
int main() {
    for (int w : {1, 4}) {
        run("row: pass A packed (per pair)", rowk<0>, w, 1);
        run("row: pass A scalar", rowk<5>, w, 1);
        run("row: argmax FFMA2+FMNMX3", rowk<1>, w, 1);
        run("row: argmax FFMA scalar+FMNMX3", rowk<3>, w, 1);
        run("row: argmax FFMA2+FMNMX(2-in)", rowk<4>, w, 1);
        run("row: pass A + argmax (kernel form)", rowk<2>, w, 1);
        run("class: Horner packed NB=1 (per 16 cols x 4 rows)", hornk<0>, w, 1, 1024);
        run("class: Horner packed NB=2", hornk<1>, w, 2, 1024);
        run("class: Horner scalar NB=1", hornk<2>, w, 1, 1024);
        run("class: Horner scalar NB=2", hornk<3>, w, 2, 1024);
        run("class: Horner mixed (d packed, h scalar) NB=1", hornk<4>, w, 1, 1024);
    }
    return 0;
}
