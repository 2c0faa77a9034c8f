// Microbenchmark: scalar FFMA vs packed FFMA2 vs MUFU.EX2 issue/throughput on sm_100a.
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
template<int MODE> __global__ void k(float* out, int iters, float x) {
    float a[8]; unsigned long long p[8];
    for (int i=0;i<8;++i){a[i]=x+i; p[i]=pk(x+i,x-i);}
    unsigned long long m = pk(1.0001f,0.9999f), c = pk(0.001f,0.002f);
    for (int it=0; it<iters; ++it) {
#pragma unroll
        for (int i=0;i<8;++i) {
            if (MODE==0) a[i] = fmaf(a[i], 1.0001f, 0.001f);
            if (MODE==1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m), "l"(c));
            if (MODE==2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE==3) { a[i] = fmaf(a[i], 1.0001f, 0.001f); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m), "l"(c)); }
            if (MODE==4) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m), "l"(c)); if (i<2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
        }
    }
    float s=0; for(int i=0;i<8;++i){ s+=a[i]; float lo,hi; asm("mov.b64 {%0,%1}, %2;":"=f"(lo),"=f"(hi):"l"(p[i])); s+=lo+hi;}
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, double ops_per_iter_per_thread) {
    float* d; cudaMalloc(&d, 148*8*256*4);
    int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148*8,256>>>(d, 100, 1.f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148*8,256>>>(d, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double warp_instr = 148.0*8*8*iters*8; // warps * iters * 8 instr (per MODE unit)
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %.3f ms  -> %.2f warp-units/clk/SM (at %d MHz nominal), %.2f Tops/s\n", name, ms, warp_instr/(ms*1e-3)/(clk*1e3)/148, clk/1000, 148.0*8*256*iters*8*ops_per_iter_per_thread/(ms*1e-3)/1e12);
    cudaFree(d);
}
int main(){ run<0>("FFMA scalar",1); run<1>("FFMA2 packed",2); run<2>("MUFU.EX2",1); run<3>("FFMA + FFMA2 interleaved",3); run<4>("FFMA2 + 1/4 MUFU",2.25); return 0; }
