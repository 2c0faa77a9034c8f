// Microbenchmark: the K2 strip kernel's two inner loops on registers only, at several occupancies.
//   A : per class  e chain (15 dependent FMUL2) + 16 FADD2 accumulations          (pass A)
//   A2: same with the chain split into even/odd columns (two chains of 8 with rho^2)
//   B : per class  Horner h/d sweep, 29 FFMA2                                      (pass B)
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ u64 mul2(u64 a, u64 b){u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
template<int MODE> __global__ void k(float* out, int classes, float x) {
    u64 S[16];
    for (int j=0;j<16;++j) S[j]=pk(x*j,x+j);
    u64 e0 = pk(0.5f+x,0.4f+x), r0 = pk(0.999f,1.001f), e1 = pk(0.3f+x,0.2f), r1 = pk(1.0001f,0.9999f);
    for (int c=0;c<classes;c+=2) {
        if (MODE==0) {
            u64 a=e0,b=e1;
#pragma unroll
            for (int j=0;j<16;++j){ S[j]=add2(S[j],a); S[j]=add2(S[j],b); if(j<15){a=mul2(a,r0); b=mul2(b,r1);} }
        }
        if (MODE==1) {
            u64 a=e0,b=e1, a1=mul2(a,r0), b1=mul2(b,r1), ra=mul2(r0,r0), rb=mul2(r1,r1);
#pragma unroll
            for (int j=0;j<16;j+=2){ S[j]=add2(S[j],a); S[j+1]=add2(S[j+1],a1); S[j]=add2(S[j],b); S[j+1]=add2(S[j+1],b1);
                if(j<14){a=mul2(a,ra); a1=mul2(a1,ra); b=mul2(b,rb); b1=mul2(b1,rb);} }
        }
        if (MODE==2) {
            u64 h=S[15], d=h, h2=S[15], d2=h2;
            h=fma2(h,r0,S[14]); h2=fma2(h2,r1,S[14]);
#pragma unroll
            for (int j=13;j>=0;--j){ d=fma2(d,r0,h); h=fma2(h,r0,S[j]); d2=fma2(d2,r1,h2); h2=fma2(h2,r1,S[j]); }
            e0=add2(e0,mul2(h,d)); e1=add2(e1,mul2(h2,d2));
        }
        r0=add2(r0,pk(1e-9f,1e-9f)); r1=add2(r1,pk(-1e-9f,1e-9f));
    }
    float s=0; for(int j=0;j<16;++j){ float lo,hi; asm("mov.b64 {%0,%1}, %2;":"=f"(lo),"=f"(hi):"l"(S[j])); s+=lo+hi;}
    { float lo,hi; asm("mov.b64 {%0,%1}, %2;":"=f"(lo),"=f"(hi):"l"(e0)); s+=lo+hi; asm("mov.b64 {%0,%1}, %2;":"=f"(lo),"=f"(hi):"l"(e1)); s+=lo+hi; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, int warps_per_sm, double packed_per_class) {
    float* d; cudaMalloc(&d, 148*32*32*4);
    int classes=40000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int threads = 32*warps_per_sm;   // one CTA per SM
    dim3 grid(148), block(threads > 1024 ? 1024 : threads);
    if (threads > 1024) { grid.x = 296; block.x = threads/2; }
    k<MODE><<<grid,block>>>(d, 100, 1.f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<grid,block>>>(d, classes, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double packed = (double)warps_per_sm*classes*packed_per_class;   // per SM
    printf("%-10s warps/SM %2d: %.3f ms -> %.3f packed warp-instr/clk/SM (1965 MHz)\n", name, warps_per_sm, ms, packed/(ms*1e-3)/1.965e9);
    cudaFree(d);
}
int main(){
    int ws[] = {4, 8, 12, 16, 20, 24, 32};
    for (int w : ws) run<0>("A", w, 31+1);
    for (int w : ws) run<1>("A2 split", w, 32+2+1);
    for (int w : ws) run<2>("B horner", w, 29+2+1);
    return 0;
}
