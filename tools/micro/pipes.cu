// Microbenchmark (sm_100a): issue cost of the packed / scalar fp32 forms the k23_rc loops are made of, per SMSP, for 1..8
// warps per SMSP; and the dependent-issue latency of each.  Prints cycles per warp-instruction per SMSP.
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ float lo(u64 p){float a,b; asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(p)); return a+b;}

// MODE: 0 FFMA scalar (3 distinct regs) 1 FFMA2 (3 distinct pairs) 2 FMUL2 3 FADD2 4 FMNMX3 5 FFMA2 with broadcast scalar operand
//       6 FFMA scalar with immediate   7 MUFU.EX2   8 mix FFMA2+FMNMX3 (1:1)   9 FSEL/SEL (alu)
template<int MODE, int NCHAIN> __global__ void k(float* out, int iters, float x, long long* cyc) {
    u64 p[NCHAIN], q[NCHAIN], r[NCHAIN]; float a[NCHAIN], b[NCHAIN], c[NCHAIN];
#pragma unroll
    for (int i=0;i<NCHAIN;++i){a[i]=x+i; b[i]=1.0001f+i*1e-6f; c[i]=0.001f*i; p[i]=pk(x+i,x-i); q[i]=pk(1.0001f+i*1e-6f,0.9999f); r[i]=pk(0.001f*i,0.002f);}
    __syncthreads();
    long long t0 = clock64();
    for (int it=0; it<iters; ++it) {
#pragma unroll
        for (int u=0;u<4;++u)
#pragma unroll
        for (int i=0;i<NCHAIN;++i) {
            if (MODE==0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
            if (MODE==1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q[i]), "l"(r[i]));
            if (MODE==2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
            if (MODE==3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(r[i]));
            if (MODE==4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
            if (MODE==5) { u64 bq = pk(b[i], b[i]); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bq), "l"(r[i])); }
            if (MODE==6) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, %1;" : "+f"(a[i]) : "f"(c[i]));
            if (MODE==7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE==8) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q[i]), "l"(r[i])); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i])); }
            if (MODE==9) asm volatile("{.reg .pred pp; setp.gt.f32 pp, %1, %2; selp.f32 %0, %1, %0, pp;}" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
        }
    }
    long long t1 = clock64();
    float s=0;
#pragma unroll
    for(int i=0;i<NCHAIN;++i){ s+=a[i]+lo(p[i]); }
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    if (threadIdx.x==0 && blockIdx.x==0) *cyc = t1-t0;
}
template<int MODE, int NCHAIN> void run(const char* name, int warps_per_smsp, int instr_per_unit) {
    float* d; long long* dc; cudaMalloc(&d, 148*1024*4); cudaMalloc(&dc, 8);
    int iters=2000; int threads = warps_per_smsp*4*32;
    k<MODE,NCHAIN><<<148,threads>>>(d, 10, 1.f, dc); cudaDeviceSynchronize();
    k<MODE,NCHAIN><<<148,threads>>>(d, iters, 1.f, dc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double n = (double)iters*4*NCHAIN*instr_per_unit;                    // warp-instructions per warp
    printf("%-34s chains=%d warps/SMSP=%d : %.2f cyc per warp-instr per warp, %.2f cyc per instr per SMSP\n", name, NCHAIN, warps_per_smsp,
           c/n, c/(n*warps_per_smsp));
    cudaFree(d); cudaFree(dc);
}
#define RUNALL(M,name,ipu) run<M,1>(name,1,ipu); run<M,8>(name,1,ipu); run<M,8>(name,2,ipu); run<M,8>(name,4,ipu); run<M,2>(name,4,ipu);
int main(){
    RUNALL(0,"FFMA scalar 3-reg",1) RUNALL(6,"FFMA scalar imm",1) RUNALL(1,"FFMA2 3 pairs",1) RUNALL(5,"FFMA2 bcast-scalar operand",1)
    RUNALL(2,"FMUL2",1) RUNALL(3,"FADD2",1) RUNALL(4,"FMNMX3",1) RUNALL(8,"FFMA2+FMNMX3",2) RUNALL(7,"MUFU.EX2",1) RUNALL(9,"FSETP+SEL",2)
    return 0; }
