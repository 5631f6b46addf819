// Microbenchmark: achieved MUFU (ex2 / rcp) throughput per SM per clock on this GPU, alone and
// mixed with FP32 work, at several warps per SM.  Build: nvcc -arch=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float rcpa(float x){float y; asm volatile("rcp.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float tanha(float x){float y; asm volatile("tanh.approx.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ unsigned packh2(float lo, float hi){unsigned p; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;":"=r"(p):"f"(hi),"f"(lo)); return p;}
template<int MODE> __global__ void k(float* out, int iters, float seed){
    float x[8];
    #pragma unroll
    for(int i=0;i<8;++i) x[i]=seed+i*0.01f+threadIdx.x*1e-4f;
    for(int it=0;it<iters;++it){
        #pragma unroll
        for(int i=0;i<8;++i){
            if(MODE==0) x[i]=ex2a(x[i]);                       // 1 MUFU
            if(MODE==1) x[i]=rcpa(x[i]);                       // 1 MUFU
            if(MODE==2) x[i]=ex2a(x[i])*0.5f+0.25f;            // 1 MUFU + 1 FFMA
            if(MODE==3) {float e=ex2a(fminf(x[i],30.f)); float d=1.f+e; x[i]=rcpa(d)*0.7f-0.3f;}   // 2 MUFU + 4 fp
            if(MODE==4) {float e=ex2a(fminf(x[i],30.f)); float d=1.f+e; float a=d*0.9f; a=a*1.1f+0.2f; a=a*0.8f-0.1f; a=a*1.05f+0.01f; x[i]=a*0.01f;} // 1 MUFU + 7 fp
            if(MODE==6) x[i]=tanha(x[i]);                      // 1 MUFU (tanh)
            if(MODE==7) x[i]=tanha(x[i]+0.3f)*0.9f+0.1f;       // 1 MUFU (tanh) + 2 fp
            if(MODE==8) x[i]=__uint_as_float(packh2(x[i],x[(i+1)&7])&0x3fff3fffu);  // 1 F2FP + 1 LOP
            if(MODE==9) {float t=tanha(x[i]); x[i]=__uint_as_float(packh2(t,x[(i+1)&7])&0x3fff3fffu);}  // 1 MUFU + 1 F2FP + LOP
            if(MODE==5) {float a=x[i]; a=a*1.1f+0.2f; a=a*0.8f-0.1f; a=a*1.05f+0.01f; a=a*0.9f+0.1f; x[i]=a;} // 4 FFMA only
        }
    }
    float s=0; 
    #pragma unroll
    for(int i=0;i<8;++i) s+=x[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, int mufu_per, int fp_per){
    float* out; cudaMalloc(&out, 148*1024*sizeof(float));
    int dev; cudaGetDevice(&dev); int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    for(int warps: {4,8,16,32}){
        int iters=20000; dim3 grid(148), block(warps*32);
        k<MODE><<<grid,block>>>(out,100,0.5f); cudaDeviceSynchronize();
        cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a); k<MODE><<<grid,block>>>(out,iters,0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms,a,b);
        double ops=(double)iters*8*warps*32;    // per SM element-iterations
        double clks=ms*1e-3*clk*1e3;
        printf("%-28s warps/SM=%2d  %.3f ms  elem/clk/SM=%.2f  MUFU/clk/SM=%.2f  FP/clk/SM=%.1f\n",name,warps,ms,ops/clks,ops*mufu_per/clks,ops*fp_per/clks);
    }
    cudaFree(out);
}
int main(){
    run<0>("ex2 only",1,0); run<1>("rcp only",1,0); run<2>("ex2 + ffma",1,1);
    run<3>("ex2+rcp sigmoid (2M+4F)",2,4); run<4>("1 MUFU + 7 fp",1,7); run<5>("4 ffma only",0,4);
    run<8>("f2fp pack + lop",0,2); run<9>("tanh + f2fp + lop",1,2);
    run<6>("tanh only",1,0); run<7>("tanh + 2 fp",1,2);
    return 0;
}
