// Microbenchmark: FP64 DMMA (mma.sync m8n8k4) vs DFMA issue throughput on sm_100a.
// Establishes the FP64 "tensor" roofline denominator that MEASURED_PEAKS.json lacks.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void dmma_kernel(double* out, double a, double b, int iters){
  double c[ILP][2];
  #pragma unroll
  for(int j=0;j<ILP;j++){c[j][0]=0;c[j][1]=0;}
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c[j][0]+c[j][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int ILP>
__global__ void dfma_kernel(double* out, double a, double b, int iters){
  double c[ILP];
  #pragma unroll
  for(int j=0;j<ILP;j++) c[j]=j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++)
      asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[j]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mixed: per loop ILP DMMAs and MF DFMAs
template<int ILP, int MF>
__global__ void mixed_kernel(double* out, double a, double b, int iters){
  double c[ILP][2]; double f[MF];
  #pragma unroll
  for(int j=0;j<ILP;j++){c[j][0]=0;c[j][1]=0;}
  #pragma unroll
  for(int j=0;j<MF;j++) f[j]=j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++){
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
      if (j < MF) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[j]) : "d"(a), "d"(b));
    }
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c[j][0]+c[j][1];
  #pragma unroll
  for(int j=0;j<MF;j++) s+=f[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F>
float timeit(F f){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<5;r++){
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best) best=ms;
  }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("device %s sms=%d clockRate=%d kHz\n", p.name, sms, clk);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*1024*4));
  int iters=20000;
  int warps_list[]={1,2,4,8,16,32};
  printf("# DMMA m8n8k4: flops = 2*8*8*4 = 512 per warp-instr\n");
  for(int wi=0; wi<6; wi++){
    int w=warps_list[wi]; int threads=w*32;
    #define RUN_DMMA(ILP) { float ms=timeit([&]{dmma_kernel<ILP><<<sms,threads>>>(out,1.0,1.0,iters);}); \
       double fl=(double)sms*w*iters*ILP*512.0; printf("dmma warps/SM=%2d ILP=%d  %.3f ms  %.2f TF/s  (%.2f clk/DMMA/SMSP @%.0fMHz nominal)\n", w, ILP, ms, fl/ms/1e9, ms*1e-3*clk*1e3/((double)iters*ILP*((w+3)/4)), clk/1e3); }
    RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8)
  }
  printf("# DFMA: flops = 64 per warp-instr\n");
  for(int wi=0; wi<6; wi++){
    int w=warps_list[wi]; int threads=w*32;
    #define RUN_DFMA(ILP) { float ms=timeit([&]{dfma_kernel<ILP><<<sms,threads>>>(out,1.0000001,1e-9,iters);}); \
       double fl=(double)sms*w*iters*ILP*64.0; printf("dfma warps/SM=%2d ILP=%d  %.3f ms  %.2f TF/s\n", w, ILP, ms, fl/ms/1e9); }
    RUN_DFMA(1) RUN_DFMA(4) RUN_DFMA(8)
  }
  printf("# mixed DMMA+DFMA (8 DMMA + MF DFMA per iter)\n");
  for(int wi=2; wi<6; wi++){
    int w=warps_list[wi]; int threads=w*32;
    #define RUN_MIX(MF) { float ms=timeit([&]{mixed_kernel<8,MF><<<sms,threads>>>(out,1.0000001,1e-9,iters);}); \
       double fl=(double)sms*w*iters*(8*512.0+MF*64.0); printf("mixed warps/SM=%2d MF=%d  %.3f ms  %.2f TF/s total (dmma part %.2f)\n", w, MF, ms, fl/ms/1e9, (double)sms*w*iters*8*512.0/ms/1e9); }
    RUN_MIX(2) RUN_MIX(8)
  }
  return 0;
}
