// FP64 peak microbenchmarks for B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate,
// their interaction (shared pipe or not), CUDA log() throughput, and cuBLAS DGEMM.
// Output: one JSON object on stdout.  Test/measurement tooling, not product code.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template<int NMMA, int NFMA>
__global__ void __launch_bounds__(512) mix_kernel(double* out, int iters, double seed){
  double c[16][2]; double f[16];
#pragma unroll
  for(int i=0;i<16;i++){c[i][0]=seed*i;c[i][1]=seed+i;f[i]=seed*(i+threadIdx.x);}
  double a=seed+threadIdx.x*1e-9, b=seed*0.5+threadIdx.x*1e-9;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int i=0;i<16;i++){
      if(i<NMMA) dmma(c[i][0],c[i][1],a,b);
      if(i<NFMA) asm volatile("fma.rn.f64 %0,%1,%2,%0;":"+d"(f[i]):"d"(a),"d"(b));
    }
  }
  double s=0;
#pragma unroll
  for(int i=0;i<16;i++) s+=c[i][0]+c[i][1]+f[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

__global__ void __launch_bounds__(512) log_kernel(double* out, int iters, double seed){
  double x[4]; for(int i=0;i<4;i++) x[i]=seed+threadIdx.x*1e-3+i;
  double s=0;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int i=0;i<4;i++){ s+=log(x[i]); x[i]+=1e-3; }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void __launch_bounds__(512) sqrt_kernel(double* out, int iters, double seed){
  double x[4]; for(int i=0;i<4;i++) x[i]=seed+threadIdx.x*1e-3+i;
  double s=0;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int i=0;i<4;i++){ s+=sqrt(x[i]); x[i]+=1e-3; }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float time_ms(F f, int rep=3){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<rep;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}

template<int NMMA,int NFMA> void run_mix(double* out,int sms,int ctas_per_sm,int threads,const char* name,bool last=false){
  int iters=20000; int grid=sms*ctas_per_sm;
  float ms=time_ms([&]{ mix_kernel<NMMA,NFMA><<<grid,threads>>>(out,iters,1.0000001); });
  double warps=(double)grid*threads/32;
  double mma_flops=warps*iters*(double)NMMA*512.0;       // 8x8x4 FMA = 256 FMA = 512 flop
  double fma_flops=warps*iters*(double)NFMA*32*2.0;
  printf("  \"%s\": {\"ms\": %.3f, \"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"threads\": %d, \"ctas_per_sm\": %d}%s\n",
         name, ms, mma_flops/ms/1e9, fma_flops/ms/1e9, threads, ctas_per_sm, last?"":",");
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*1024));
  printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
  run_mix<16,0>(out,sms,1,512,"dmma_only_16w");
  run_mix<16,0>(out,sms,2,512,"dmma_only_32w");
  run_mix<16,0>(out,sms,1,256,"dmma_only_8w");
  run_mix<16,0>(out,sms,1,128,"dmma_only_4w");
  run_mix<8,0>(out,sms,1,128,"dmma_only_4w_ilp8");
  run_mix<4,0>(out,sms,1,128,"dmma_only_4w_ilp4");
  run_mix<2,0>(out,sms,1,128,"dmma_only_4w_ilp2");
  run_mix<1,0>(out,sms,1,128,"dmma_only_4w_ilp1");
  run_mix<0,16>(out,sms,1,512,"dfma_only_16w");
  run_mix<0,16>(out,sms,2,512,"dfma_only_32w");
  run_mix<16,2>(out,sms,1,512,"mix_16mma_2fma");
  run_mix<16,4>(out,sms,1,512,"mix_16mma_4fma");
  run_mix<16,8>(out,sms,1,512,"mix_16mma_8fma");
  run_mix<16,16>(out,sms,1,512,"mix_16mma_16fma");
  run_mix<8,16>(out,sms,1,512,"mix_8mma_16fma");
  {
    int iters=4000; int grid=sms*2;
    float ms=time_ms([&]{ log_kernel<<<grid,512>>>(out,iters,1.5); });
    printf("  \"log_f64\": {\"ms\": %.3f, \"gevals_per_s\": %.3f},\n", ms, (double)grid*512*iters*4/ms/1e6);
    ms=time_ms([&]{ sqrt_kernel<<<grid,512>>>(out,iters,1.5); });
    printf("  \"sqrt_f64\": {\"ms\": %.3f, \"gevals_per_s\": %.3f},\n", ms, (double)grid*512*iters*4/ms/1e6);
  }
  {
    cublasHandle_t h; cublasCreate(&h);
    for(int n : {4096, 8192}){
      double *A,*B,*C; size_t bytes=sizeof(double)*n*n;
      CK(cudaMalloc(&A,bytes)); CK(cudaMalloc(&B,bytes)); CK(cudaMalloc(&C,bytes));
      CK(cudaMemset(A,0,bytes)); CK(cudaMemset(B,0,bytes));
      double one=1,zero=0;
      float ms=time_ms([&]{ cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n); },5);
      printf("  \"cublas_dgemm_%d\": {\"ms\": %.3f, \"tflops\": %.3f},\n", n, ms, 2.0*n*n*n/ms/1e9);
      // sustained: back to back for ~3 s
      int reps=(int)(3000.0f/ms)+1; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); for(int r=0;r<reps;r++) cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float tot; cudaEventElapsedTime(&tot,e0,e1);
      printf("  \"cublas_dgemm_%d_sustained\": {\"ms\": %.3f, \"tflops\": %.3f, \"reps\": %d},\n", n, tot/reps, 2.0*n*n*n/(tot/reps)/1e9, reps);
      cudaFree(A);cudaFree(B);cudaFree(C);
    }
    cublasDestroy(h);
  }
  run_mix<16,0>(out,sms,1,512,"dmma_only_16w_again",true);
  printf("}\n");
  return 0;
}
