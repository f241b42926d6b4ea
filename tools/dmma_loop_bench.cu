// Which ingredient of the K1 inner loop costs DMMA issue rate?  Measurement tooling.
//  V0: 4x4 register tile, distinct A/B registers, no loads
//  V1: + LDS.64 fragment loads every k4 step
//  V2: + __syncthreads every 16 k4 steps
//  V3: + 24 dependent DFMA per 64 DMMA (the Phi chain), interleaved
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b){
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template<int V, int NW>
__global__ void __launch_bounds__(NW*32,1) k(double* out, int iters, double seed){
  extern __shared__ double sm[];
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5;
  for(int i=tid;i<16384;i+=blockDim.x) sm[i]=seed*(i%97)*1e-3;
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++){acc[i][j][0]=seed*i;acc[i][j][1]=seed*j;}
  double a[4],b[4];
#pragma unroll
  for(int i=0;i<4;i++){a[i]=seed+i+lane*1e-6;b[i]=seed*0.5+i+lane*1e-6;}
  double f=seed*lane;
  const double* pa=sm+lane, *pb=sm+8192+(warp%16)*128+lane;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int q=0;q<4;q++){
      if(V>=1){
#pragma unroll
        for(int i=0;i<4;i++){a[i]=pa[((it*4+q)&15)*128+i*32]; b[i]=pb[((it*4+q)&3)*2048+i*32];}
      }
#pragma unroll
      for(int i=0;i<4;i++)
#pragma unroll
        for(int j=0;j<4;j++) dmma(acc[i][j][0],acc[i][j][1],a[i],b[j]);
      if(V>=3){
#pragma unroll
        for(int r=0;r<6;r++) f=fma(f,1.0000001,seed);
      }
    }
    if(V>=2) __syncthreads();
  }
  double s=f;
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++) s+=acc[i][j][0]+acc[i][j][1];
  out[blockIdx.x*blockDim.x+tid]=s;
}
template<int V,int NW> void run(double* out,int sms,const char* name){
  int iters=4000; size_t smem=16384*8;
  CK(cudaFuncSetAttribute(k<V,NW>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V,NW><<<sms,NW*32,smem>>>(out,iters,1.0000001); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<3;r++){ cudaEventRecord(e0); k<V,NW><<<sms,NW*32,smem>>>(out,iters,1.0000001); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  double flops=(double)sms*NW*iters*64.0*512.0;
  printf("  \"%s\": {\"ms\": %.3f, \"dmma_tflops\": %.3f},\n",name,best,flops/best/1e9);
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out,sizeof(double)*sms*1024));
  printf("{\n");
  run<0,16>(out,sms,"v0_regs_16w"); run<1,16>(out,sms,"v1_lds_16w"); run<2,16>(out,sms,"v2_lds_bar_16w"); run<3,16>(out,sms,"v3_lds_bar_phi_16w");
  run<0,8>(out,sms,"v0_regs_8w"); run<1,8>(out,sms,"v1_lds_8w"); run<2,8>(out,sms,"v2_lds_bar_8w"); run<3,8>(out,sms,"v3_lds_bar_phi_8w");
  run<0,4>(out,sms,"v0_regs_4w"); run<1,4>(out,sms,"v1_lds_4w"); run<2,4>(out,sms,"v2_lds_bar_4w"); run<3,4>(out,sms,"v3_lds_bar_phi_4w");
  printf("  \"end\": 0\n}\n");
  return 0;
}
