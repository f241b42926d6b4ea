// Does interleaving FP64 ALU ops between DMMAs cost pipe time beyond their own occupancy?  Measurement tooling.
// 64 DMMAs (4 k4-steps of a 4x4 register tile, fragments from shared memory) + NF dependent DFMAs per
// warp-iteration, __syncthreads per iteration; the DFMAs are placed (asm volatile keeps source order)
//   MODE 0: none        MODE 1: one burst in front of the DMMAs      MODE 2: 1 DFMA after every (64/NF)th DMMA
//   MODE 3: bursts of 6 after each k4-step
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma(double &f, double a, double b){ asm volatile("fma.rn.f64 %0,%0,%1,%2;" : "+d"(f) : "d"(a), "d"(b)); }
template<int MODE, int NF, bool BAR>
__global__ void __launch_bounds__(512,1) k(double* out, int iters, double seed){
  extern __shared__ double sm[];
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5;
  for(int i=tid;i<16384;i+=blockDim.x) sm[i]=seed*(i%97)*1e-3;
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++){acc[i][j][0]=seed*i;acc[i][j][1]=seed*j;}
  double a[4],b[4]; double f=seed*lane;
  const double* pa=sm+lane, *pb=sm+8192+(warp%16)*128+lane;
  for(int it=0;it<iters;it++){
    if(MODE==1){
#pragma unroll
      for(int r=0;r<NF;r++) dfma(f,1.0000001,seed);
    }
    int cnt=0;
#pragma unroll
    for(int q=0;q<4;q++){
#pragma unroll
      for(int i=0;i<4;i++){a[i]=pa[((it*4+q)&15)*128+i*32]; b[i]=pb[((it*4+q)&3)*2048+i*32];}
#pragma unroll
      for(int i=0;i<4;i++)
#pragma unroll
        for(int j=0;j<4;j++){
          dmma(acc[i][j][0],acc[i][j][1],a[i],b[j]);
          ++cnt;
          if(MODE==2 && NF>0 && (cnt % (64/NF))==0) dfma(f,1.0000001,seed);
        }
      if(MODE==3){
#pragma unroll
        for(int r=0;r<NF/4;r++) dfma(f,1.0000001,seed);
      }
    }
    if(BAR) __syncthreads();
  }
  double s=f;
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++) s+=acc[i][j][0]+acc[i][j][1];
  out[blockIdx.x*blockDim.x+tid]=s;
}
template<int MODE,int NF,bool BAR> void run(double* out,int sms,const char* name){
  int iters=4000; size_t smem=16384*8;
  CK(cudaFuncSetAttribute(k<MODE,NF,BAR>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE,NF,BAR><<<sms,512,smem>>>(out,iters,1.0000001); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<3;r++){ cudaEventRecord(e0); k<MODE,NF,BAR><<<sms,512,smem>>>(out,iters,1.0000001); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  double flops=(double)sms*16*iters*64.0*512.0;
  printf("  \"%s\": {\"ms\": %.3f, \"dmma_tflops\": %.3f},\n",name,best,flops/best/1e9);
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out,sizeof(double)*sms*1024));
  printf("{\n");
  run<0,0,false>(out,sms,"none_nobar"); run<0,0,true>(out,sms,"none_bar");
  run<1,16,true>(out,sms,"front16_bar"); run<2,16,true>(out,sms,"every4th_16_bar"); run<3,16,true>(out,sms,"perq4_16_bar");
  run<1,32,true>(out,sms,"front32_bar"); run<2,32,true>(out,sms,"every2nd_32_bar"); run<3,32,true>(out,sms,"perq8_32_bar");
  run<1,32,false>(out,sms,"front32_nobar"); run<2,32,false>(out,sms,"every2nd_32_nobar"); run<3,32,false>(out,sms,"perq8_32_nobar");
  run<1,64,true>(out,sms,"front64_bar"); run<2,64,true>(out,sms,"every1_64_bar");
  printf("  \"end\": 0\n}\n");
  return 0;
}
