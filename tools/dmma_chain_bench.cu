// How does the cost of FP64 ALU work mixed into a DMMA stream depend on its dependency structure and placement?
// Measurement tooling (follow-up to dmma_mix_bench.cu).  64 DMMAs (4 k4-steps of a 4x4 register tile, fragments from
// shared memory) + NF DFMAs per warp-iteration, no CTA barrier; the DFMAs form NCH independent chains (round-robin)
//   MODE 0: none   MODE 1: one burst in front of the DMMAs   MODE 2: evenly spread between the DMMAs
//   MODE 4: one burst behind the DMMAs   MODE 5: second half of the DMMA block + behind it (what ptxas does in K1)
// asm volatile keeps the source order.  Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma(double &f, double a, double b){ asm volatile("fma.rn.f64 %0,%0,%1,%2;" : "+d"(f) : "d"(a), "d"(b)); }
template<int MODE, int NF, int NCH, int WARPS>
__global__ void __launch_bounds__(WARPS*32,1) k(double* out, int iters, double seed){
  extern __shared__ double sm[];
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5;
  for(int i=tid;i<16384;i+=blockDim.x) sm[i]=seed*(i%97)*1e-3;
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++){acc[i][j][0]=seed*i;acc[i][j][1]=seed*j;}
  double a[4],b[4]; double f[NCH];
#pragma unroll
  for(int c=0;c<NCH;c++) f[c]=seed*(lane+c);
  const double* pa=sm+lane, *pb=sm+8192+(warp%16)*128+lane;
  for(int it=0;it<iters;it++){
    int nf=0;
    if(MODE==1){
#pragma unroll
      for(int r=0;r<NF;r++) dfma(f[r%NCH],1.0000001,seed);
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
          if(MODE==2 && NF>0){
            // spread: after DMMA number cnt issue DFMAs so that nf == cnt*NF/64
#pragma unroll
            for(int r=0;r<(NF+63)/64;r++) if(nf < cnt*NF/64){ dfma(f[nf%NCH],1.0000001,seed); ++nf; }
          }
          if(MODE==5 && cnt>32){
#pragma unroll
            for(int r=0;r<(NF/2+31)/32;r++) if(nf < (cnt-32)*(NF/2)/32){ dfma(f[nf%NCH],1.0000001,seed); ++nf; }
          }
        }
    }
    if(MODE==4){
#pragma unroll
      for(int r=0;r<NF;r++) dfma(f[r%NCH],1.0000001,seed);
    }
    if(MODE==5){
#pragma unroll
      for(int r=0;r<NF/2;r++) dfma(f[r%NCH],1.0000001,seed);
    }
  }
  double s=0;
#pragma unroll
  for(int c=0;c<NCH;c++) s+=f[c];
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++) s+=acc[i][j][0]+acc[i][j][1];
  out[blockIdx.x*blockDim.x+tid]=s;
}
static float base_ms[33];
template<int MODE,int NF,int NCH,int WARPS> void run(double* out,int sms,const char* name){
  int iters=4000; size_t smem=16384*8;
  CK(cudaFuncSetAttribute(k<MODE,NF,NCH,WARPS>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE,NF,NCH,WARPS><<<sms,WARPS*32,smem>>>(out,iters,1.0000001); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<3;r++){ cudaEventRecord(e0); k<MODE,NF,NCH,WARPS><<<sms,WARPS*32,smem>>>(out,iters,1.0000001); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  double flops=(double)sms*WARPS*iters*64.0*512.0;
  if(MODE==0) base_ms[WARPS]=best;
  // pipe clocks per DFMA: extra time relative to the DMMA-only run, in units of one DMMA (16.2 clocks measured)
  double per_dmma_ms = base_ms[WARPS]/(64.0*iters);   // per warp-DMMA wall share
  double clk = (NF>0) ? (best-base_ms[WARPS])/(NF*(double)iters)/per_dmma_ms*16.2 : 0.0;
  printf("  \"%s\": {\"ms\": %.3f, \"dmma_tflops\": %.3f, \"clocks_per_dfma\": %.2f},\n",name,best,flops/best/1e9,clk);
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out,sizeof(double)*sms*1024));
  printf("{\n");
  run<0,0,1,16>(out,sms,"w16_none");
  run<1,64,1,16>(out,sms,"w16_front64_ch1"); run<1,64,2,16>(out,sms,"w16_front64_ch2"); run<1,64,4,16>(out,sms,"w16_front64_ch4");
  run<1,64,8,16>(out,sms,"w16_front64_ch8"); run<1,64,16,16>(out,sms,"w16_front64_ch16");
  run<2,64,1,16>(out,sms,"w16_spread64_ch1"); run<2,64,4,16>(out,sms,"w16_spread64_ch4"); run<2,64,8,16>(out,sms,"w16_spread64_ch8");
  run<2,64,16,16>(out,sms,"w16_spread64_ch16");
  run<4,64,8,16>(out,sms,"w16_back64_ch8"); run<5,64,8,16>(out,sms,"w16_half64_ch8"); run<5,64,1,16>(out,sms,"w16_half64_ch1");
  run<2,96,8,16>(out,sms,"w16_spread96_ch8"); run<5,96,8,16>(out,sms,"w16_half96_ch8"); run<1,96,8,16>(out,sms,"w16_front96_ch8");
  run<2,16,1,16>(out,sms,"w16_spread16_ch1"); run<2,16,4,16>(out,sms,"w16_spread16_ch4"); run<1,16,1,16>(out,sms,"w16_front16_ch1");
  run<0,0,1,32>(out,sms,"w32_none");
  run<1,64,1,32>(out,sms,"w32_front64_ch1"); run<1,64,8,32>(out,sms,"w32_front64_ch8"); run<2,64,8,32>(out,sms,"w32_spread64_ch8");
  run<2,64,1,32>(out,sms,"w32_spread64_ch1");
  run<0,0,1,8>(out,sms,"w8_none");
  run<1,64,8,8>(out,sms,"w8_front64_ch8"); run<2,64,8,8>(out,sms,"w8_spread64_ch8"); run<2,64,1,8>(out,sms,"w8_spread64_ch1");
  printf("  \"end\": 0\n}\n");
  return 0;
}
