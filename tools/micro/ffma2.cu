#include <cuda_runtime.h>
#include <cstdio>
__global__ void k1(float* out, int iters, float a, float b) {
  float x0=threadIdx.x, x1=x0+1, x2=x0+2, x3=x0+3, x4=x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
  for (int i=0;i<iters;++i){ x0=fmaf(x0,a,b);x1=fmaf(x1,a,b);x2=fmaf(x2,a,b);x3=fmaf(x3,a,b);x4=fmaf(x4,a,b);x5=fmaf(x5,a,b);x6=fmaf(x6,a,b);x7=fmaf(x7,a,b);}
  float s=x0+x1+x2+x3+x4+x5+x6+x7; if (s==123456.f) out[0]=s;
}
__global__ void k2(float* out, int iters, float a, float b) {
  float2 A={a,a}, B={b,b};
  float2 x0={threadIdx.x,1.f}, x1={2.f,3.f}, x2={4.f,5.f}, x3={6.f,7.f}, x4={8.f,1.f}, x5={2.f,9.f}, x6={4.f,1.f}, x7={6.f,2.f};
  for (int i=0;i<iters;++i){ x0=__ffma2_rn(x0,A,B);x1=__ffma2_rn(x1,A,B);x2=__ffma2_rn(x2,A,B);x3=__ffma2_rn(x3,A,B);x4=__ffma2_rn(x4,A,B);x5=__ffma2_rn(x5,A,B);x6=__ffma2_rn(x6,A,B);x7=__ffma2_rn(x7,A,B);}
  float s=x0.x+x1.x+x2.x+x3.x+x4.y+x5.y+x6.y+x7.y+x0.y+x1.y+x2.y+x3.y+x4.x+x5.x+x6.x+x7.x; if (s==123456.f) out[0]=s;
}
int main(){ float* d; cudaMalloc(&d,64); int it=1<<14; int blocks=148*8, th=256;
 cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
 for(int r=0;r<2;++r){ cudaEventRecord(e0); k1<<<blocks,th>>>(d,it,0.9999f,1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);}
 printf("FFMA  : %.2f TFLOP/s, %.2f T warp-instr/s\n", 2.0*8*it*(double)blocks*th/ms/1e9, 8.0*it*(double)blocks*th/32/ms/1e9);
 for(int r=0;r<2;++r){ cudaEventRecord(e0); k2<<<blocks,th>>>(d,it,0.9999f,1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);}
 printf("FFMA2 : %.2f TFLOP/s, %.2f T warp-instr/s\n", 4.0*8*it*(double)blocks*th/ms/1e9, 8.0*it*(double)blocks*th/32/ms/1e9);
 return 0; }
