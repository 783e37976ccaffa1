#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b){
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b){
  unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template<int MODE>
__global__ void k(float* out, int iters, float s){
  // 8 independent chains
  if (MODE==0){
    float a[16]; 
    #pragma unroll
    for(int i=0;i<16;i++) a[i]=threadIdx.x*0.001f+i;
    float m = s, c = s*0.5f;
    for(int it=0; it<iters; it++){
      #pragma unroll
      for(int i=0;i<16;i++) a[i]=fmaf(a[i],m,c);
    }
    float r=0; 
    #pragma unroll
    for(int i=0;i<16;i++) r+=a[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=r;
  } else if (MODE==1) {
    unsigned long long a[8];
    #pragma unroll
    for(int i=0;i<8;i++){ float2 f=make_float2(threadIdx.x*0.001f+i, i); a[i]=*(unsigned long long*)&f; }
    float2 mf=make_float2(s,s), cf=make_float2(s*0.5f,s*0.25f);
    unsigned long long m=*(unsigned long long*)&mf, c=*(unsigned long long*)&cf;
    for(int it=0; it<iters; it++){
      #pragma unroll
      for(int i=0;i<8;i++) a[i]=fma2(a[i],m,c);
    }
    float r=0;
    #pragma unroll
    for(int i=0;i<8;i++){ float2 f=*(float2*)&a[i]; r+=f.x+f.y; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=r;
  } else if (MODE==2) { // scalar FADD
    float a[16];
    #pragma unroll
    for(int i=0;i<16;i++) a[i]=threadIdx.x*0.001f+i;
    float c = s;
    for(int it=0; it<iters; it++){
      #pragma unroll
      for(int i=0;i<16;i++) a[i]=a[i]+c;
    }
    float r=0;
    #pragma unroll
    for(int i=0;i<16;i++) r+=a[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=r;
  } else { // packed add
    unsigned long long a[8];
    #pragma unroll
    for(int i=0;i<8;i++){ float2 f=make_float2(threadIdx.x*0.001f+i, i); a[i]=*(unsigned long long*)&f; }
    float2 cf=make_float2(s*0.5f,s*0.25f);
    unsigned long long c=*(unsigned long long*)&cf;
    for(int it=0; it<iters; it++){
      #pragma unroll
      for(int i=0;i<8;i++) a[i]=add2(a[i],c);
    }
    float r=0;
    #pragma unroll
    for(int i=0;i<8;i++){ float2 f=*(float2*)&a[i]; r+=f.x+f.y; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=r;
  }
}
template<int MODE> void run(const char* name){
  float* out; cudaMalloc(&out, 148*8*256*4);
  int iters=20000;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148*8,256>>>(out,100,1.0001f);
  cudaEventRecord(e0);
  k<MODE><<<148*8,256>>>(out,iters,1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double lanes = 148.0*8*256*16.0*iters; // fp32 lane-ops
  printf("%s: %.3f ms, %.2f T lane-ops/s\n", name, ms, lanes/ms/1e9);
}
int main(){ run<0>("FFMA scalar"); run<1>("FFMA2 packed"); run<2>("FADD scalar"); run<3>("FADD2 packed"); return 0; }
