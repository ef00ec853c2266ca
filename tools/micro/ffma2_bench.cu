// Microbenchmark: issue throughput of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pack(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(unsigned long long v, float&a, float&b){ asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template<int MODE> __global__ void k(float* out, int iters, float s){
  float a[8]; unsigned long long p[8];
  for (int i=0;i<8;++i){ a[i] = threadIdx.x*0.001f + i; p[i] = pack(a[i], a[i]+1.f); }
  unsigned long long ps = pack(s, s);
  for (int it=0; it<iters; ++it){
    if (MODE==0){
      #pragma unroll
      for (int i=0;i<8;++i) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(s));
    } else if (MODE==1) {
      #pragma unroll
      for (int i=0;i<8;++i) p[i] = fma2(p[i], ps, p[i]);
    } else {  // mixed: 4 FFMA2 + 4 alu ops (integer adds) to see co-issue
      #pragma unroll
      for (int i=0;i<4;++i) p[i] = fma2(p[i], ps, p[i]);
      #pragma unroll
      for (int i=4;i<8;++i) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(s));
    }
  }
  float acc=0; for (int i=0;i<8;++i){ float u,v; unpack(p[i],u,v); acc += a[i]+u+v; }
  out[blockIdx.x*blockDim.x+threadIdx.x] = acc;
}
int main(){
  float* d; cudaMalloc(&d, 148*8*256*4*4);
  const int iters = 20000; const int blocks = 148*8, threads = 256;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode=0; mode<3; ++mode){
    for (int rep=0; rep<2; ++rep){
      cudaEventRecord(e0);
      if (mode==0) k<0><<<blocks,threads>>>(d, iters, 1.0001f); else if (mode==1) k<1><<<blocks,threads>>>(d, iters, 1.0001f); else k<2><<<blocks,threads>>>(d, iters, 1.0001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double winstr = (double)blocks*threads/32*iters*8;
      if (rep==1) printf("mode %d: %.3f ms, %.1f G warp-instr/s, %.2f TFLOP/s (fp32)\n", mode, ms, winstr/ms/1e6, winstr*32*2*(mode==0?1:(mode==1?2:1.5))/ms/1e9);
    }
  }
  return 0;
}
