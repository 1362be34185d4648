// Micro-benchmarks that decide the MFCC kernel design on B200 (sm_100a):
// scalar vs packed (f32x2) FP32 issue rate, shared-memory bandwidth, shuffle rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

constexpr int ITER = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float r[8];
  #pragma unroll
  for (int j=0;j<8;j++) r[j] = threadIdx.x + j;
  for (int i=0;i<ITER;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) r[j] = fmaf(r[j], a, b);
  }
  float s=0; for (int j=0;j<8;j++) s+=r[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_fadd(float* out, float a, float b) {
  float r[8];
  #pragma unroll
  for (int j=0;j<8;j++) r[j] = threadIdx.x + j;
  for (int i=0;i<ITER;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) r[j] = r[j] + r[(j+1)&7];
  }
  float s=0; for (int j=0;j<8;j++) s+=r[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  float2 r[8]; float2 A=make_float2(a,a*1.0001f), B=make_float2(b,b);
  #pragma unroll
  for (int j=0;j<8;j++) r[j] = make_float2(threadIdx.x + j, j);
  for (int i=0;i<ITER;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) r[j] = __ffma2_rn(r[j], A, B);
  }
  float s=0; for (int j=0;j<8;j++) s+=r[j].x+r[j].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_fadd2(float* out, float a, float b) {
  float2 r[8];
  #pragma unroll
  for (int j=0;j<8;j++) r[j] = make_float2(threadIdx.x + j, j+a);
  for (int i=0;i<ITER;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) r[j] = __fadd2_rn(r[j], r[(j+1)&7]);
  }
  float s=0; for (int j=0;j<8;j++) s+=r[j].x+r[j].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// mixed: 3 FADD2 per FFMA2 with distinct registers (butterfly-like)
__global__ void k_mix2(float* out, float a, float b) {
  float2 r[16]; float2 A=make_float2(a,a*1.0001f);
  #pragma unroll
  for (int j=0;j<16;j++) r[j] = make_float2(threadIdx.x + j, j+b);
  for (int i=0;i<ITER/2;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) { float2 u=r[j], v=r[j+8];
      r[j]=__fadd2_rn(u,v); float2 d=__fadd2_rn(u,make_float2(-v.x,-v.y)); r[j+8]=__ffma2_rn(d,A,r[j]); }
  }
  float s=0; for (int j=0;j<16;j++) s+=r[j].x+r[j].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_mix1(float* out, float a, float b) {
  float r[16];
  #pragma unroll
  for (int j=0;j<16;j++) r[j] = threadIdx.x + j + b;
  for (int i=0;i<ITER/2;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) { float u=r[j], v=r[j+8];
      r[j]=u+v; float d=u-v; r[j+8]=fmaf(d,a,r[j]); }
  }
  float s=0; for (int j=0;j<16;j++) s+=r[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int VEC> __global__ void k_lds(float* out) {
  extern __shared__ float4 sm[];
  int t = threadIdx.x;
  for (int i=t;i<4096;i+=blockDim.x) sm[i]=make_float4(i,i,i,i);
  __syncthreads();
  float acc=0;
  for (int i=0;i<ITER/4;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) {
      int idx = (t + j*blockDim.x + i) & 4095;
      if (VEC==4) { float4 v = sm[idx]; acc += v.x+v.w; }
      else if (VEC==2) { float2 v = ((float2*)sm)[idx]; acc += v.x+v.y; }
      else { float v = ((float*)sm)[idx]; acc += v; }
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
template<int VEC> __global__ void k_sts(float* out) {
  extern __shared__ float4 sm[];
  int t = threadIdx.x;
  for (int i=0;i<ITER/4;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) {
      int idx = (t + j*blockDim.x + i) & 4095;
      if (VEC==4) sm[idx] = make_float4(i,j,t,idx);
      else if (VEC==2) ((float2*)sm)[idx] = make_float2(i,j);
      else ((float*)sm)[idx] = i;
    }
  }
  __syncthreads();
  out[blockIdx.x*blockDim.x+threadIdx.x]=sm[t].x;
}
__global__ void k_shfl(float* out) {
  float r[8];
  #pragma unroll
  for (int j=0;j<8;j++) r[j]=threadIdx.x*j;
  int src = (16 - (threadIdx.x&15)) & 15;
  for (int i=0;i<ITER/4;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) r[j] = __shfl_sync(0xffffffffu, r[j], src, 16) + 1.0f;
  }
  float s=0; for (int j=0;j<8;j++) s+=r[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// FFMA2 stream + LDS.128 stream together: do they overlap?
__global__ void k_ffma2_lds(float* out, float a, float b) {
  extern __shared__ float4 sm[];
  int t = threadIdx.x;
  for (int i=t;i<4096;i+=blockDim.x) sm[i]=make_float4(i,i,i,i);
  __syncthreads();
  float2 r[8]; float2 A=make_float2(a,a*1.0001f), B=make_float2(b,b);
  #pragma unroll
  for (int j=0;j<8;j++) r[j] = make_float2(threadIdx.x + j, j);
  float acc=0;
  for (int i=0;i<ITER/4;i++) {
    float4 v = sm[(t + i) & 4095];
    #pragma unroll
    for (int k=0;k<4;k++)
    #pragma unroll
    for (int j=0;j<8;j++) r[j] = __ffma2_rn(r[j], A, B);
    acc += v.x + v.w;
  }
  float s=acc; for (int j=0;j<8;j++) s+=r[j].x+r[j].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_i2f(float* out, const short* in) {
  int t = threadIdx.x; float acc = 0; int v = in[t];
  for (int i=0;i<ITER;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) { acc += (float)(short)(v + j); v = v*3+i; }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
__global__ void k_log(float* out, float a) {
  float r[8];
  #pragma unroll
  for (int j=0;j<8;j++) r[j] = threadIdx.x + j + a;
  for (int i=0;i<ITER/8;i++) {
    #pragma unroll
    for (int j=0;j<8;j++) r[j] = logf(r[j]) + 100.f;
  }
  float s=0; for (int j=0;j<8;j++) s+=r[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<class F> float timeit(F f) {
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for(int i=0;i<5;i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms/5;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s sms=%d clock=%d kHz smem/SM=%zu\n", p.name, sms, clk, p.sharedMemPerMultiprocessor);
  float* out; CK(cudaMalloc(&out, sizeof(float)*sms*8*1024));
  short* in; CK(cudaMalloc(&in, 4096)); cudaMemset(in,1,4096);
  for (int thr : {256, 512, 1024}) {
    int blocks = sms * (2048/thr > 2 ? 2 : 2048/thr);
    if (thr==1024) blocks = sms*2; if (thr==512) blocks=sms*2; if (thr==256) blocks=sms*4;
    double tot = (double)blocks*thr;
    printf("--- threads/block=%d blocks=%d (resident warps/SM=%d)\n", thr, blocks, blocks/sms*thr/32);
    float ms;
    ms=timeit([&]{k_ffma<<<blocks,thr>>>(out,1.0001f,0.5f);});  printf("ffma   scalar: %.3f ms  %.1f FMA/clk/SM @1.9GHz-equiv  %.2f TFLOP/s\n", ms, tot*ITER*8/(ms*1e-3)/sms/1.9e9, 2*tot*ITER*8/(ms*1e-3)/1e12);
    ms=timeit([&]{k_ffma2<<<blocks,thr>>>(out,1.0001f,0.5f);}); printf("ffma2  packed: %.3f ms  %.1f FMA/clk/SM  %.2f TFLOP/s\n", ms, 2*tot*ITER*8/(ms*1e-3)/sms/1.9e9, 4*tot*ITER*8/(ms*1e-3)/1e12);
    ms=timeit([&]{k_fadd<<<blocks,thr>>>(out,1.0001f,0.5f);});  printf("fadd   scalar: %.3f ms  %.1f ADD/clk/SM\n", ms, tot*ITER*8/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_fadd2<<<blocks,thr>>>(out,1.0001f,0.5f);}); printf("fadd2  packed: %.3f ms  %.1f ADD/clk/SM\n", ms, 2*tot*ITER*8/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_mix1<<<blocks,thr>>>(out,1.0001f,0.5f);});  printf("mix    scalar: %.3f ms  %.1f op/clk/SM\n", ms, tot*(ITER/2)*8*3/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_mix2<<<blocks,thr>>>(out,1.0001f,0.5f);});  printf("mix2   packed: %.3f ms  %.1f op/clk/SM (scalar-equivalent)\n", ms, 2*tot*(ITER/2)*8*3/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_lds<4><<<blocks,thr,65536>>>(out);}); printf("lds.128: %.3f ms  %.1f B/clk/SM\n", ms, tot*(ITER/4)*8*16/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_lds<2><<<blocks,thr,65536>>>(out);}); printf("lds.64 : %.3f ms  %.1f B/clk/SM\n", ms, tot*(ITER/4)*8*8/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_lds<1><<<blocks,thr,65536>>>(out);}); printf("lds.32 : %.3f ms  %.1f B/clk/SM\n", ms, tot*(ITER/4)*8*4/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_sts<4><<<blocks,thr,65536>>>(out);}); printf("sts.128: %.3f ms  %.1f B/clk/SM\n", ms, tot*(ITER/4)*8*16/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_sts<2><<<blocks,thr,65536>>>(out);}); printf("sts.64 : %.3f ms  %.1f B/clk/SM\n", ms, tot*(ITER/4)*8*8/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_shfl<<<blocks,thr>>>(out);}); printf("shfl   : %.3f ms  %.1f lanes/clk/SM\n", ms, tot*(ITER/4)*8/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_ffma2_lds<<<blocks,thr,65536>>>(out,1.0001f,0.5f);}); printf("ffma2+lds128 (32 ffma2 per lds128): %.3f ms  %.1f FMA/clk/SM, %.1f B/clk/SM\n", ms, 2*tot*(ITER/4)*32/(ms*1e-3)/sms/1.9e9, tot*(ITER/4)*16/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_i2f<<<blocks,thr>>>(out,in);}); printf("i2f    : %.3f ms  %.1f cvt/clk/SM\n", ms, tot*ITER*8/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_log<<<blocks,thr>>>(out,2.0f);}); printf("logf   : %.3f ms  %.1f logf/clk/SM\n", ms, tot*(ITER/8)*8/(ms*1e-3)/sms/1.9e9);
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
