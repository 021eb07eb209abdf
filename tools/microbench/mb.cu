// Microbenchmarks that decide the paint/sort design on B200 (run under gpurun).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cufft.h>
#include <cub/cub.cuh>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s){ s = s*1664525u+1013904223u; return s>>8; }

template<typename T, int MODE>  // MODE 0 atomic, 1 plain RMW (racy; throughput only)
__global__ void smem_atom(int iters, int S, T* out){
  extern __shared__ unsigned char raw[];
  T* a = (T*)raw;
  for(int i=threadIdx.x;i<S;i+=blockDim.x) a[i]=T(0);
  __syncthreads();
  uint32_t s = threadIdx.x*7919u + blockIdx.x*104729u + 1u;
  for(int i=0;i<iters;i++){
    uint32_t r = lcg(s) % S;
    if(MODE==0) atomicAdd(&a[r], T(1));
    else a[r] = a[r] + T(1);
  }
  __syncthreads();
  T acc=T(0);
  for(int i=threadIdx.x;i<S;i+=blockDim.x) acc+=a[i];
  if(acc==T(123456789)) out[0]=acc;
}

// sorted-ish addresses: lane l hits base+l (conflict free), base random
template<typename T>
__global__ void smem_atom_cf(int iters, int S, T* out){
  extern __shared__ unsigned char raw[];
  T* a = (T*)raw;
  for(int i=threadIdx.x;i<S;i+=blockDim.x) a[i]=T(0);
  __syncthreads();
  uint32_t s = (threadIdx.x/32)*7919u + blockIdx.x*104729u + 1u;
  int lane = threadIdx.x&31;
  for(int i=0;i<iters;i++){
    uint32_t r = (lcg(s) % (S-32)) + lane;
    atomicAdd(&a[r], T(1));
  }
  __syncthreads();
  T acc=T(0);
  for(int i=threadIdx.x;i<S;i+=blockDim.x) acc+=a[i];
  if(acc==T(123456789)) out[0]=acc;
}

__global__ void gatom(int n, uint32_t mask, int* ctr, int* rank){
  int i = blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=n) return;
  uint32_t h = (uint32_t)i*2654435761u; h ^= h>>15; h*=2246822519u; h^=h>>13;
  rank[i] = atomicAdd(&ctr[h & mask], 1);
}
__global__ void gred(int n, uint32_t mask, int* ctr){
  int i = blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=n) return;
  uint32_t h = (uint32_t)i*2654435761u; h ^= h>>15; h*=2246822519u; h^=h>>13;
  atomicAdd(&ctr[h & mask], 1);
}

__global__ void conv_f2ll(int iters, float* out){
  float x = threadIdx.x*1.0001f+1.f; long long acc=0;
  for(int i=0;i<iters;i++){ acc += __float2ll_rn(x); x = x*1.0000001f+0.5f; }
  if(acc==12345) out[0]=x;
}
__global__ void conv_f2i(int iters, float* out){
  float x = threadIdx.x*1.0001f+1.f; long long acc=0;
  for(int i=0;i<iters;i++){ acc += (long long)__float2int_rn(x); x = x*1.0000001f+0.5f; }
  if(acc==12345) out[0]=x;
}
__global__ void conv_ll2f(int iters, float* out){
  long long v = threadIdx.x*977+13; float acc=0;
  for(int i=0;i<iters;i++){ acc += __ll2float_rn(v); v = v*3+1; }
  if(acc==12345.f) out[0]=acc;
}
__global__ void shfl_k(int iters, float* out){
  float x = threadIdx.x; 
  for(int i=0;i<iters;i++){ x += __shfl_up_sync(0xffffffffu, x, 1); }
  if(x==12345.f) out[0]=x;
}
__global__ void match_k(int iters, float* out){
  uint32_t s = threadIdx.x*7919u+1u; unsigned acc=0;
  for(int i=0;i<iters;i++){ acc += __match_any_sync(0xffffffffu, lcg(s)&255u); }
  if(acc==12345u) out[0]=acc;
}
__global__ void copyk(const float4* __restrict__ a, float4* __restrict__ b, size_t n){
  size_t i = blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st = (size_t)gridDim.x*blockDim.x;
  for(;i<n;i+=st) b[i]=a[i];
}

template<class F> float timeit(F f, int rep=5){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<rep;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("dev %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm = p.multiProcessorCount;
  float* dout; CK(cudaMalloc(&dout, 64));
  // ---- smem atomics
  {
    int iters=2000, S=4096, thr=256, blocks=nsm*4;
    double ops = (double)iters*thr*blocks;
    #define RUN(T,MODE,name) { float ms=timeit([&]{ smem_atom<T,MODE><<<blocks,thr,S*sizeof(T)>>>(iters,S,(T*)dout); }); \
       printf("smem %-18s random S=%d: %.3f ms  %.2f Gops/s  %.3f ops/clk/SM(@1.9GHz)\n", name, S, ms, ops/ms*1e-6, ops/ms*1e-6/nsm/1.9); }
    RUN(int,0,"atomicAdd int32"); RUN(unsigned long long,0,"atomicAdd u64"); RUN(float,0,"atomicAdd f32"); RUN(double,0,"atomicAdd f64");
    RUN(int,1,"plain RMW int32"); RUN(float,1,"plain RMW f32"); RUN(unsigned long long,1,"plain RMW u64");
    #define RUNCF(T,name) { float ms=timeit([&]{ smem_atom_cf<T><<<blocks,thr,S*sizeof(T)>>>(iters,S,(T*)dout); }); \
       printf("smem %-18s confl-free S=%d: %.3f ms  %.2f Gops/s  %.3f ops/clk/SM\n", name, S, ms, ops/ms*1e-6, ops/ms*1e-6/nsm/1.9); }
    RUNCF(int,"atomicAdd int32"); RUNCF(unsigned long long,"atomicAdd u64"); RUNCF(float,"atomicAdd f32");
  }
  // ---- conversions / shuffles
  {
    int iters=4000, thr=256, blocks=nsm*8; double ops=(double)iters*thr*blocks;
    float ms;
    ms=timeit([&]{conv_f2ll<<<blocks,thr>>>(iters,dout);}); printf("F2I.S64.F32: %.3f ms %.2f ops/clk/SM\n", ms, ops/ms*1e-6/nsm/1.9);
    ms=timeit([&]{conv_f2i<<<blocks,thr>>>(iters,dout);});  printf("F2I.S32.F32(+add64): %.3f ms %.2f ops/clk/SM\n", ms, ops/ms*1e-6/nsm/1.9);
    ms=timeit([&]{conv_ll2f<<<blocks,thr>>>(iters,dout);}); printf("I2F.F32.S64: %.3f ms %.2f ops/clk/SM\n", ms, ops/ms*1e-6/nsm/1.9);
    ms=timeit([&]{shfl_k<<<blocks,thr>>>(iters,dout);});    printf("SHFL.UP(dep chain): %.3f ms %.2f lane-ops/clk/SM\n", ms, ops/ms*1e-6/nsm/1.9);
    ms=timeit([&]{match_k<<<blocks,thr>>>(iters,dout);});   printf("MATCH.ANY: %.3f ms %.2f lane-ops/clk/SM\n", ms, ops/ms*1e-6/nsm/1.9);
  }
  // ---- global atomics
  {
    int n=10000000; int *ctr,*rank; CK(cudaMalloc(&ctr,(size_t)(1<<27)*4)); CK(cudaMalloc(&rank,(size_t)n*4));
    for(int lg : {12,15,18,20,24,27}){
      uint32_t mask=(1u<<lg)-1;
      CK(cudaMemset(ctr,0,(size_t)(1<<lg)*4));
      float ms=timeit([&]{gatom<<<(n+255)/256,256>>>(n,mask,ctr,rank);});
      float ms2=timeit([&]{gred<<<(n+255)/256,256>>>(n,mask,ctr);});
      printf("global atomicAdd 10M ops on 2^%d counters: ret+store %.3f ms, red %.3f ms\n", lg, ms, ms2);
    }
    cudaFree(ctr); cudaFree(rank);
  }
  // ---- copy bandwidth
  {
    size_t n = (size_t)1<<30; float4 *a,*b; CK(cudaMalloc(&a,n)); CK(cudaMalloc(&b,n)); CK(cudaMemset(a,1,n));
    float ms=timeit([&]{copyk<<<nsm*16,512>>>(a,b,n/16);});
    printf("copy 1GiB: %.3f ms  %.1f GB/s (r+w)\n", ms, 2.0*n/ms*1e-6);
    float ms2=timeit([&]{cudaMemsetAsync(b,0,n);});
    printf("memset 1GiB: %.3f ms  %.1f GB/s\n", ms2, 1.0*n/ms2*1e-6);
    cudaFree(a); cudaFree(b);
  }
  // ---- CUB radix sort pairs
  {
    int n=10000000; uint32_t *k0,*k1,*v0,*v1; CK(cudaMalloc(&k0,n*4));CK(cudaMalloc(&k1,n*4));CK(cudaMalloc(&v0,n*4));CK(cudaMalloc(&v1,n*4));
    std::vector<uint32_t> h(n); uint32_t s=1; for(int i=0;i<n;i++){ s=s*1664525u+1013904223u; h[i]=s>>5; }
    CK(cudaMemcpy(k0,h.data(),n*4,cudaMemcpyHostToDevice));
    void* tmp=nullptr; size_t tb=0; cub::DeviceRadixSort::SortPairs(tmp,tb,k0,k1,v0,v1,n,0,27);
    CK(cudaMalloc(&tmp,tb));
    for(int bits : {12,15,18,20,24,27}){
      float ms=timeit([&]{cub::DeviceRadixSort::SortPairs(tmp,tb,k0,k1,v0,v1,n,0,bits);});
      printf("cub SortPairs 10M u32/u32 bits=%d: %.3f ms\n", bits, ms);
    }
    float ms=timeit([&]{cub::DeviceRadixSort::SortKeys(tmp,tb,k0,k1,n,0,27);});
    printf("cub SortKeys 10M u32 bits=27: %.3f ms\n", ms);
    // scan
    size_t tb2=0; cub::DeviceScan::ExclusiveSum(nullptr,tb2,k0,k1,1<<24);
    void* tmp2; CK(cudaMalloc(&tmp2,tb2));
    CK(cudaFree(k0)); CK(cudaMalloc(&k0,(size_t)(1<<24)*4)); CK(cudaFree(k1)); CK(cudaMalloc(&k1,(size_t)(1<<24)*4));
    ms=timeit([&]{cub::DeviceScan::ExclusiveSum(tmp2,tb2,k0,k1,1<<24);});
    printf("cub ExclusiveSum 2^24 u32: %.3f ms\n", ms);
  }
  // ---- cuFFT
  {
    for(int N : {128,256,512}){
      size_t M=(size_t)N*N*N, Mc=(size_t)N*N*(N/2+1);
      int nb_f = (N==512)?2:4, nb_i=(N==512)?4:12;
      float* r; cufftComplex* c; CK(cudaMalloc(&r,M*4*nb_i)); CK(cudaMalloc(&c,Mc*8*nb_i)); CK(cudaMemset(r,0,M*4*nb_i)); CK(cudaMemset(c,0,Mc*8*nb_i));
      int dims[3]={N,N,N};
      cufftHandle pf,pi,pf1,pi1; size_t ws;
      cufftPlanMany(&pf,3,dims,nullptr,1,0,nullptr,1,0,CUFFT_R2C,nb_f);
      cufftPlanMany(&pi,3,dims,nullptr,1,0,nullptr,1,0,CUFFT_C2R,nb_i);
      cufftPlan3d(&pf1,N,N,N,CUFFT_R2C); cufftPlan3d(&pi1,N,N,N,CUFFT_C2R);
      cufftGetSize(pi,&ws);
      float ms;
      ms=timeit([&]{cufftExecR2C(pf,r,c);}); printf("cufft R2C %d^3 batch %d: %.3f ms (%.3f per xform; alg bytes %.0f MB -> %.0f GB/s)\n",N,nb_f,ms,ms/nb_f,(M*4+Mc*8)*1e-6,(M*4+Mc*8)*nb_f/ms*1e-6);
      ms=timeit([&]{cufftExecC2R(pi,c,r);}); printf("cufft C2R %d^3 batch %d: %.3f ms (%.3f per xform, ws %.0f MB) -> %.0f GB/s\n",N,nb_i,ms,ms/nb_i,ws*1e-6,(M*4+Mc*8)*nb_i/ms*1e-6);
      ms=timeit([&]{cufftExecR2C(pf1,r,c);}); printf("cufft R2C %d^3 single: %.3f ms\n",N,ms);
      ms=timeit([&]{cufftExecC2R(pi1,c,r);}); printf("cufft C2R %d^3 single: %.3f ms\n",N,ms);
      // in-place padded
      ms=timeit([&]{cufftExecR2C(pf1,(float*)c,c);}); printf("cufft R2C %d^3 single inplace: %.3f ms\n",N,ms);
      cufftDestroy(pf);cufftDestroy(pi);cufftDestroy(pf1);cufftDestroy(pi1);
      if(N==256){
        double* rd; cufftDoubleComplex* cd; CK(cudaMalloc(&rd,M*8*4)); CK(cudaMalloc(&cd,Mc*16*4)); CK(cudaMemset(rd,0,M*8*4)); CK(cudaMemset(cd,0,Mc*16*4));
        cufftHandle pd,pdi; cufftPlanMany(&pd,3,dims,nullptr,1,0,nullptr,1,0,CUFFT_D2Z,4); cufftPlanMany(&pdi,3,dims,nullptr,1,0,nullptr,1,0,CUFFT_Z2D,4);
        ms=timeit([&]{cufftExecD2Z(pd,rd,cd);}); printf("cufft D2Z 256^3 batch 4: %.3f ms\n",ms);
        ms=timeit([&]{cufftExecZ2D(pdi,cd,rd);}); printf("cufft Z2D 256^3 batch 4: %.3f ms\n",ms);
        cufftDestroy(pd);cufftDestroy(pdi); cudaFree(rd); cudaFree(cd);
      }
      cudaFree(r); cudaFree(c);
    }
  }
  return 0;
}
