// Microbenchmark (B200): throughput of random global atomics / loads versus footprint, to size the hash-job design.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench atomics_bench.cu && ./atomics_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t fmix64(uint64_t k){k^=k>>33;k*=0xff51afd7ed558ccdull;k^=k>>33;k*=0xc4ceb9fe1a85ec53ull;k^=k>>33;return k;}
template<int MODE> __global__ void k(unsigned long long* tab, uint32_t* bits, uint64_t mask, int64_t n, unsigned long long* sink){
  unsigned long long acc=0;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=(int64_t)gridDim.x*blockDim.x){
    uint64_t h=fmix64((uint64_t)i*0x9e3779b97f4a7c15ull+12345);
    uint64_t slot=h&mask;
    if(MODE==0){ acc+=atomicOr(&bits[slot>>5],1u<<(slot&31)); }                 // atomicOr u32 (returning)
    else if(MODE==1){ acc+=atomicCAS(&tab[slot],0xFFFFFFFFFFFFFFFFull,h|1); }    // CAS u64
    else if(MODE==2){ acc+=tab[slot]; }                                          // random 8B load
    else if(MODE==3){ atomicOr(&bits[slot>>5],1u<<(slot&31)); }                  // RED (non-returning)
    else if(MODE==4){ acc+=atomicAdd(&bits[slot],1u); }                          // atomicAdd u32 returning
  }
  if(acc==0x1234567) *sink=acc;
}
int main(){
  const int64_t n=200000000; unsigned long long* tab; uint32_t* bits; unsigned long long* sink;
  size_t maxb=(size_t)4<<30; cudaMalloc(&tab,maxb); cudaMalloc(&sink,8); bits=(uint32_t*)tab;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  const char* names[]={"atomicOr.u32 (ret) bitmap","atomicCAS.u64","load.u64","red.or.u32 bitmap","atomicAdd.u32 (ret) array"};
  for(int mode=0;mode<5;mode++){
    for(int lg=20;lg<=32;lg+=2){ // footprint bytes = 2^lg
      uint64_t slots = mode==0||mode==3 ? ((uint64_t)1<<lg)*8 : mode==4 ? ((uint64_t)1<<lg)/4 : ((uint64_t)1<<lg)/8;
      cudaMemset(tab,0xFF,(size_t)1<<lg);
      float best=1e9;
      for(int rep=0;rep<3;rep++){
        cudaEventRecord(a);
        switch(mode){case 0:k<0><<<148*8,256>>>(tab,bits,slots-1,n,sink);break;case 1:k<1><<<148*8,256>>>(tab,bits,slots-1,n,sink);break;
          case 2:k<2><<<148*8,256>>>(tab,bits,slots-1,n,sink);break;case 3:k<3><<<148*8,256>>>(tab,bits,slots-1,n,sink);break;case 4:k<4><<<148*8,256>>>(tab,bits,slots-1,n,sink);break;}
        cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;
      }
      printf("%-28s footprint 2^%d B: %.3f ms  %.1f Gops/s\n",names[mode],lg,best,n/best/1e6);
    }
  }
  return 0;
}
