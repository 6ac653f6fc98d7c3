// micro-benchmark: H2D/D2H copy shapes used by nav24_orb_detect_batch (pinned host memory)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
int main(){
  const int F=256,H=376,W=1241,P=1280;
  size_t nb=(size_t)F*H*W;
  unsigned char *h,*d,*hp; CK(cudaHostAlloc(&h,nb,0)); CK(cudaMalloc(&d,(size_t)F*H*P)); memset(h,1,nb);
  hp=(unsigned char*)malloc(nb); memset(hp,1,nb);
  cudaStream_t s; CK(cudaStreamCreate(&s)); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for(int it=0;it<3;++it){
    cudaEventRecord(a,s); CK(cudaMemcpy2DAsync(d,P,h,W,W,(size_t)H*F,cudaMemcpyHostToDevice,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D 2D pinned  %.3f ms  %.1f GB/s\n",ms,nb/ms/1e6);
    cudaEventRecord(a,s); CK(cudaMemcpyAsync(d,h,nb,cudaMemcpyHostToDevice,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D 1D pinned  %.3f ms  %.1f GB/s\n",ms,nb/ms/1e6);
    cudaEventRecord(a,s); CK(cudaMemcpyAsync(d,hp,nb,cudaMemcpyHostToDevice,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D 1D pageable %.3f ms  %.1f GB/s\n",ms,nb/ms/1e6);
    size_t ob=(size_t)F*2064*60;
    cudaEventRecord(a,s); CK(cudaMemcpyAsync(h,d,ob,cudaMemcpyDeviceToHost,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("D2H 1D pinned %zu B %.3f ms  %.1f GB/s\n",ob,ms,ob/ms/1e6);
    // chunked 1D copies of 32 frames
    cudaEventRecord(a,s); for(int c=0;c<8;++c) CK(cudaMemcpyAsync(d+(size_t)c*32*H*W,h+(size_t)c*32*H*W,(size_t)32*H*W,cudaMemcpyHostToDevice,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D 8x1D pinned  %.3f ms  %.1f GB/s\n",ms,nb/ms/1e6);
  }
  // bidirectional
  cudaStream_t s2; cudaStreamCreate(&s2); unsigned char* h2; CK(cudaHostAlloc(&h2,nb,0));
  cudaEventRecord(a,s); CK(cudaMemcpyAsync(d,h,nb,cudaMemcpyHostToDevice,s)); CK(cudaMemcpyAsync(h2,d+nb/2,nb/4,cudaMemcpyDeviceToHost,s2)); cudaStreamSynchronize(s2); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
  printf("bidir H2D %zu + D2H %zu: %.3f ms\n",nb,nb/4,ms);
  return 0;
}
