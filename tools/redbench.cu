#include <cstdio>
#include <cuda_runtime.h>
// each lane does `iters` REDs; mode 0: lanes hit consecutive doubles (coalesced), base varies per warp per iter
// mode 1: lanes hit random doubles within an N*N matrix; mode 2: all lanes of a warp same address (then warp-reduced? no: raw)
__global__ void k(double* g, long n, int iters, int mode, unsigned seed){
  unsigned s = seed ^ (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u;
  int lane = threadIdx.x&31; unsigned ws = seed ^ ((blockIdx.x*blockDim.x+threadIdx.x)>>5)*2246822519u;
  for(int i=0;i<iters;++i){
    s = s*1664525u+1013904223u; ws = ws*1664525u+1013904223u;
    long idx;
    if(mode==0) idx = ((ws>>4) % (n>32?n-32:1)) + (n>32?lane:0);
    else if(mode==1) idx = (s>>4) % n;
    else idx = (ws>>4)%n;
    atomicAdd(g+idx, 1.0);
  }
}
int main(){
  long n=400*400; double* g; cudaMalloc(&g,n*8); cudaMemset(g,0,n*8);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  long ns[5]={400*400, 400*8, 400, 32, 4};
  for(int t=0;t<5;++t){ n=ns[t]; printf("hot set = %ld doubles\n", n);
  for(int mode=1;mode<3;++mode){
    int blocks=148*16, threads=256, iters=200;
    k<<<blocks,threads>>>(g,n,10,mode,1); cudaDeviceSynchronize();
    cudaEventRecord(a); k<<<blocks,threads>>>(g,n,iters,mode,7); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms,a,b);
    double ops=(double)blocks*threads*iters;
    printf("mode %d: %.3f ms, %.3e lane-REDs/s (%.2f per clk per SM @1.9GHz)\n",mode,ms,ops/(ms*1e-3),ops/(ms*1e-3)/148/1.9e9);
  }}
  return 0;
}
