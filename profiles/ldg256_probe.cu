// standalone probe: 256-bit read-only loads with lane-per-row addressing
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float* __restrict__ p, float* o, int rows){
  int r = (blockIdx.x*blockDim.x + threadIdx.x) % rows;
  float v[8];
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]),"=f"(v[1]),"=f"(v[2]),"=f"(v[3]),"=f"(v[4]),"=f"(v[5]),"=f"(v[6]),"=f"(v[7]) : "l"(p + (size_t)r*128 + 32));
  float s=0; for(int i=0;i<8;i++) s+=v[i];
  o[blockIdx.x*blockDim.x + threadIdx.x]=s;
}
int main(){
  int rows=8193; float *p,*o; cudaMalloc(&p, rows*512); cudaMalloc(&o, 1<<20);
  cudaMemset(p,0,rows*512);
  k<<<256,256>>>(p,o,rows);
  cudaError_t e=cudaDeviceSynchronize(); printf("sync: %s\n", cudaGetErrorString(e));
  return 0;
}
