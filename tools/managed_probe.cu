// Diagnostic: does GPU access to cudaMallocManaged memory work on this box?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void touch(int* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }
int main() {
  int* m = nullptr; int* d = nullptr;
  printf("cudaMallocManaged: %s\n", cudaGetErrorString(cudaMallocManaged(&m, 2048)));
  touch<<<8, 64>>>(m, 512);
  printf("kernel on managed: launch %s, sync %s\n", cudaGetErrorString(cudaGetLastError()), cudaGetErrorString(cudaDeviceSynchronize()));
  printf("host read m[5] = %d\n", m ? m[5] : -1);
  cudaGetLastError();
  printf("cudaMalloc: %s\n", cudaGetErrorString(cudaMalloc(&d, 2048)));
  touch<<<8, 64>>>(d, 512);
  printf("kernel on device mem: sync %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
