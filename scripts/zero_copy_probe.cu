// probe: SM-driven writes of the output span of struct part into pinned host memory vs a DMA copy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(char *dst, const float *src, long n) {
  long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  char *b = dst + 128 * p;
  float v = src[p];
  // a_hydro 52..64, h 64, rho 76, entropy_dt 80, force union 84..112, bytes 113..117
  for (int o = 52; o < 112; o += 4) *(float *)(b + o) = v + o;
  b[113] = 1; b[114] = 2; b[116] = 3;
}
__global__ void k_write16(char *dst, const float *src, long n) {  // 16-byte stores where aligned
  long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  char *b = dst + 128 * p;
  float v = src[p];
  *(float *)(b + 52) = v; *(float *)(b + 56) = v; *(float *)(b + 60) = v;
  *(float4 *)(b + 64) = make_float4(v, v, v, v);
  *(float4 *)(b + 80) = make_float4(v, v, v, v);
  *(float4 *)(b + 96) = make_float4(v, v, v, v);
  *(int *)(b + 112) = 7; *(int *)(b + 116) = 3;
}
int main() {
  const long n = 2097152;
  char *h; float *s; char *d;
  cudaHostAlloc(&h, n * 128, cudaHostAllocDefault);
  cudaMalloc(&s, n * 4); cudaMalloc(&d, n * 128);
  cudaMemset(s, 0, n * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); cudaMemcpyAsync(h, d, n * 128, cudaMemcpyDeviceToHost); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("DMA full 268MB: %.3f ms\n", ms);
    cudaEventRecord(e0); k_write<<<(n + 255) / 256, 256>>>(h, s, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("zero-copy scalar stores: %.3f ms (%s)\n", ms, cudaGetErrorString(cudaGetLastError()));
    cudaEventRecord(e0); k_write16<<<(n + 255) / 256, 256>>>(h, s, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("zero-copy vector stores: %.3f ms (%s)\n", ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
