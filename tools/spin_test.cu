// concurrency probe: kernel A spins on a flag that kernel B (other stream, launched later) sets
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin (volatile unsigned long long* f, int* err)
{
  unsigned long long t0, t1; asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*f < 1ull) { asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t1)); if (t1 - t0 > 3000000000ull) { *err = 1; break; } __nanosleep(200); }
}
__global__ void setf (unsigned long long* f) { __threadfence_system(); *((volatile unsigned long long*) f) = 1ull; }
__global__ void busy (double* x, long n) { for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) x[i] = x[i] * 1.0001 + 1.0; }
int main ()
{
  unsigned long long* f; int* err; double* x; long n = 1 << 26;
  cudaMalloc(&f, 8); cudaMalloc(&err, 4); cudaMalloc(&x, n * 8); cudaMemset(f, 0, 8); cudaMemset(err, 0, 4); cudaMemset(x, 0, n * 8);
  cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  // warm every kernel (lazy loading)
  setf<<<1, 1, 0, b>>>(f); spin<<<1, 1, 0, a>>>(f, err); busy<<<1024, 256, 0, b>>>(x, n); cudaDeviceSynchronize(); cudaMemset(f, 0, 8);
  for (int variant = 0; variant < 2; variant++)
    {
      cudaMemset(f, 0, 8); cudaMemset(err, 0, 4); cudaDeviceSynchronize();
      spin<<<1, 1, 0, a>>>(f, err);
      if (variant == 1) busy<<<100000, 128, 0, b>>>(x, n);
      setf<<<1, 1, 0, b>>>(f);
      cudaDeviceSynchronize();
      int e; cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost);
      printf("variant %d: %s (%s)\n", variant, e ? "TIMEOUT" : "ok", cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
