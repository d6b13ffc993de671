// Micro-test: fragment layout of tcgen05.st.16x256b (written) as seen by tcgen05.ld.32x32b (thread = lane row).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t* out) {
  __shared__ uint32_t tb;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&tb)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tb;
  const int i = threadIdx.x;
  // x2: 8 regs, 16 columns; two 16-lane halves
  for (int h = 0; h < 2; ++h) {
    uint32_t r[8];
    for (int q = 0; q < 8; ++q) r[q] = (uint32_t)(h * 10000 + i * 100 + q);
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(base + ((uint32_t)(16 * h) << 16)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 16; ++c) out[i * 16 + c] = v[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 32 * 16 * 4);
  k<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  uint32_t h[512]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int r = 0; r < 32; ++r) { printf("row %2d:", r); for (int c = 0; c < 16; ++c) printf(" %5u", h[r * 16 + c]); printf("\n"); }
  return 0;
}
