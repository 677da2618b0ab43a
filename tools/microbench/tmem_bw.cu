// tmem_bw.cu — how fast can an SM read its tensor memory?  (tcgen05.ld 32x32b, fp32 columns)
// The attention kernels read every score from TMEM exactly once (flash: 128 x 128 x 4 B per KV tile; cross-IP: twice);
// this microbenchmark measures the ceiling of that read path on the device it runs on.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/microbench/tmem_bw.cu -o tools/microbench/bin/tmem_bw
//   tools/microbench/bin/tmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[X]);
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// X columns per load, U loads in flight before one tcgen05.wait::ld
template <int X, int U>
__global__ void __launch_bounds__(512, 1) tmem_read(long long* cycles, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[U][X];
#pragma unroll
    for (int u = 0; u < U; ++u) ld<X>(base + ((i * U + u) * X) % (512 - X + 1) / X * X, r[u]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < X; ++j) acc ^= r[u][j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int X, int U>
static void run(int warps, long long* d_cycles, uint32_t* d_sink, int sms) {
  const int iters = 4096;
  tmem_read<X, U><<<sms, warps * 32>>>(d_cycles, d_sink, iters);
  tmem_read<X, U><<<sms, warps * 32>>>(d_cycles, d_sink, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("x%d u%d warps %d: %s\n", X, U, warps, cudaGetErrorString(e));
    return;
  }
  long long h[256];
  cudaMemcpy(h, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = static_cast<double>(warps) * iters * U * 32.0 * X * 4.0;
  printf("tcgen05.ld 32x32b.x%-2d  %d in flight  %2d warps/SM : %7.1f B/clk/SM  (%lld clk)\n", X, U, warps, bytes / mx, mx);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, sizeof(long long) * 256);
  cudaMalloc(&d_sink, 4);
  printf("%s, %d SMs: TMEM read bandwidth per SM (every warp reads its own 32-lane quadrant)\n", prop.name, sms);
  for (int warps : {4, 8, 16}) {
    run<16, 1>(warps, d_cycles, d_sink, sms);
    run<32, 1>(warps, d_cycles, d_sink, sms);
    run<16, 2>(warps, d_cycles, d_sink, sms);
    run<32, 2>(warps, d_cycles, d_sink, sms);
  }
  return 0;
}
