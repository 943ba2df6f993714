// L2 -> shared-memory streaming bandwidth with cp.async.bulk (what the fused MLP's weight ring needs).
// Every CTA streams the same 12 MB arena (13 "layers") `reps` times through a ring of STAGES tiles.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

template <int TILE, int STAGES>
__global__ void __launch_bounds__(128, 1) stream_kernel(const char* arena, size_t arena_bytes, int reps, int skew, unsigned long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = (uint64_t*)(smem + (size_t)TILE * STAGES);
  if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const size_t tiles = arena_bytes / TILE;
  const size_t total = tiles * reps;
  const size_t start = skew ? (size_t)blockIdx.x * 37 % tiles : 0;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    size_t issued = 0, done = 0;
    uint32_t phase[STAGES];
    for (int i = 0; i < STAGES; ++i) phase[i] = 0;
    for (; issued < STAGES && issued < total; ++issued) {
      int st = issued % STAGES;
      mbar_expect(&bars[st], TILE);
      tma1d(smem + (size_t)st * TILE, arena + ((start + issued) % tiles) * TILE, TILE, &bars[st]);
    }
    for (; done < total; ++done) {
      int st = done % STAGES;
      mbar_wait(&bars[st], phase[st]); phase[st] ^= 1;
      if (issued < total) {
        mbar_expect(&bars[st], TILE);
        tma1d(smem + (size_t)st * TILE, arena + ((start + issued) % tiles) * TILE, TILE, &bars[st]);
        ++issued;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int TILE, int STAGES>
void run(const char* arena, size_t bytes, int grid, int skew) {
  unsigned long long* d_c; cudaMalloc(&d_c, grid * 8);
  size_t smem = (size_t)TILE * STAGES + 256;
  cudaFuncSetAttribute(stream_kernel<TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int reps = 20;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  stream_kernel<TILE, STAGES><<<grid, 128, smem>>>(arena, bytes, 2, skew, d_c);
  cudaEventRecord(a);
  stream_kernel<TILE, STAGES><<<grid, 128, smem>>>(arena, bytes, reps, skew, d_c);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double total = (double)bytes * reps * grid;
  printf("tile %6d stages %2d grid %3d skew %d: %.2f TB/s aggregate, %.1f GB/s per SM (%s)\n", TILE, STAGES, grid, skew,
         total / (ms * 1e-3) / 1e12, total / grid / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_c);
}

int main() {
  size_t bytes = 12u << 20;
  char* arena; cudaMalloc(&arena, bytes); cudaMemset(arena, 1, bytes);
  run<16384, 4>(arena, bytes, 148, 0);
  run<16384, 8>(arena, bytes, 148, 0);
  run<16384, 8>(arena, bytes, 148, 1);
  run<32768, 4>(arena, bytes, 148, 0);
  run<8192, 8>(arena, bytes, 148, 0);
  run<16384, 8>(arena, bytes, 74, 0);
  run<16384, 2>(arena, bytes, 148, 0);
  return 0;
}
