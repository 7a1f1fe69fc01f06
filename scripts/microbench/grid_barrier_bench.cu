// Grid-barrier variants on a cooperative 148 x 320 launch (the persistent kernels' shape): ns per barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/grid_barrier_bench scripts/microbench/grid_barrier_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

template <int VAR>
__device__ __forceinline__ void barrier(unsigned* counter, unsigned& target, int G) {
  if (VAR == 0) {            // product version
    fence_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      target += (unsigned)G;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
    }
    __syncthreads();
    fence_async();
  } else if (VAR == 1) {     // no proxy fences
    __syncthreads();
    if (threadIdx.x == 0) {
      target += (unsigned)G;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
    }
    __syncthreads();
  } else if (VAR == 2) {     // relaxed polling, one acquire fence at the end
    __syncthreads();
    if (threadIdx.x == 0) {
      target += (unsigned)G;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
  } else if (VAR == 3) {     // atom (returns the count): the last arriver knows without polling
    __syncthreads();
    if (threadIdx.x == 0) {
      target += (unsigned)G;
      unsigned old;
      asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
      if (old + 1 != target) {
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
      } else {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
    }
    __syncthreads();
  } else if (VAR == 4) {     // cooperative groups
    cg::this_grid().sync();
  } else if (VAR == 5) {     // a warp polls (32 requests in flight, first to see it wins)
    __syncthreads();
    if (threadIdx.x < 32) {
      target += (unsigned)G;
      if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (!__any_sync(0xffffffffu, (int)(v - target) >= 0));
    }
    __syncthreads();
  } else if (VAR == 6) {     // like 0 but every CTA leaves 64 floats of fresh stores behind (the slot writes of a GEMM phase)
    fence_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      target += (unsigned)G;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
    }
    __syncthreads();
    fence_async();
  }
}

template <int VAR>
__global__ void __launch_bounds__(320, 1) k(unsigned* counter, float* sink, int iters, long long* out) {
  extern __shared__ unsigned char smem[];
  unsigned target = 0;
  const int G = gridDim.x;
  if (VAR == 5) { /* target must advance in the polling warp only */ }
  barrier<VAR>(counter, target, G);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (VAR == 6) {
#pragma unroll
      for (int u = 0; u < 64; ++u) __stcg(sink + ((long)blockIdx.x * 64 + u) * 320 + threadIdx.x + (long)(i & 1) * 148 * 64 * 320, (float)i);
    }
    barrier<VAR>(counter, target, G);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int VAR>
static void run(const char* name, unsigned* counter, float* sink, long long* out, int G, int khz) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaMemset(counter, 0, 64);
  void* args[4] = {&counter, &sink, (void*)&iters, &out};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(counter, 0, 64);
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchCooperativeKernel((void*)k<VAR>, dim3(G), dim3(320), args, 200 * 1024, 0);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (err != cudaSuccess || e2 != cudaSuccess) { printf("%s: error %s %s\n", name, cudaGetErrorString(err), cudaGetErrorString(e2)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 1) printf("%-52s %7.0f ns per barrier (events)\n", name, ms * 1e6 / iters);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int G = p.multiProcessorCount;
  unsigned* counter; float* sink; long long* out;
  cudaMalloc(&counter, 64); cudaMalloc(&sink, sizeof(float) * 2 * 148 * 64 * 320); cudaMalloc(&out, sizeof(long long) * G);
  printf("G = %d CTAs x 320 threads\n", G);
  run<0>("0 product: proxy fences + red.release + ld.acquire", counter, sink, out, G, 0);
  run<1>("1 no proxy fences", counter, sink, out, G, 0);
  run<2>("2 relaxed polling + one acq_rel fence", counter, sink, out, G, 0);
  run<3>("3 atom.release, last arriver skips the poll", counter, sink, out, G, 0);
  run<4>("4 cooperative_groups grid.sync()", counter, sink, out, G, 0);
  run<5>("5 warp-wide polling", counter, sink, out, G, 0);
  run<6>("6 product + 64 fresh __stcg per thread before it", counter, sink, out, G, 0);
  return 0;
}
