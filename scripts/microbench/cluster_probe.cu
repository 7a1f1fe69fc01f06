// Probe for the cluster split-K word-step design: can 16 clusters of 8 CTAs (320 threads, 207 KB
// dynamic shared memory each) be co-resident on a B200, does a cooperative + cluster launch work, and what do the
// building blocks cost: cluster barrier, the reduce-scatter of eight 128 x 64 fp32 partial tiles through distributed
// shared memory (pull: ld.shared::cluster; push: st.shared::cluster), a grid barrier over 128 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/cluster_probe scripts/microbench/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int THREADS = 320;
constexpr int SMEM = 207 * 1024;

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ float ld_cluster(unsigned addr) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void st_cluster(unsigned addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

__device__ void grid_barrier(unsigned* counter, unsigned& target, int G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += (unsigned)G;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
  }
  __syncthreads();
}

// mode 0: cluster barrier only; 1: pull reduce-scatter; 2: push reduce-scatter; 3: grid barrier; 4: pull + cell-like math + grid barrier
template <int CS>
__global__ void __launch_bounds__(THREADS, 1) probe_kernel(int mode, int iters, unsigned* counter, float* out, int* placement, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* part = reinterpret_cast<float*>(smem);                    // [128 rows][64 captions] partial tile (32 KB)
  float* recv = reinterpret_cast<float*>(smem + 64 * 1024);        // push target: [CS][4 gates][4 units][64] (32 KB at CS=8)
  const unsigned rank = cluster_ctarank();
  if (threadIdx.x == 0) { placement[blockIdx.x * 2] = (int)smid(); placement[blockIdx.x * 2 + 1] = (int)rank; }
  for (int i = threadIdx.x; i < 128 * 64; i += THREADS) part[i] = (float)(i % 97) * 0.01f + rank;
  __syncthreads();
  cluster_arrive(); cluster_wait();
  unsigned target = 0;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      cluster_arrive(); cluster_wait();
    } else if (mode == 1 || mode == 4) {
      cluster_arrive(); cluster_wait();                               // partial tiles complete everywhere
      if (threadIdx.x < 256) {
        const int u = threadIdx.x >> 6, c = threadIdx.x & 63;
        float v[4][CS];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = g * 32 + (int)rank * (32 / CS) + (u % (32 / CS));
          const unsigned a = smem_u32(part + n * 64 + c);
#pragma unroll
          for (int k = 0; k < CS; ++k) v[g][k] = ld_cluster(mapa(a, k));
        }
        float z[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) { z[g] = 0.f;
#pragma unroll
          for (int k = 0; k < CS; ++k) z[g] += v[g][k]; }
        if (mode == 4) {
          const float ig = 1.f / (1.f + __expf(-z[0])), fg = 1.f / (1.f + __expf(-z[1])), og = 1.f / (1.f + __expf(-z[2]));
          const float gg = 1.f - 2.f / (1.f + __expf(2.f * z[3]));
          const float c2 = fg * acc + ig * gg;
          acc = og * (1.f - 2.f / (1.f + __expf(2.f * c2)));
          out[(blockIdx.x * 256 + threadIdx.x)] = acc;
        } else {
          acc += z[0] + z[1] + z[2] + z[3];
        }
      }
      cluster_arrive(); cluster_wait();                               // reads done before the tiles are overwritten
      if (mode == 4) grid_barrier(counter, target, gridDim.x);
    } else if (mode == 2) {
      // every thread < 128 owns tile row n (64 captions): push to the CTA that owns that row's hidden unit
      if (threadIdx.x < 128) {
        const int n = threadIdx.x, g = n >> 5, uu = n & 31, owner = uu / (32 / CS), ul = uu % (32 / CS);
        const unsigned dst = mapa(smem_u32(recv + (((int)rank * 4 + g) * (32 / CS) + ul) * 64), owner);
#pragma unroll 16
        for (int c = 0; c < 64; ++c) st_cluster(dst + 4 * c, part[n * 64 + c]);
      }
      cluster_arrive(); cluster_wait();
      if (threadIdx.x < 256) {
        const int u = threadIdx.x >> 6, c = threadIdx.x & 63;
        float z = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int k = 0; k < CS; ++k) z += recv[((k * 4 + g) * (32 / CS) + (u % (32 / CS))) * 64 + c];
        acc += z;
      }
      cluster_arrive(); cluster_wait();
    } else if (mode == 3) {
      grid_barrier(counter, target, gridDim.x);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) out[0] = acc;
}

template <int CS>
int run(int grid, bool coop) {
  cudaFuncSetAttribute(probe_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  if (CS > 8) cudaFuncSetAttribute(probe_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = 0;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
  int nclusters = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, probe_kernel<CS>, &cfg);
  printf("cluster size %2d, grid %3d, coop %d: max active clusters %d (%s) -> %d CTAs\n", CS, grid, (int)coop, nclusters, cudaGetErrorString(e), nclusters * CS);
  if (e != cudaSuccess || nclusters * CS < grid) { cudaGetLastError(); return 0; }
  unsigned* counter; float* out; int* place; long long* cyc;
  cudaMalloc(&counter, 4); cudaMalloc(&out, 4 * 256 * 256); cudaMalloc(&place, 8 * 256); cudaMalloc(&cyc, 8 * 256);
  const char* names[5] = {"cluster barrier", "pull reduce-scatter (2 cluster barriers + 32 DSMEM loads/thread)", "push reduce-scatter (64 DSMEM stores/thread + 2 cluster barriers)",
                          "grid barrier", "pull reduce + cell + grid barrier"};
  for (int mode = 0; mode < 5; ++mode) {
    const int iters = 2000;
    cudaMemset(counter, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int warm = 10;
    void* args1[6] = {(void*)&mode, (void*)&warm, (void*)&counter, (void*)&out, (void*)&place, (void*)&cyc};
    e = cudaLaunchKernelExC(&cfg, (void*)probe_kernel<CS>, args1);
    if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return 0; }
    cudaDeviceSynchronize();
    cudaMemset(counter, 0, 4);
    int it2 = iters;
    void* args[6] = {(void*)&mode, (void*)&it2, (void*)&counter, (void*)&out, (void*)&place, (void*)&cyc};
    cudaEventRecord(e0);
    e = cudaLaunchKernelExC(&cfg, (void*)probe_kernel<CS>, args);
    cudaEventRecord(e1);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  run failed: %s\n", cudaGetErrorString(e)); return 0; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("  mode %d %-70s %8.0f ns per iteration\n", mode, names[mode], ms * 1e6 / iters);
  }
  int h[512]; cudaMemcpy(h, place, sizeof(int) * 2 * grid, cudaMemcpyDeviceToHost);
  printf("  placement (cta: smid/rank):");
  for (int i = 0; i < grid && i < 24; ++i) printf(" %d:%d/%d", i, h[2 * i], h[2 * i + 1]);
  printf("\n");
  return 1;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<8>(128, true);
  run<8>(128, false);
  run<8>(144, false);
  run<4>(128, true);
  run<4>(148, false);
  run<2>(148, true);
  run<16>(128, false);
  return 0;
}
