// How fast can ONE SM pull L2-resident data into shared memory, and how does that scale with the number of SMs pulling?
// (The GEMM phases of the word-step kernels run at ~48 KB per 1700 cycles per SM, the attention at ~19 B/clk per SM.)
//   mode 0: cp.async.bulk global->shared in `chunk`-byte pieces, all pieces of a pass in flight, one mbarrier per pass
//   mode 1: ld.global.cg float4 (coalesced), 8 loads in flight per thread, summed
//   mode 2: 2-D TMA tensor boxes (64 x 128 B rows, swizzle 128B: the operand tiles of the GEMM phases)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o gpurun_out/l2_ingest scripts/microbench/l2_ingest.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int THREADS = 320;
constexpr int REGION = 160 * 1024;        // bytes per CTA per pass

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned a, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c)); }
__device__ __forceinline__ void expect_tx(unsigned a, unsigned b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(a), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(unsigned dst, const CUtensorMap* tm, unsigned bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) ingest_kernel(const float* __restrict__ src, const __grid_constant__ CUtensorMap tm, int mode, int chunk,
                                                          int iters, int active, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if ((int)blockIdx.x >= active) return;
  const char* mine = reinterpret_cast<const char*>(src) + (size_t)blockIdx.x * REGION;
  float acc = 0.f;
  long long t0 = 0;
  for (int it = -2; it < iters; ++it) {       // two warm-up passes bring the region into L2
    if (it == 0) t0 = clock64();
    if (mode == 0) {
      if (threadIdx.x == 0) {
        expect_tx(b, REGION);
        for (int o = 0; o < REGION; o += chunk) bulk_g2s(smem_u32(smem + o), mine + o, chunk, b);
      }
      mbar_wait(b, (unsigned)(it + 2) & 1);
    } else if (mode == 2) {
      if (threadIdx.x == 0) {
        expect_tx(b, REGION);
        // region = 1280 rows of 128 bytes; box = 64 rows x 128 B (8 KB)
        for (int r = 0; r < REGION / 128; r += 64) tma_2d(smem_u32(smem + r * 128), &tm, b, 0, (int)blockIdx.x * (REGION / 128) + r);
      }
      mbar_wait(b, (unsigned)(it + 2) & 1);
    } else {
      const float4* p = reinterpret_cast<const float4*>(mine);
      for (int i = threadIdx.x; i < REGION / 16; i += THREADS * 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int j = i + q * THREADS; v[q] = j < REGION / 16 ? __ldcg(p + j) : make_float4(0, 0, 0, 0); }
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += v[q].x + v[q].y + v[q].z + v[q].w;
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc + smem[threadIdx.x];
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* src; long long* cyc; float* sink;
  const size_t bytes = (size_t)sms * REGION;
  cudaMalloc(&src, bytes); cudaMemset(src, 0, bytes);
  cudaMalloc(&cyc, sizeof(long long) * sms); cudaMalloc(&sink, 4);
  CUtensorMap tm;
  {
    cuuint64_t dims[2] = {32, (cuuint64_t)(bytes / 128)};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); return 1; }
  }
  const int smem = REGION + 2048;
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 40;
  const char* names[3] = {"cp.async.bulk", "ld.global.cg float4", "TMA 2-D 8 KB boxes"};
  for (int mode = 0; mode < 3; ++mode)
    for (int chunk : {4096, 16384, 32768}) {
      if (mode != 0 && chunk != 4096) continue;
      for (int active : {1, 16, 64, sms}) {
        ingest_kernel<<<sms, THREADS, smem>>>(src, tm, mode, chunk, iters, active, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(sms);
        cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        std::sort(h.begin(), h.begin() + active);
        const double med = (double)h[active / 2] / iters, worst = (double)h[active - 1] / iters;
        printf("%-22s chunk %6d  active SMs %3d : %8.0f cycles per %d KB pass (median), %8.0f (slowest)  = %5.1f B/clk per SM\n", names[mode],
               mode == 0 ? chunk : 0, active, med, REGION / 1024, worst, REGION / med);
      }
    }
  return 0;
}
