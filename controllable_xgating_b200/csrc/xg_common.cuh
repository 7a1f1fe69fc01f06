// Common device/host helpers for libxgating (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/xgating.h"

namespace xg {

// ------------------------------------------------------------------------------------
// error plumbing: no exceptions cross the C ABI; every launcher returns xg_status.
// ------------------------------------------------------------------------------------
struct ErrorSink {
  std::string msg;
  void set(const char* file, int line, const char* what, const char* detail) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s:%d: %s%s%s", file, line, what, detail ? ": " : "", detail ? detail : "");
    msg = buf;
  }
};

#define XG_CUDA_TRY(sink, expr)                                              \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      (sink).set(__FILE__, __LINE__, #expr, cudaGetErrorString(_e));         \
      return XG_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

#define XG_LAUNCH_CHECK(sink)                                                \
  do {                                                                       \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) {                                                 \
      (sink).set(__FILE__, __LINE__, "kernel launch", cudaGetErrorString(_e)); \
      return XG_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

#define XG_TRY(expr)                      \
  do {                                    \
    int _s = (expr);                      \
    if (_s != XG_OK) return _s;           \
  } while (0)

#define XG_REQUIRE(sink, cond, code, text)                 \
  do {                                                     \
    if (!(cond)) {                                         \
      (sink).set(__FILE__, __LINE__, text, #cond);         \
      return (code);                                       \
    }                                                      \
  } while (0)

// ------------------------------------------------------------------------------------
// bump allocator over a caller-provided workspace
// ------------------------------------------------------------------------------------
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  bool overflow;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0), overflow(false) {}
  static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (base == nullptr) {  // sizing pass
      off += bytes;
      return nullptr;
    }
    if (off + bytes > cap) {
      overflow = true;
      off += bytes;
      return nullptr;
    }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

// ------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case XG_ACT_RELU: return v > 0.f ? v : 0.f;
    case XG_ACT_TANH: return tanhf(v);
    case XG_ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------
// counter-based dropout: Philox4x32-10 keyed by the step seed; counter = (element index, site).
// The mask of logical element `idx` at `site` is a pure function of (seed, site, idx), so the
// backward pass regenerates it instead of storing it, and tests can replay it in the oracle.
// ------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint64_t p0 = (uint64_t)M0 * c[0];
  uint64_t p1 = (uint64_t)M1 * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  uint32_t n0 = hi1 ^ c[1] ^ k0;
  uint32_t n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__host__ __device__ __forceinline__ uint32_t philox_word(uint64_t seed, uint32_t site, uint64_t idx) {
  uint32_t c[4] = {(uint32_t)(idx >> 2), (uint32_t)(idx >> 34), site, 0x58474154u /* "XGAT" */};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c[idx & 3];
}

// uniform in [0,1) with 24 bits
__host__ __device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t site, uint64_t idx) {
  return (float)(philox_word(seed, site, idx) >> 8) * (1.0f / 16777216.0f);
}

// multiplicative dropout factor: 0 with probability p, else 1/(1-p)
__host__ __device__ __forceinline__ float drop_factor(uint64_t seed, uint32_t site, uint64_t idx, float p,
                                                      float keep_scale) {
  return philox_uniform(seed, site, idx) >= p ? keep_scale : 0.0f;
}

struct DropSpec {
  uint64_t seed;
  uint32_t site;
  float p;           // 0 => disabled
  float keep_scale;  // 1/(1-p)
  uint64_t base;     // added to the logical element index
  __host__ __device__ bool on() const { return p > 0.f; }
  __host__ __device__ float factor(uint64_t idx) const {
    return on() ? drop_factor(seed, site, base + idx, p, keep_scale) : 1.0f;
  }
};

inline DropSpec make_drop(bool train, float p, uint64_t seed, uint32_t site, uint64_t base = 0) {
  DropSpec d;
  d.seed = seed;
  d.site = site;
  d.p = (train && p > 0.f) ? p : 0.f;
  d.keep_scale = d.p > 0.f ? 1.0f / (1.0f - d.p) : 1.0f;
  d.base = base;
  return d;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }

}  // namespace xg
