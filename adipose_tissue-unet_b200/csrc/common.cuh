// common.cuh — shared host/device helpers for libadipose_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#define ADP_DEVINL __device__ __forceinline__

namespace adp {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define ADP_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw adp::Error(-2, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                               std::to_string(__LINE__) + ")");                                    \
  } while (0)

#define ADP_REQUIRE(cond, msg)                                     \
  do {                                                             \
    if (!(cond)) throw adp::Error(-1, std::string("invalid argument: ") + (msg)); \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int pad16(int c) { return (c + 15) / 16 * 16; }

// ---- element type helpers (activations are float or bf16, accumulation always fp32) ----
template <typename T> ADP_DEVINL float to_f(T v);
template <> ADP_DEVINL float to_f<float>(float v) { return v; }
template <> ADP_DEVINL float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> ADP_DEVINL T from_f(float v);
template <> ADP_DEVINL float from_f<float>(float v) { return v; }
template <> ADP_DEVINL __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Counter-based dropout hash (kernels_train.cuh: dropout kernels; conv_tc.cuh: fused into the conv epilogue)
__host__ __device__ inline uint32_t hash_u32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
// salt = hash(high half of the element index ^ high half of the seed) ^ low half of the seed: constant over a launch whenever
// the index fits 32 bits (hoisted out of the element loop - the kernel was issue-bound on two hashes per channel pair)
__host__ __device__ inline uint32_t dropout_salt(uint64_t seed, uint32_t idx_hi) { return hash_u32(idx_hi ^ (uint32_t)(seed >> 32)) ^ (uint32_t)seed; }

// Dihedral source index: aug[i][j] = img[src(i,j)]  (op codes in adipose_b200.h)
ADP_DEVINL void d4_src(int op, int i, int j, int n, int &si, int &sj) {
  // branch-free: 3 bits per op = (transpose, mirror the row index, mirror the column index)
  //   0: (i, j)         1: (j, n-1-i)     2: (n-1-i, n-1-j)  3: (n-1-j, i)
  //   4: (i, n-1-j)     5: (n-1-i, j)     6: (n-1-j, n-1-i)  7: (j, i)
  // (an 8-way switch compiles to ~16 predicated compares per call, which made the TTA kernel instruction-bound)
  constexpr unsigned kTable = (0u << 0) | (5u << 3) | (6u << 6) | (3u << 9) | (4u << 12) | (2u << 15) | (7u << 18) | (1u << 21);
  const unsigned b = (kTable >> (3 * (op & 7))) & 7u;
  const int u = (b & 1u) ? j : i, v = (b & 1u) ? i : j;
  si = (b & 2u) ? n - 1 - u : u;
  sj = (b & 4u) ? n - 1 - v : v;
}
__host__ __device__ inline int d4_inverse(int op) {
  return op == 1 ? 3 : (op == 3 ? 1 : op);
}

// Flat item index -> (image n, row y, channel group g, column x) of a row-planar tensor walked as [n][y][g][x].
// 32-bit divisions whenever the index fits (it always does at the path's sizes: <= 2^26 items per launch): the 64-bit
// form costs ~300 integer instructions per 16..48-byte item, which capped the elementwise kernels near 4.8 TB/s.
struct RpIndex { int n, y, g, x; };
ADP_DEVINL RpIndex rp_index(size_t i, int W, int G, int H) {
  RpIndex r;
  if (i <= 0xFFFFFFFFull) {
    uint32_t j = (uint32_t)i;
    if (((W & (W - 1)) | (H & (H - 1))) == 0) {          // power-of-two width and height (every size of the path): one division
      r.x = (int)(j & (uint32_t)(W - 1)); j >>= (__ffs(W) - 1);
      const uint32_t q2 = j / (uint32_t)G; r.g = (int)(j - q2 * (uint32_t)G);
      r.y = (int)(q2 & (uint32_t)(H - 1)); r.n = (int)(q2 >> (__ffs(H) - 1));
      return r;
    }
    const uint32_t q = j / (uint32_t)W; r.x = (int)(j - q * (uint32_t)W); j = q;
    const uint32_t q2 = j / (uint32_t)G; r.g = (int)(j - q2 * (uint32_t)G); j = q2;
    const uint32_t q3 = j / (uint32_t)H; r.y = (int)(j - q3 * (uint32_t)H); r.n = (int)q3;
  } else {
    r.x = (int)(i % W); size_t t = i / W;
    r.g = (int)(t % G); t /= G;
    r.y = (int)(t % H); r.n = (int)(t / H);
  }
  return r;
}

ADP_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ADP_DEVINL double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ADP_DEVINL unsigned warp_sum_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace adp
