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

// Dihedral source index: aug[i][j] = img[src(i,j)]  (op codes in adipose_b200.h)
ADP_DEVINL void d4_src(int op, int i, int j, int n, int &si, int &sj) {
  switch (op) {
    case 0: si = i; sj = j; break;
    case 1: si = j; sj = n - 1 - i; break;
    case 2: si = n - 1 - i; sj = n - 1 - j; break;
    case 3: si = n - 1 - j; sj = i; break;
    case 4: si = i; sj = n - 1 - j; break;
    case 5: si = n - 1 - i; sj = j; break;
    case 6: si = n - 1 - j; sj = n - 1 - i; break;
    default: si = j; sj = i; break;
  }
}
__host__ __device__ inline int d4_inverse(int op) {
  return op == 1 ? 3 : (op == 3 ? 1 : op);
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
