// warp_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile dex_retargeting_b200/csrc/dexr_kernels.cuh (the solver the CUDA
// kernels instantiate) for the host: CUDA qualifiers become no-ops, the math intrinsics map to libm, and the warp
// collectives (__shfl_sync, __shfl_xor_sync, __ballot_sync, __syncwarp) become rendezvous points of 32 cooperatively
// scheduled fibers (emu_driver.cpp).  Between two collectives the lanes run ONE AFTER THE OTHER, so a cross-lane shared
// memory dependency that is not separated by a collective reads stale data here (stricter than the hardware), and lanes
// that reach different collectives, or a collective that some lanes never reach, abort with a message.
// Arithmetic differs from the GPU in the last bits (libm sincosf / sqrt, no MUFU approximations): this checks the LOGIC of
// the solver source -- including the compile-time experiment switches -- against the oracle, not bit patterns.
#pragma once
#define __CUDA_RUNTIME_H__  // the header's #include <cuda_runtime.h> becomes a no-op
#include <cmath>
#include <cstdint>
#include <cstring>

#define DEXR_HOST_EMULATION 1
#define __device__
#define __host__
#define __global__
#define __shared__
#define __forceinline__ inline __attribute__((always_inline))
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

using std::isfinite;

struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct emu_dim3 { unsigned x, y, z; };
extern emu_dim3 threadIdx, blockDim, blockIdx, gridDim;  // only load_shared_table looks at them (run as one thread)

namespace emu {
enum Op { OP_SHFL = 1, OP_BALLOT = 2, OP_SYNC = 3 };
int lane();                                   // lane of the running fiber
const uint32_t* rendezvous(int op, uint32_t value);  // returns the 32 values of this round
}  // namespace emu

template <typename T>
static inline T emu_shfl(T v, int src_lane_abs) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  const uint32_t* all = emu::rendezvous(emu::OP_SHFL, bits);
  T out;
  std::memcpy(&out, &all[src_lane_abs & 31], 4);
  return out;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int l = emu::lane();
  return emu_shfl(v, (l & ~(width - 1)) + (src & (width - 1)));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32) {
  const int l = emu::lane();
  const int src = l ^ lane_mask;
  return emu_shfl(v, ((src & ~(width - 1)) == (l & ~(width - 1))) ? src : l);
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
  const uint32_t* all = emu::rendezvous(emu::OP_BALLOT, pred ? 1u : 0u);
  unsigned b = 0;
  for (int i = 0; i < 32; ++i) b |= (all[i] & 1u) << i;
  return b;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::rendezvous(emu::OP_SYNC, 0u); }

static inline void emu_fast_sincosf(float x, float* s, float* c) { sincosf(x, s, c); }
#define __sincosf emu_fast_sincosf  // glibc already declares a __sincosf
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
