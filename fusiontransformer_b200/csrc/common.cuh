// Shared helpers for libft3d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>

#include "../../include/ft3d.h"

#define FT3D_OK 0
#define FT3D_ERR_ARG 1
#define FT3D_ERR_CUDA 2
#define FT3D_ERR_WORKSPACE 3

namespace ft3d {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FT3D_ERR_CUDA;
  }
  return FT3D_OK;
}

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FT3D_ERR_CUDA;
  }
  return FT3D_OK;
}

#define FT3D_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ft3d::set_error(__VA_ARGS__);        \
      return FT3D_ERR_ARG;                 \
    }                                      \
  } while (0)

#define FT3D_CUDA(expr)                                            \
  do {                                                             \
    int _rc = ft3d::check_cuda((expr), #expr);                     \
    if (_rc) return _rc;                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL).  Every kernel launched through launch_pdl() starts with pdl_enter():
// griddepcontrol.wait blocks until the preceding kernel of the stream has completed and flushed, THEN
// launch_dependents lets the next kernel's CTAs be scheduled and run their prologue while this one executes.  Because
// a kernel triggers only after its own wait, a dependent that has started knows everything up to its
// grand-predecessor is complete; it still waits for its predecessor before touching memory.  Launch latency and CTA
// ramp-up of the ~700 short, strictly dependent kernels of a training step overlap instead of adding up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}

int row_batch();    // FT3D_ROWBATCH = 1 | 2 | 4: rows a thread of the streaming kernels keeps in flight (default 2, measured best on B200)
int col_cta_cap();  // FT3D_COL_CTAS: upper bound on the CTAs of a column-reduction / BatchNorm streaming launch (<= 8 per SM)
bool pdl_enabled();   // FT3D_PDL=0 turns the launch attribute off (the device-side instructions are then no-ops)

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through check_launch()
}

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// grid for a grid-stride loop over n items: enough CTAs for ~4 resident per SM, capped by the work.
static inline int grid_for(int64_t n, int block, int per_sm = 8) {
  int64_t need = (n + block - 1) / block;
  int64_t cap = (int64_t)kNumSMs * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- FNV-1a 64 over four 32-bit words, folded to 60 bits (torchsparse v1.1.0 `hash_kernel`)
__host__ __device__ __forceinline__ uint64_t fnv1a_fold(int x, int y, int z, int b) {
  uint64_t h = 14695981039346656037ULL;
  h ^= (uint32_t)x; h *= 1099511628211ULL;
  h ^= (uint32_t)y; h *= 1099511628211ULL;
  h ^= (uint32_t)z; h *= 1099511628211ULL;
  h ^= (uint32_t)b; h *= 1099511628211ULL;
  return (h >> 60) ^ (h & 0x0FFFFFFFFFFFFFFFULL);
}

// ---- FNV-1 64 (multiply then xor) over three 64-bit words, unfolded (torchsparse `fnv_hash_vec`)
__host__ __device__ __forceinline__ uint64_t fnv1_vec3(int64_t x, int64_t y, int64_t z) {
  uint64_t h = 14695981039346656037ULL;
  h *= 1099511628211ULL; h ^= (uint64_t)x;
  h *= 1099511628211ULL; h ^= (uint64_t)y;
  h *= 1099511628211ULL; h ^= (uint64_t)z;
  return h;
}

// ---- open-addressing table: slot mixer (murmur3 finaliser) -- table layout is private to libft3d
__device__ __forceinline__ uint32_t slot_of(uint64_t key, uint32_t mask) {
  key ^= key >> 33; key *= 0xff51afd7ed558ccdULL;
  key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ULL;
  key ^= key >> 33;
  return (uint32_t)key & mask;
}
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFULL;

__device__ __forceinline__ int table_lookup(const unsigned long long* __restrict__ tkeys,
                                            const int* __restrict__ tvals, uint32_t mask,
                                            unsigned long long key) {
  uint32_t s = slot_of(key, mask);
  #pragma unroll 1
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    unsigned long long k = __ldg(tkeys + s);
    if (k == key) return __ldg(tvals + s);
    if (k == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
  return -1;
}

}  // namespace ft3d
