// Hashing, coordinate scaling and the open-addressing coordinate table (K1, K2, K3, a1).
// All kernels are HBM/L2-bound integer work: one element per thread, 16-byte coordinate loads,
// grid-stride loops sized in multiples of the SM count.
#include <cstdlib>
#include "common.cuh"

namespace ft3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int row_batch() {
  static int rb = 0;
  if (rb == 0) {
    const char* e = getenv("FT3D_ROWBATCH");
    rb = e ? atoi(e) : 2;
    if (rb != 1 && rb != 4) rb = 2;
  }
  return rb;
}

int col_cta_cap() {
  const char* e = getenv("FT3D_COL_CTAS");       // read per call: tools/bn_probe.py sweeps it inside one process
  int v = e ? atoi(e) : 0;
  if (v < 1 || v > kNumSMs * 8) v = kNumSMs * 8;
  return v;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("FT3D_PDL");
    on = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return on != 0;
}

// ------------------------------------------------------------------ K1
__global__ void hash_kernel(const int4* __restrict__ coords, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    out[i] = (int64_t)fnv1a_fold(c.x, c.y, c.z, c.w);
  }
}

// ------------------------------------------------------------------ K2  (offset-major output [K,n])
__global__ void kernel_hash_kernel(const int4* __restrict__ coords, int64_t n,
                                   const int32_t* __restrict__ offsets, int K,
                                   int64_t* __restrict__ out) {
  extern __shared__ int s_off[];
  for (int t = threadIdx.x; t < K * 3; t += blockDim.x) s_off[t] = offsets[t];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    for (int k = 0; k < K; ++k)
      out[(int64_t)k * n + i] =
          (int64_t)fnv1a_fold(c.x + s_off[3 * k], c.y + s_off[3 * k + 1], c.z + s_off[3 * k + 2], c.w);
  }
}

// ------------------------------------------------------------------ a1
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  // ordered-int trick, valid for any sign (no NaNs in LiDAR points)
  if (v == 0.f) v = 0.f;  // fold -0.0 onto +0.0 (its int pattern would sort below every negative)
  if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v));
  else atomicMax((unsigned int*)addr, __float_as_uint(v));
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void scan_min_kernel(const float* __restrict__ pts, const int32_t* __restrict__ scan_id,
                                int64_t n, float scale, float* __restrict__ mins) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = scan_id[i];
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = __fmul_rn(pts[i * 3 + d], scale);
      // warp-aggregate when a full warp sits in one scan (the common case)
      unsigned m = __activemask();
      bool uniform = false;
      if (m == 0xffffffffu) {
        int s0 = __shfl_sync(m, s, 0);
        uniform = __all_sync(m, s == s0);
      }
      if (uniform) {
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0) atomic_min_float(mins + s * 3 + d, v);
      } else {
        atomic_min_float(mins + s * 3 + d, v);
      }
    }
  }
}

__global__ void scale_coords_kernel(const float* __restrict__ pts, const int32_t* __restrict__ scan_id,
                                    int64_t n, float scale, int full_scale,
                                    const float* __restrict__ mins, int4* __restrict__ coords,
                                    uint8_t* __restrict__ keep) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = scan_id[i];
    int c[3];
    bool ok = true;
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
      // numpy: (points * scale) then (coords -= coords.min(0)), both rounded to fp32; no FMA contraction
      float v = __fsub_rn(__fmul_rn(pts[i * 3 + d], scale), mins[s * 3 + d]);
      c[d] = (int)v;  // astype(int64) truncation; values are >= 0
      ok = ok && (c[d] >= 0) && (c[d] < full_scale);
    }
    coords[i] = make_int4(c[0], c[1], c[2], s);
    keep[i] = ok ? 1 : 0;
  }
}

// ------------------------------------------------------------------ a1 with the augmentation branch
// data/utils/augmentation_3d.py:22-51.  The random numbers are drawn on the host in the reference's order (they are a
// dozen per scan); the device applies them to every point with numpy's float32 arithmetic: the [n,3] x [3,3] product
// as sgemm evaluates it (rounded product, then two fused multiply-adds, k ascending -- verified bit-exact against
// numpy/OpenBLAS), scale, per-scan minimum, and the translation offset computed in float64 from the float32 extent.
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v == 0.f) v = 0.f;
  if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
  else atomicMin((unsigned int*)addr, __float_as_uint(v));
}

__device__ __forceinline__ float aug_coord(const float* __restrict__ pts, int64_t i, const float* __restrict__ rot, int s,
                                           int d, float scale) {
  float v = pts[i * 3 + d];
  if (rot != nullptr) {
    const float* r = rot + s * 9;
    v = __fmul_rn(pts[i * 3], r[d]);
    v = __fmaf_rn(pts[i * 3 + 1], r[3 + d], v);
    v = __fmaf_rn(pts[i * 3 + 2], r[6 + d], v);
  }
  return __fmul_rn(v, scale);
}

__global__ void scan_minmax_kernel(const float* __restrict__ pts, const int32_t* __restrict__ scan_id, int64_t n,
                                   float scale, const float* __restrict__ rot, float* __restrict__ mins,
                                   float* __restrict__ maxs) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = scan_id[i];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = aug_coord(pts, i, rot, s, d, scale);
      atomic_min_float(mins + s * 3 + d, v);
      atomic_max_float(maxs + s * 3 + d, v);
    }
  }
}

__global__ void augment_scale_kernel(const float* __restrict__ pts, const int32_t* __restrict__ scan_id, int64_t n,
                                     float scale, int full_scale, const float* __restrict__ rot,
                                     const double* __restrict__ transl_u, const float* __restrict__ mins,
                                     const float* __restrict__ maxs, int4* __restrict__ coords,
                                     uint8_t* __restrict__ keep) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = scan_id[i];
    int c[3];
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float mn = mins[s * 3 + d];
      float v = __fsub_rn(aug_coord(pts, i, rot, s, d, scale), mn);                   // :43, :46
      if (transl_u != nullptr) {
        // :50-51  offset = clip(full_scale - coords.max(0) - 0.001, 0) * rand(3): float32 up to the product with the
        // float64 draw; coords += offset is evaluated in float64 and rounded back to float32
        const float mx = __fsub_rn(maxs[s * 3 + d], mn);                               // max of the shifted column
        float room = __fsub_rn(__fsub_rn((float)full_scale, mx), 0.001f);
        room = room > 0.f ? room : 0.f;
        v = (float)((double)v + (double)room * transl_u[s * 3 + d]);
      }
      c[d] = (int)v;
      ok = ok && (c[d] >= 0) && (c[d] < full_scale);
    }
    coords[i] = make_int4(c[0], c[1], c[2], s);
    keep[i] = ok ? 1 : 0;
  }
}

// ------------------------------------------------------------------ K10 helper
__device__ __forceinline__ int floor_to(int v, int r) {
  // torch: floor(floor(float(v)/r)*r); coordinates are < 2^24 so the float path is exact
  int q = v / r;
  if ((v % r != 0) && ((v < 0) != (r < 0))) --q;
  return q * r;
}

__global__ void coarsen_hash_kernel(const int4* __restrict__ coords, int64_t n, int ratio,
                                    int4* __restrict__ coarse, int64_t* __restrict__ hash) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    int4 q = make_int4(floor_to(c.x, ratio), floor_to(c.y, ratio), floor_to(c.z, ratio), c.w);
    coarse[i] = q;
    hash[i] = (int64_t)fnv1a_fold(q.x, q.y, q.z, q.w);
  }
}

__global__ void gather_rows_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ first,
                                       int64_t m, int width, int32_t* __restrict__ out) {
  int64_t total = m * width;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = t / width;
    int w = (int)(t - g * width);
    out[t] = src[(int64_t)first[g] * width + w];
  }
}

// ------------------------------------------------------------------ K3
__global__ void table_init_kernel(unsigned long long* __restrict__ tkeys, int* __restrict__ tvals, int64_t cap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
    tkeys[i] = kEmptyKey;
    tvals[i] = 0x7FFFFFFF;
  }
}

__global__ void table_insert_kernel(const int64_t* __restrict__ keys, int64_t n,
                                    unsigned long long* __restrict__ tkeys, int* __restrict__ tvals,
                                    uint32_t mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = (unsigned long long)keys[i];
    uint32_t s = slot_of(key, mask);
    for (uint32_t probe = 0; probe <= mask; ++probe) {
      unsigned long long prev = atomicCAS(tkeys + s, kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) {
        atomicMin(tvals + s, (int)i);   // duplicates: smallest row wins, deterministic
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

__global__ void table_query_kernel(const int64_t* __restrict__ q, int64_t m,
                                   const unsigned long long* __restrict__ tkeys,
                                   const int* __restrict__ tvals, uint32_t mask,
                                   int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int64_t)table_lookup(tkeys, tvals, mask, (unsigned long long)q[i]);
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

int ft3d_version(void) { return FT3D_VERSION; }
const char* ft3d_last_error(void) { return g_err; }

int ft3d_hash(const int32_t* coords, int64_t n, int64_t* out, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(coords && out && n > 0, "ft3d_hash: bad arguments");
  hash_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const int4*)coords, n, out);
  return check_launch("ft3d_hash");
}

int ft3d_kernel_hash(const int32_t* coords, int64_t n, const int32_t* offsets, int32_t K,
                     int64_t* out, ft3d_stream_t stream) {
  if (n == 0 || K == 0) return FT3D_OK;
  FT3D_REQUIRE(coords && offsets && out && n > 0 && K > 0 && K <= 4096, "ft3d_kernel_hash: bad arguments");
  kernel_hash_kernel<<<grid_for(n, 256), 256, K * 3 * sizeof(int), (cudaStream_t)stream>>>(
      (const int4*)coords, n, offsets, K, out);
  return check_launch("ft3d_kernel_hash");
}

int ft3d_scale_coords(const float* points, const int32_t* scan_id, int64_t n, int32_t num_scans,
                      float scale, int32_t full_scale, int32_t* coords_out, uint8_t* keep_out,
                      float* min_ws, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(points && scan_id && coords_out && keep_out && min_ws && num_scans > 0,
               "ft3d_scale_coords: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  fill_f32_kernel<<<1, 256, 0, s>>>(min_ws, (int64_t)num_scans * 3, INFINITY);
  scan_min_kernel<<<grid_for(n, 256), 256, 0, s>>>(points, scan_id, n, scale, min_ws);
  scale_coords_kernel<<<grid_for(n, 256), 256, 0, s>>>(points, scan_id, n, scale, full_scale, min_ws,
                                                       (int4*)coords_out, keep_out);
  return check_launch("ft3d_scale_coords");
}

int ft3d_augment_scale_coords(const float* points, const int32_t* scan_id, int64_t n, int32_t num_scans, float scale,
                              int32_t full_scale, const float* rot, const double* transl_u, int32_t* coords_out,
                              uint8_t* keep_out, float* ws, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(points && scan_id && coords_out && keep_out && ws && num_scans > 0,
               "ft3d_augment_scale_coords: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  float* mins = ws;
  float* maxs = ws + (int64_t)num_scans * 3;
  fill_f32_kernel<<<1, 256, 0, s>>>(mins, (int64_t)num_scans * 3, INFINITY);
  fill_f32_kernel<<<1, 256, 0, s>>>(maxs, (int64_t)num_scans * 3, -INFINITY);
  scan_minmax_kernel<<<grid_for(n, 256), 256, 0, s>>>(points, scan_id, n, scale, rot, mins, maxs);
  augment_scale_kernel<<<grid_for(n, 256), 256, 0, s>>>(points, scan_id, n, scale, full_scale, rot, transl_u, mins, maxs,
                                                        (int4*)coords_out, keep_out);
  return check_launch("ft3d_augment_scale_coords");
}

int ft3d_coarsen_hash(const int32_t* coords, int64_t n, int32_t ratio, int32_t* coarse_out,
                      int64_t* hash_out, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(coords && coarse_out && hash_out && ratio > 0, "ft3d_coarsen_hash: bad arguments");
  coarsen_hash_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const int4*)coords, n, ratio, (int4*)coarse_out, hash_out);
  return check_launch("ft3d_coarsen_hash");
}

int ft3d_gather_rows_i32(const int32_t* src, const int32_t* first, int64_t m, int32_t width,
                         int32_t* out, ft3d_stream_t stream) {
  if (m == 0) return FT3D_OK;
  FT3D_REQUIRE(src && first && out && width > 0, "ft3d_gather_rows_i32: bad arguments");
  gather_rows_i32_kernel<<<grid_for(m * width, 256), 256, 0, (cudaStream_t)stream>>>(src, first, m, width, out);
  return check_launch("ft3d_gather_rows_i32");
}

int64_t ft3d_table_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

int ft3d_table_build(const int64_t* keys, int64_t n, uint64_t* table_keys, int32_t* table_vals,
                     int64_t cap, ft3d_stream_t stream) {
  FT3D_REQUIRE(table_keys && table_vals && cap >= 2 && (cap & (cap - 1)) == 0 && cap <= (1LL << 31),
               "ft3d_table_build: capacity must be a power of two <= 2^31");
  FT3D_REQUIRE(n >= 0 && 2 * n <= cap, "ft3d_table_build: capacity %lld too small for %lld keys (need >= 2n)",
               (long long)cap, (long long)n);
  cudaStream_t s = (cudaStream_t)stream;
  table_init_kernel<<<grid_for(cap, 256), 256, 0, s>>>((unsigned long long*)table_keys, table_vals, cap);
  if (n > 0) {
    FT3D_REQUIRE(keys != nullptr, "ft3d_table_build: null keys");
    table_insert_kernel<<<grid_for(n, 256), 256, 0, s>>>(keys, n, (unsigned long long*)table_keys, table_vals,
                                                         (uint32_t)(cap - 1));
  }
  return check_launch("ft3d_table_build");
}

int ft3d_table_query(const int64_t* queries, int64_t m, const uint64_t* table_keys,
                     const int32_t* table_vals, int64_t cap, int64_t* out, ft3d_stream_t stream) {
  if (m == 0) return FT3D_OK;
  FT3D_REQUIRE(queries && table_keys && table_vals && out && cap >= 2 && (cap & (cap - 1)) == 0,
               "ft3d_table_query: bad arguments");
  table_query_kernel<<<grid_for(m, 256), 256, 0, (cudaStream_t)stream>>>(
      queries, m, (const unsigned long long*)table_keys, table_vals, (uint32_t)(cap - 1), out);
  return check_launch("ft3d_table_query");
}

}  // extern "C"
