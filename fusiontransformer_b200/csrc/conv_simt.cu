// fp32 CUDA-core sparse convolution (exact-precision mode, FT3D_CONV=f32).
// Same output-stationary formulation and the same maps as the tcgen05 path in conv_tc.cu; used for
// 1e-5 parity against the oracle and as the on-device cross-check of the bf16 tensor-core kernels.
#include "common.cuh"

namespace ft3d {

constexpr int kMaxColsPerLane = 12;  // ncols <= 384

// one warp per output row; lanes stride over output columns; weights stream through L1/L2.
__global__ void __launch_bounds__(128)
conv_gather_f32_kernel(const float* __restrict__ in, const int32_t* __restrict__ nbr, int64_t n_out, int K,
                       int kpad, int kflip, int red, int ncols, const float* __restrict__ w, int w_transposed,
                       float* __restrict__ out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= n_out) return;
  float acc[kMaxColsPerLane];
  #pragma unroll
  for (int t = 0; t < kMaxColsPerLane; ++t) acc[t] = 0.f;
  if (red == 4 && K <= 32 && (((uintptr_t)in) & 15) == 0) {
    // 4-channel input (the stem convolution, models/spvcnn.py:87-93): lane k fetches neighbour k's index and then its
    // whole 16-byte feature row, so the warp makes two dependent round trips in total instead of one per offset; the
    // offsets are then walked in ascending order through shuffles (same summation order as the general loop below).
    const int src = lane < K ? __ldg(nbr + row * kpad + lane) : -1;
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src >= 0) a4 = __ldg(reinterpret_cast<const float4*>(in + (int64_t)src * 4));
    unsigned mask = __ballot_sync(0xffffffffu, src >= 0);
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      const float av[4] = {__shfl_sync(0xffffffffu, a4.x, k), __shfl_sync(0xffffffffu, a4.y, k),
                           __shfl_sync(0xffffffffu, a4.z, k), __shfl_sync(0xffffffffu, a4.w, k)};
      const int kw = kflip ? (K - 1 - k) : k;
      const float* wk = w + (int64_t)kw * 4 * ncols;
      #pragma unroll
      for (int r = 0; r < 4; ++r) {
        #pragma unroll
        for (int t = 0; t < kMaxColsPerLane; ++t) {
          const int c = lane + 32 * t;
          if (c < ncols) {
            const float bv = w_transposed ? __ldg(wk + (int64_t)c * 4 + r) : __ldg(wk + (int64_t)r * ncols + c);
            acc[t] = fmaf(av[r], bv, acc[t]);
          }
        }
      }
    }
    #pragma unroll
    for (int t = 0; t < kMaxColsPerLane; ++t) {
      const int c = lane + 32 * t;
      if (c < ncols) out[row * ncols + c] = acc[t];
    }
    return;
  }
  for (int k = 0; k < K; ++k) {
    int src = __ldg(nbr + row * kpad + k);
    if (src < 0) continue;
    const int kw = kflip ? (K - 1 - k) : k;
    const float* a = in + (int64_t)src * red;
    const float* wk = w + (int64_t)kw * red * ncols;
    for (int r = 0; r < red; ++r) {
      float av = __ldg(a + r);
      #pragma unroll
      for (int t = 0; t < kMaxColsPerLane; ++t) {
        int c = lane + 32 * t;
        if (c < ncols) {
          float bv = w_transposed ? __ldg(wk + (int64_t)c * red + r) : __ldg(wk + (int64_t)r * ncols + c);
          acc[t] = fmaf(av, bv, acc[t]);
        }
      }
    }
  }
  #pragma unroll
  for (int t = 0; t < kMaxColsPerLane; ++t) {
    int c = lane + 32 * t;
    if (c < ncols) out[row * ncols + c] = acc[t];
  }
}

constexpr int kWgradChunk = 64;

// work item -> (offset k, pair range) from the device-side prefix array
__device__ __forceinline__ bool wgrad_item(const int32_t* __restrict__ off, int K, int chunk, int item, int* k_out,
                                           int* begin, int* end) {
  int acc = 0;
  for (int k = 0; k < K; ++k) {
    int b = __ldg(off + k), e = __ldg(off + k + 1);
    int nc = (e - b + chunk - 1) / chunk;
    if (item < acc + nc) {
      *k_out = k;
      *begin = b + (item - acc) * chunk;
      *end = min(*begin + chunk, e);
      return true;
    }
    acc += nc;
  }
  return false;
}

__global__ void __launch_bounds__(256)
conv_wgrad_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, const int2* __restrict__ pairs,
                      const int32_t* __restrict__ off, int K, int ca, int cin, int cout, float* __restrict__ gw) {
  pdl_enter();
  __shared__ int s_a[kWgradChunk], s_b[kWgradChunk];
  int k, begin, end;
  if (!wgrad_item(off, K, kWgradChunk, blockIdx.x, &k, &begin, &end)) return;
  const int np = end - begin;
  for (int t = threadIdx.x; t < np; t += blockDim.x) {
    int2 p = __ldg(pairs + begin + t);
    s_a[t] = ca ? p.y : p.x;
    s_b[t] = ca ? p.x : p.y;
  }
  __syncthreads();
  const int total = cin * cout;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int ci = e / cout, co = e - ci * cout;
    float s = 0.f;
    for (int t = 0; t < np; ++t)
      s = fmaf(__ldg(a + (int64_t)s_a[t] * cin + ci), __ldg(b + (int64_t)s_b[t] * cout + co), s);
    atomicAdd(gw + ((int64_t)k * cin + ci) * cout + co, s);
  }
}

// Deterministic variant (FT3D_DETERMINISTIC): a thread owns ONE element gw[k, ci, co] and walks all pairs of offset k
// in list order -- no atomics, no workspace; slower (a serial walk of L_k pairs per thread), used for the exact-fp32
// layers (the 4-channel stem in tensor-core mode).
__global__ void __launch_bounds__(128)
conv_wgrad_f32_det_kernel(const float* __restrict__ a, const float* __restrict__ b, const int2* __restrict__ pairs,
                          const int32_t* __restrict__ off, int ca, int cin, int cout, float* __restrict__ gw) {
  pdl_enter();
  const int k = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cin * cout) return;
  const int ci = e / cout, co = e - ci * cout;
  const int begin = __ldg(off + k), end = __ldg(off + k + 1);
  float s = 0.f;
  for (int t = begin; t < end; ++t) {
    const int2 p = __ldg(pairs + t);
    s = fmaf(__ldg(a + (int64_t)(ca ? p.y : p.x) * cin + ci), __ldg(b + (int64_t)(ca ? p.x : p.y) * cout + co), s);
  }
  gw[((int64_t)k * cin + ci) * cout + co] += s;
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

int ft3d_conv_gather_f32(const float* in, const int32_t* nbr, int64_t n_out, int32_t K, int32_t kpad,
                         int32_t kflip, int32_t red, int32_t ncols, const float* w, int32_t w_transposed,
                         float* out, ft3d_stream_t stream) {
  if (n_out == 0) return FT3D_OK;
  FT3D_REQUIRE(in && nbr && w && out, "ft3d_conv_gather_f32: null pointer");
  FT3D_REQUIRE(K > 0 && K <= kpad && red > 0 && ncols > 0 && ncols <= 32 * kMaxColsPerLane,
               "ft3d_conv_gather_f32: unsupported shape K=%d kpad=%d red=%d ncols=%d", K, kpad, red, ncols);
  launch_pdl(conv_gather_f32_kernel, dim3((unsigned)((n_out + 3) / 4)), dim3(128), 0, (cudaStream_t)stream, in, nbr, n_out, K, kpad, kflip, red, ncols, w, w_transposed, out);
  return check_launch("ft3d_conv_gather_f32");
}

int ft3d_conv_wgrad_f32(const float* a, const float* b, const int32_t* pairs, const int32_t* pair_offsets,
                        int32_t K, int32_t ca, int32_t cin, int32_t cout, int64_t max_pairs, float* gw,
                        ft3d_stream_t stream) {
  if (max_pairs == 0) return FT3D_OK;
  FT3D_REQUIRE(a && b && pairs && pair_offsets && gw && K > 0 && cin > 0 && cout > 0,
               "ft3d_conv_wgrad_f32: bad arguments");
  int64_t items = (max_pairs + kWgradChunk - 1) / kWgradChunk + K;
  launch_pdl(conv_wgrad_f32_kernel, dim3((unsigned)items), dim3(256), 0, (cudaStream_t)stream, a, b, (const int2*)pairs, pair_offsets,
                                                                            K, ca, cin, cout, gw);
  return check_launch("ft3d_conv_wgrad_f32");
}

int ft3d_conv_wgrad_f32_det(const float* a, const float* b, const int32_t* pairs, const int32_t* pair_offsets,
                            int32_t K, int32_t ca, int32_t cin, int32_t cout, int64_t max_pairs, float* gw,
                            ft3d_stream_t stream) {
  if (max_pairs == 0) return FT3D_OK;
  FT3D_REQUIRE(a && b && pairs && pair_offsets && gw && K > 0 && cin > 0 && cout > 0,
               "ft3d_conv_wgrad_f32_det: bad arguments");
  launch_pdl(conv_wgrad_f32_det_kernel, dim3((unsigned)((cin * cout + 127) / 128), (unsigned)K), dim3(128), 0,
             (cudaStream_t)stream, a, b, (const int2*)pairs, pair_offsets, (int)ca, (int)cin, (int)cout, gw);
  return check_launch("ft3d_conv_wgrad_f32_det");
}

}  // extern "C"
