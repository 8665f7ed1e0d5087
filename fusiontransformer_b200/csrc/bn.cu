// spnn.BatchNorm + spnn.ReLU + residual add fused around the sparse convolution (a11; models/spvcnn.py:26-31,57-78).
//
// The reference runs conv -> BatchNorm1d -> ReLU (-> add -> ReLU) as separate ATen kernels: three or more full
// read+write passes over every activation plus a separate fp32->bf16 pass for the next convolution's operands.
// Here:
//   forward   statistics are produced by the convolution's sorted scatter (conv_pairs_tc.cu) or by bn_stats_kernel;
//             bn_apply_kernel then makes ONE pass: normalise, affine, (+ residual), (ReLU), and writes the fp32
//             activation and the bf16 copy the next convolution gathers from.
//   backward  bn_bwd_reduce_kernel (one pass: ReLU mask, sum g, sum g*xhat -> dgamma, dbeta) and bn_bwd_apply_kernel
//             (one pass: dx, written as the bf16 operand of dgrad/wgrad, plus the masked gradient of the residual).
// All are HBM-bound streaming kernels: float4 per thread, channel-fastest so a warp touches contiguous 512 B.
#include "bn_common.cuh"
#include "tc_common.cuh"

namespace ft3d {

using tc::pack_bf16x2;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 bf16x4_to_float4(uint2 v) {
  float4 f;
  f.x = __uint_as_float(v.x << 16);
  f.y = __uint_as_float(v.x & 0xFFFF0000u);
  f.z = __uint_as_float(v.y << 16);
  f.w = __uint_as_float(v.y & 0xFFFF0000u);
  return f;
}

// ------------------------------------------------------------------------------------------------ statistics
template <int RB>
__global__ void __launch_bounds__(kColThreads)
bn_stats_kernel(const float* __restrict__ y, int64_t n, int channels, int rows_per_cta,
                float* __restrict__ partials, const int32_t* __restrict__ valid_rows,
                const float* __restrict__ pivot) {
  pdl_enter();
  __shared__ float4 s_stage[kColStageFloat4];
  n = effective_rows(n, valid_rows);
  const int ch = threadIdx.x * 4;
  const float4 pv = stat_pivot(pivot, ch);
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row1 = row0 + rows_per_cta < n ? row0 + rows_per_cta : n;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const int64_t rstep = blockDim.y;
  for (int64_t r = row0 + threadIdx.y; r < row1; r += RB * rstep) {      // four independent row loads in flight
    float4 v[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i)
      v[i] = r + i * rstep < row1 ? ld4(y + (r + i * rstep) * channels + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (r + i * rstep < row1) stat_add(s1, s2, v[i], pv);
  }
  col_publish(s1, s2, partials, channels, s_stage);
}

// ------------------------------------------------------------------------------------------------ forward apply
// z = [relu]( (y - mean) * rstd * gamma + beta [+ res] );  writes z (fp32, nullable) and z16 (bf16, nullable).
// Same 2-D thread layout as the reductions (a thread owns 4 channels: its statistics and affine parameters are loaded
// once) and four rows in flight per thread.
template <int RB>
__global__ void __launch_bounds__(kColThreads)
bn_apply_kernel(const float* __restrict__ y, int64_t n, int channels, int rows_per_cta, const float* __restrict__ stat,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ res,
                int relu, float* __restrict__ z, __nv_bfloat16* __restrict__ z16, int64_t ldz,
                const int32_t* __restrict__ valid_rows) {
  pdl_enter();
  const int64_t nv = effective_rows(n, valid_rows);
  const int ch = threadIdx.x * 4;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row1 = row0 + rows_per_cta < n ? row0 + rows_per_cta : n;
  const float4 m = ld4(stat + ch), rs = ld4(stat + channels + ch), g = ld4(gamma + ch), b = ld4(beta + ch);
  const int64_t rstep = blockDim.y;
  for (int64_t r = row0 + threadIdx.y; r < row1; r += RB * rstep) {
    float4 v[RB], rr[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t row = r + i * rstep;
      const bool live = row < row1 && row < nv;
      v[i] = live ? ld4(y + row * channels + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
      rr[i] = (live && res != nullptr) ? ld4(res + row * channels + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t row = r + i * rstep;
      if (row >= row1) break;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);         // padding rows of a capacity-sized activation stay exactly zero
      if (row < nv) {
        o.x = fmaf((v[i].x - m.x) * rs.x, g.x, b.x);
        o.y = fmaf((v[i].y - m.y) * rs.y, g.y, b.y);
        o.z = fmaf((v[i].z - m.z) * rs.z, g.z, b.z);
        o.w = fmaf((v[i].w - m.w) * rs.w, g.w, b.w);
        if (res != nullptr) add4(o, rr[i]);
        if (relu) {
          o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
      }
      // ldz: row pitch of the outputs (> channels when z is the left part of a concatenation buffer)
      if (z != nullptr) *reinterpret_cast<float4*>(z + row * ldz + ch) = o;
      if (z16 != nullptr)
        *reinterpret_cast<uint2*>(z16 + row * ldz + ch) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// g' = gz * [z > 0]   (mask from the saved output: bf16 copy if present, else fp32; no mask when both are null)
__device__ __forceinline__ float4 masked_grad(const float* __restrict__ gz, const __nv_bfloat16* __restrict__ z16,
                                              const float* __restrict__ z, int64_t off) {
  float4 g = ld4(gz + off);
  if (z16 != nullptr) {
    const float4 o = bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(z16 + off)));
    g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
  } else if (z != nullptr) {
    const float4 o = ld4(z + off);
    g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
  }
  return g;
}

// red = [c1[C], c2[C]] = [sum g'/n, sum g'*xhat/n];  dgamma = sum g'*xhat;  dbeta = sum g'
template <int RB>
__global__ void __launch_bounds__(kColThreads)
bn_bwd_reduce_kernel(const float* __restrict__ gz, const float* __restrict__ y, const __nv_bfloat16* __restrict__ z16,
                     const float* __restrict__ z, int64_t n, int channels, int rows_per_cta,
                     const float* __restrict__ stat, float* __restrict__ partials, int64_t ldg,
                     const int32_t* __restrict__ valid_rows) {
  pdl_enter();
  __shared__ float4 s_stage[kColStageFloat4];
  n = effective_rows(n, valid_rows);
  const int ch = threadIdx.x * 4;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row1 = row0 + rows_per_cta < n ? row0 + rows_per_cta : n;
  const float4 m = ld4(stat + ch), rs = ld4(stat + channels + ch);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const int64_t rstep = blockDim.y;
  for (int64_t r = row0 + threadIdx.y; r < row1; r += RB * rstep) {      // 4 rows x (gz, mask, y) requested together
    float4 g[RB], v[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t rr = r + i * rstep < row1 ? r + i * rstep : r;     // tail lanes re-read row r and are discarded
      g[i] = masked_grad(gz, z16, z, rr * ldg + ch);                    // ldg: row pitch of gz and of the saved output
      v[i] = ld4(y + rr * channels + ch);
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      if (r + i * rstep >= row1) break;
      const float4 xh = make_float4((v[i].x - m.x) * rs.x, (v[i].y - m.y) * rs.y, (v[i].z - m.z) * rs.z,
                                    (v[i].w - m.w) * rs.w);
      add4(s1, g[i]);
      fma4(s2, g[i], xh);
    }
  }
  col_publish(s1, s2, partials, channels, s_stage);
}

// gy = gamma * rstd * (g' - c1 - xhat * c2)   (training);   gy = gamma * rstd * g'   (frozen statistics: red == null)
template <int RB>
__global__ void __launch_bounds__(kColThreads)
bn_bwd_apply_kernel(const float* __restrict__ gz, const float* __restrict__ y, const __nv_bfloat16* __restrict__ z16,
                    const float* __restrict__ z, int64_t n, int channels, int rows_per_cta,
                    const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ red,
                    float* __restrict__ gy, __nv_bfloat16* __restrict__ gy16, float* __restrict__ gres, int64_t ldg,
                    const int32_t* __restrict__ valid_rows) {
  pdl_enter();
  const int64_t nv = effective_rows(n, valid_rows);
  const int ch = threadIdx.x * 4;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row1 = row0 + rows_per_cta < n ? row0 + rows_per_cta : n;
  const float4 m = ld4(stat + ch), rs = ld4(stat + channels + ch), ga = ld4(gamma + ch);
  float4 c1 = make_float4(0.f, 0.f, 0.f, 0.f), c2 = c1;
  if (red != nullptr) {
    c1 = ld4(red + ch);
    c2 = ld4(red + channels + ch);
  }
  const int64_t rstep = blockDim.y;
  for (int64_t r = row0 + threadIdx.y; r < row1; r += RB * rstep) {
    float4 g[RB], v[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t row = r + i * rstep;
      const int64_t rr = (row < row1 && row < nv) ? row : (r < nv ? r : 0);   // dead lanes re-read a live row
      g[i] = masked_grad(gz, z16, z, rr * ldg + ch);
      v[i] = red != nullptr ? ld4(y + rr * channels + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t row = r + i * rstep;
      if (row >= row1) break;
      const int64_t off = row * channels + ch;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f), gg = o;     // padding rows carry no gradient
      if (row < nv) {
        gg = g[i];
        if (red != nullptr) {
          o.x = ga.x * rs.x * (gg.x - c1.x - (v[i].x - m.x) * rs.x * c2.x);
          o.y = ga.y * rs.y * (gg.y - c1.y - (v[i].y - m.y) * rs.y * c2.y);
          o.z = ga.z * rs.z * (gg.z - c1.z - (v[i].z - m.z) * rs.z * c2.z);
          o.w = ga.w * rs.w * (gg.w - c1.w - (v[i].w - m.w) * rs.w * c2.w);
        } else {
          o = make_float4(ga.x * rs.x * gg.x, ga.y * rs.y * gg.y, ga.z * rs.z * gg.z, ga.w * rs.w * gg.w);
        }
      }
      if (gy != nullptr) *reinterpret_cast<float4*>(gy + off) = o;
      if (gy16 != nullptr) *reinterpret_cast<uint2*>(gy16 + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      if (gres != nullptr) *reinterpret_cast<float4*>(gres + off) = gg;
    }
  }
}

// dst[row, col0 + c] = src[row, c] for the fp32 tensor and (optionally) its bf16 copy: the skip half of
// torchsparse.cat([deconv(y), skip]) (models/spvcnn.py:212-228) written next to the half bn_apply wrote in place
__global__ void copy_cols_kernel(const float* __restrict__ src, const __nv_bfloat16* __restrict__ src16, int64_t n, int c,
                                 float* __restrict__ dst, __nv_bfloat16* __restrict__ dst16, int64_t ld, int col0) {
  pdl_enter();
  const int c4 = c >> 2;
  const int64_t total = n * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / c4;
    const int q = (int)(i - row * c4) * 4;
    const float4 v = ld4(src + row * c + q);
    *reinterpret_cast<float4*>(dst + row * ld + col0 + q) = v;
    if (dst16 != nullptr) {
      uint2 h;
      if (src16 != nullptr) h = __ldg(reinterpret_cast<const uint2*>(src16 + row * c + q));
      else h = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      *reinterpret_cast<uint2*>(dst16 + row * ld + col0 + q) = h;
    }
  }
}

}  // namespace ft3d

using namespace ft3d;

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" {

size_t ft3d_bn_workspace(int32_t channels) { return col_workspace_bytes(channels > 0 ? channels : 4); }

int ft3d_bn_stats(const float* y, int64_t n, int32_t channels, float eps, float momentum, float* stat,
                  float* running_mean, float* running_var, const int32_t* valid_rows, void* workspace,
                  size_t workspace_bytes, ft3d_stream_t stream) {
  FT3D_REQUIRE(n > 0, "ft3d_bn_stats: BatchNorm statistics need at least one row");
  FT3D_REQUIRE(y && stat && workspace && channels >= 4 && channels % 4 == 0 && channels <= 1024,
               "ft3d_bn_stats: bad arguments (channels=%d)", channels);
  FT3D_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "ft3d_bn_stats: running stats go together");
  FT3D_REQUIRE(aligned16(y) && aligned16(workspace), "ft3d_bn_stats: pointers must be 16-byte aligned");
  FT3D_REQUIRE(workspace_bytes >= col_workspace_bytes(channels), "ft3d_bn_stats: workspace too small");
  ColGrid g = col_grid(n, channels / 4);
  switch (row_batch()) {
    case 1: launch_pdl(bn_stats_kernel<1>, dim3(g.grid), dim3(g.block), 0, (cudaStream_t)stream, y, n, channels, g.rows_per_cta, (float*)workspace,
                                                                valid_rows, (const float*)running_mean); break;
    case 2: launch_pdl(bn_stats_kernel<2>, dim3(g.grid), dim3(g.block), 0, (cudaStream_t)stream, y, n, channels, g.rows_per_cta, (float*)workspace,
                                                                valid_rows, (const float*)running_mean); break;
    default: launch_pdl(bn_stats_kernel<4>, dim3(g.grid), dim3(g.block), 0, (cudaStream_t)stream, y, n, channels, g.rows_per_cta, (float*)workspace,
                                                                valid_rows, (const float*)running_mean); break;
  }
  launch_pdl(col_finalize_kernel<0>, dim3(channels / 4), dim3(kColThreads), 0, (cudaStream_t)stream, (const float*)workspace, g.grid, channels, n, eps, momentum, stat, running_mean, running_var, 0, valid_rows);
  return check_launch("ft3d_bn_stats");
}

int ft3d_col_sum(const float* x, int64_t n, int32_t channels, float* out, int32_t accumulate, void* workspace,
                 size_t workspace_bytes, ft3d_stream_t stream) {
  FT3D_REQUIRE(n > 0 && x && out && workspace && channels >= 4 && channels % 4 == 0 && channels <= 1024,
               "ft3d_col_sum: bad arguments (channels=%d)", channels);
  FT3D_REQUIRE(aligned16(x) && aligned16(workspace), "ft3d_col_sum: pointers must be 16-byte aligned");
  FT3D_REQUIRE(workspace_bytes >= col_workspace_bytes(channels), "ft3d_col_sum: workspace too small");
  ColGrid g = col_grid(n, channels / 4);
  launch_pdl(bn_stats_kernel<2>, dim3(g.grid), g.block, 0, (cudaStream_t)stream, x, n, channels, g.rows_per_cta,
             (float*)workspace, (const int32_t*)nullptr, (const float*)nullptr);
  launch_pdl(col_finalize_kernel<2>, dim3(channels / 4), dim3(kColThreads), 0, (cudaStream_t)stream,
             (const float*)workspace, g.grid, channels, n, 0.f, 0.f, out, (float*)nullptr, (float*)nullptr, accumulate,
             (const int32_t*)nullptr);
  return check_launch("ft3d_col_sum");
}

int ft3d_bn_apply(const float* y, int64_t n, int32_t channels, const float* stat, const float* gamma, const float* beta,
                  const float* res, int32_t relu, float* z, void* z16, int64_t ldz, const int32_t* valid_rows,
                  ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  if (ldz == 0) ldz = channels;
  FT3D_REQUIRE(ldz >= channels && ldz % 4 == 0, "ft3d_bn_apply: output pitch must be a multiple of 4 and >= channels");
  FT3D_REQUIRE(y && stat && gamma && beta && (z || z16) && channels >= 4 && channels % 4 == 0,
               "ft3d_bn_apply: bad arguments");
  FT3D_REQUIRE(aligned16(y) && aligned16(stat) && aligned16(gamma) && aligned16(beta) && aligned16(res) &&
                   aligned16(z) && ((uintptr_t)z16 & 7) == 0,
               "ft3d_bn_apply: pointers must be 16-byte aligned");
  FT3D_REQUIRE(channels <= 1024, "ft3d_bn_apply: at most 1024 channels");
  ColGrid ga = col_grid(n, channels / 4);
  switch (row_batch()) {
    case 1: launch_pdl(bn_apply_kernel<1>, dim3(ga.grid), ga.block, 0, (cudaStream_t)stream, y, n, channels, ga.rows_per_cta, stat,
             gamma, beta, res, relu, z, (__nv_bfloat16*)z16, ldz, valid_rows); break;
    case 2: launch_pdl(bn_apply_kernel<2>, dim3(ga.grid), ga.block, 0, (cudaStream_t)stream, y, n, channels, ga.rows_per_cta, stat,
             gamma, beta, res, relu, z, (__nv_bfloat16*)z16, ldz, valid_rows); break;
    default: launch_pdl(bn_apply_kernel<4>, dim3(ga.grid), ga.block, 0, (cudaStream_t)stream, y, n, channels, ga.rows_per_cta, stat,
             gamma, beta, res, relu, z, (__nv_bfloat16*)z16, ldz, valid_rows); break;
  }
  return check_launch("ft3d_bn_apply");
}

int ft3d_copy_cols(const float* src, const void* src16, int64_t n, int32_t c, float* dst, void* dst16, int64_t ld,
                   int32_t col0, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(src && dst && c >= 4 && c % 4 == 0 && col0 % 4 == 0 && ld % 4 == 0 && ld >= col0 + c,
               "ft3d_copy_cols: bad arguments");
  FT3D_REQUIRE(aligned16(src) && aligned16(dst) && ((uintptr_t)src16 & 7) == 0 && ((uintptr_t)dst16 & 7) == 0,
               "ft3d_copy_cols: pointers must be 16-byte aligned");
  launch_pdl(copy_cols_kernel, dim3(grid_for(n * (c / 4), 256)), dim3(256), 0, (cudaStream_t)stream, src,
             (const __nv_bfloat16*)src16, n, (int)c, dst, (__nv_bfloat16*)dst16, ld, (int)col0);
  return check_launch("ft3d_copy_cols");
}

int ft3d_bn_bwd_reduce(const float* gz, const float* y, const void* z16, const float* z, int64_t n, int32_t channels,
                       const float* stat, float* red, float* dgamma, float* dbeta, int32_t accumulate, int64_t ldg,
                       const int32_t* valid_rows, void* workspace, size_t workspace_bytes, ft3d_stream_t stream) {
  FT3D_REQUIRE(n > 0, "ft3d_bn_bwd_reduce: needs at least one row");
  if (ldg == 0) ldg = channels;
  FT3D_REQUIRE(ldg >= channels && ldg % 4 == 0, "ft3d_bn_bwd_reduce: gz pitch must be a multiple of 4 and >= channels");
  FT3D_REQUIRE(gz && y && stat && red && dgamma && dbeta && workspace && channels >= 4 &&
                   channels % 4 == 0 && channels <= 1024,
               "ft3d_bn_bwd_reduce: bad arguments");
  FT3D_REQUIRE(aligned16(gz) && aligned16(y) && aligned16(z) && ((uintptr_t)z16 & 7) == 0 && aligned16(stat) &&
                   aligned16(workspace),
               "ft3d_bn_bwd_reduce: pointers must be 16-byte aligned");
  FT3D_REQUIRE(workspace_bytes >= col_workspace_bytes(channels), "ft3d_bn_bwd_reduce: workspace too small");
  ColGrid g = col_grid(n, channels / 4);
  switch (row_batch()) {
    case 1: launch_pdl(bn_bwd_reduce_kernel<1>, dim3(g.grid), dim3(g.block), 0, (cudaStream_t)stream, gz, y, (const __nv_bfloat16*)z16, z, n, channels, g.rows_per_cta, stat, (float*)workspace, ldg, valid_rows); break;
    case 2: launch_pdl(bn_bwd_reduce_kernel<2>, dim3(g.grid), dim3(g.block), 0, (cudaStream_t)stream, gz, y, (const __nv_bfloat16*)z16, z, n, channels, g.rows_per_cta, stat, (float*)workspace, ldg, valid_rows); break;
    default: launch_pdl(bn_bwd_reduce_kernel<4>, dim3(g.grid), dim3(g.block), 0, (cudaStream_t)stream, gz, y, (const __nv_bfloat16*)z16, z, n, channels, g.rows_per_cta, stat, (float*)workspace, ldg, valid_rows); break;
  }
  launch_pdl(col_finalize_kernel<1>, dim3(channels / 4), dim3(kColThreads), 0, (cudaStream_t)stream, (const float*)workspace, g.grid, channels, n, 0.f, 0.f, red, dgamma, dbeta, accumulate, valid_rows);
  return check_launch("ft3d_bn_bwd_reduce");
}

int ft3d_bn_bwd_apply(const float* gz, const float* y, const void* z16, const float* z, int64_t n, int32_t channels,
                      const float* stat, const float* gamma, const float* red, float* gy, void* gy16, float* gres,
                      int64_t ldg, const int32_t* valid_rows, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  if (ldg == 0) ldg = channels;
  FT3D_REQUIRE(ldg >= channels && ldg % 4 == 0, "ft3d_bn_bwd_apply: gz pitch must be a multiple of 4 and >= channels");
  FT3D_REQUIRE(gz && stat && gamma && (y || !red) && (gy || gy16 || gres) && channels >= 4 && channels % 4 == 0,
               "ft3d_bn_bwd_apply: bad arguments");
  FT3D_REQUIRE(aligned16(gz) && aligned16(y) && aligned16(z) && ((uintptr_t)z16 & 7) == 0 && aligned16(stat) &&
                   aligned16(gamma) && aligned16(red) && aligned16(gy) && ((uintptr_t)gy16 & 7) == 0 && aligned16(gres),
               "ft3d_bn_bwd_apply: pointers must be 16-byte aligned");
  FT3D_REQUIRE(channels <= 1024, "ft3d_bn_bwd_apply: at most 1024 channels");
  ColGrid gb = col_grid(n, channels / 4);
  switch (row_batch()) {
    case 1: launch_pdl(bn_bwd_apply_kernel<1>, dim3(gb.grid), gb.block, 0, (cudaStream_t)stream, gz, y, (const __nv_bfloat16*)z16, z,
             n, channels, gb.rows_per_cta, stat, gamma, red, gy, (__nv_bfloat16*)gy16, gres, ldg, valid_rows); break;
    case 2: launch_pdl(bn_bwd_apply_kernel<2>, dim3(gb.grid), gb.block, 0, (cudaStream_t)stream, gz, y, (const __nv_bfloat16*)z16, z,
             n, channels, gb.rows_per_cta, stat, gamma, red, gy, (__nv_bfloat16*)gy16, gres, ldg, valid_rows); break;
    default: launch_pdl(bn_bwd_apply_kernel<4>, dim3(gb.grid), gb.block, 0, (cudaStream_t)stream, gz, y, (const __nv_bfloat16*)z16, z,
             n, channels, gb.rows_per_cta, stat, gamma, red, gy, (__nv_bfloat16*)gy16, gres, ldg, valid_rows); break;
  }
  return check_launch("ft3d_bn_bwd_apply");
}

}  // extern "C"
