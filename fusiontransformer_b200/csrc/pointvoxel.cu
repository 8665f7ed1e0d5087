// Point<->voxel gathers (K5-K8), fused map builders for voxel_to_point / point_to_voxel, and the
// 2D->3D lift (K16).  All HBM/L2-bound: 16-byte vectorised row accesses, one thread per 4 channels,
// consecutive threads on consecutive channels of one row so every warp access is a contiguous segment.
#include "common.cuh"

namespace ft3d {

__global__ void zero_f32_kernel(float* __restrict__ p, int64_t n) {
  pdl_enter();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n4 = n >> 2;
  float4* p4 = (float4*)p;
  for (int64_t j = i; j < n4; j += stride) p4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t j = (n4 << 2) + i; j < n; j += stride) p[j] = 0.f;
}

__global__ void zero_i32_kernel2(int32_t* __restrict__ p, int64_t n) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0;
}

// ------------------------------------------------------------------ K5
__global__ void count_kernel(const int32_t* __restrict__ idx, int64_t n, int32_t* __restrict__ cnt, int64_t m) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int v = idx[i];
    if (v >= 0 && v < m) atomicAdd(cnt + v, 1);
  }
}

// ------------------------------------------------------------------ K6
template <int VEC>
__global__ void voxelize_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                                    const int32_t* __restrict__ cnt, int64_t n, int64_t m, int c,
                                    float* __restrict__ out) {
  pdl_enter();
  const int cv = c / VEC;
  int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / cv;
    int ch = (int)(t - i * cv) * VEC;
    int v = __ldg(idx + i);
    if (v < 0 || v >= m) continue;
    float inv_den = (float)__ldg(cnt + v);
    if (VEC == 4) {
      float4 f = __ldg((const float4*)(feat + i * c + ch));
      f.x = __fdiv_rn(f.x, inv_den); f.y = __fdiv_rn(f.y, inv_den);
      f.z = __fdiv_rn(f.z, inv_den); f.w = __fdiv_rn(f.w, inv_den);
      atomicAdd((float4*)(out + (int64_t)v * c + ch), f);
    } else {
      atomicAdd(out + (int64_t)v * c + ch, __fdiv_rn(__ldg(feat + i * c + ch), inv_den));
    }
  }
}

template <int VEC>
__global__ void voxelize_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx,
                                    const int32_t* __restrict__ cnt, int64_t n, int64_t m, int c,
                                    float* __restrict__ gin) {
  pdl_enter();
  const int cv = c / VEC;
  int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / cv;
    int ch = (int)(t - i * cv) * VEC;
    int v = __ldg(idx + i);
    bool ok = (v >= 0 && v < m);
    float den = ok ? (float)__ldg(cnt + v) : 1.f;
    if (VEC == 4) {
      float4 g = ok ? __ldg((const float4*)(gout + (int64_t)v * c + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
      g.x = __fdiv_rn(g.x, den); g.y = __fdiv_rn(g.y, den); g.z = __fdiv_rn(g.z, den); g.w = __fdiv_rn(g.w, den);
      *(float4*)(gin + i * c + ch) = g;
    } else {
      gin[i * c + ch] = ok ? __fdiv_rn(__ldg(gout + (int64_t)v * c + ch), den) : 0.f;
    }
  }
}

// ------------------------------------------------------------------ K7
template <int VEC>
__global__ void devoxelize_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                                      const float* __restrict__ w, int64_t n, int64_t m, int c,
                                      float* __restrict__ out) {
  pdl_enter();
  const int cv = c / VEC;
  int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / cv;
    int ch = (int)(t - i * cv) * VEC;
    float acc[VEC];
    #pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    #pragma unroll
    for (int k = 0; k < 8; ++k) {
      int v = __ldg(idx + i * 8 + k);
      if (v < 0 || v >= m) continue;
      float wk = __ldg(w + i * 8 + k);
      if (VEC == 4) {
        float4 f = __ldg((const float4*)(feat + (int64_t)v * c + ch));
        acc[0] = fmaf(wk, f.x, acc[0]); acc[1] = fmaf(wk, f.y, acc[1]);
        acc[2] = fmaf(wk, f.z, acc[2]); acc[3] = fmaf(wk, f.w, acc[3]);
      } else {
        acc[0] = fmaf(wk, __ldg(feat + (int64_t)v * c + ch), acc[0]);
      }
    }
    if (VEC == 4) *(float4*)(out + i * c + ch) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else out[i * c + ch] = acc[0];
  }
}

template <int VEC>
__global__ void devoxelize_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx,
                                      const float* __restrict__ w, int64_t n, int64_t m, int c,
                                      float* __restrict__ gfeat) {
  pdl_enter();
  const int cv = c / VEC;
  int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / cv;
    int ch = (int)(t - i * cv) * VEC;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float g1 = 0.f;
    if (VEC == 4) g4 = __ldg((const float4*)(gout + i * c + ch));
    else g1 = __ldg(gout + i * c + ch);
    #pragma unroll
    for (int k = 0; k < 8; ++k) {
      int v = __ldg(idx + i * 8 + k);
      if (v < 0 || v >= m) continue;
      float wk = __ldg(w + i * 8 + k);
      if (wk == 0.f) continue;
      if (VEC == 4) atomicAdd((float4*)(gfeat + (int64_t)v * c + ch), make_float4(wk * g4.x, wk * g4.y, wk * g4.z, wk * g4.w));
      else atomicAdd(gfeat + (int64_t)v * c + ch, wk * g1);
    }
  }
}

// ------------------------------------------------------------------ deterministic segmented sums (FT3D_DETERMINISTIC)
// out[v,:] = sum over entries e in [off[v], off[v+1]), IN ENTRY ORDER, of  w[e] * src[row[e],:] (/ cnt[v] per term)
// The scatter kernels above add their terms with fp32 atomics in arrival order (as the reference's own torchsparse
// kernels do); here the caller has sorted the contributions by destination row (stable, so ascending source row)
// and a thread group owns a destination row: a fixed summation order, bit-identical from run to run.
template <int VEC>
__global__ void segsum_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ entry_row,
                                   const float* __restrict__ entry_w, const int32_t* __restrict__ off,
                                   const int32_t* __restrict__ cnt, int64_t m, int c, float* __restrict__ out) {
  pdl_enter();
  const int cv = c / VEC;
  const int64_t total = m * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t / cv;
    const int ch = (int)(t - v * cv) * VEC;
    const int e0 = __ldg(off + v), e1 = __ldg(off + v + 1);
    const float den = cnt != nullptr ? (float)max(__ldg(cnt + v), 1) : 1.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
      const int64_t r = __ldg(entry_row + e);
      const float wk = entry_w != nullptr ? __ldg(entry_w + e) : 1.f;
      if (VEC == 4) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src + r * c + ch));
        if (cnt != nullptr) {
          acc.x += __fdiv_rn(x.x, den); acc.y += __fdiv_rn(x.y, den); acc.z += __fdiv_rn(x.z, den); acc.w += __fdiv_rn(x.w, den);
        } else {
          acc.x += wk * x.x; acc.y += wk * x.y; acc.z += wk * x.z; acc.w += wk * x.w;
        }
      } else {
        const float x = __ldg(src + r * c + ch);
        acc.x += cnt != nullptr ? __fdiv_rn(x, den) : wk * x;
      }
    }
    if (VEC == 4) *reinterpret_cast<float4*>(out + v * c + ch) = acc;
    else out[v * c + ch] = acc.x;
  }
}

// ------------------------------------------------------------------ K8 (API layout [8,n], idx int64)
__device__ __forceinline__ float ti_weight(float x, float y, float z, float scale, int k) {
  float flx = (scale != 1.f) ? floorf(x / scale) * scale : floorf(x);
  float fly = (scale != 1.f) ? floorf(y / scale) * scale : floorf(y);
  float flz = (scale != 1.f) ? floorf(z / scale) * scale : floorf(z);
  float fx = (k & 4) ? __fsub_rn(x, flx) : __fsub_rn(__fadd_rn(flx, scale), x);
  float fy = (k & 2) ? __fsub_rn(y, fly) : __fsub_rn(__fadd_rn(fly, scale), y);
  float fz = (k & 1) ? __fsub_rn(z, flz) : __fsub_rn(__fadd_rn(flz, scale), z);
  float wk = __fmul_rn(__fmul_rn(fx, fy), fz);
  if (scale != 1.f) wk = __fdiv_rn(wk, scale * scale * scale);
  return wk;
}

__global__ void ti_weights_kernel(const float4* __restrict__ pc, const int64_t* __restrict__ idx, int64_t n,
                                  float scale, float* __restrict__ w_out) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = __ldg(pc + i);
    float wk[8];
    float sum = 0.f;
    #pragma unroll
    for (int k = 0; k < 8; ++k) {
      wk[k] = (idx[(int64_t)k * n + i] == -1) ? 0.f : ti_weight(p.x, p.y, p.z, scale, k);
      sum += wk[k];
    }
    float den = sum + 1e-8f;
    #pragma unroll
    for (int k = 0; k < 8; ++k) w_out[(int64_t)k * n + i] = __fdiv_rn(wk[k], den);
  }
}

// ------------------------------------------------------------------ fused voxel_to_point map build
// 8 lanes per point: lane k probes corner k (dx=k>>2, dy=(k>>1)&1, dz=k&1, KernelRegion(2) order).
__global__ void v2p_build_kernel(const float4* __restrict__ pc, int64_t n, int stride,
                                 const unsigned long long* __restrict__ tkeys, const int* __restrict__ tvals,
                                 uint32_t mask, int32_t* __restrict__ idx_out, float* __restrict__ w_out) {
  pdl_enter();
  const float scale = (float)stride;
  int64_t total = n * 8;
  int64_t padded = (total + 31) / 32 * 32;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < padded; t += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(t & 7);
    int64_t i = t >> 3;
    bool live = i < n;
    float wk = 0.f;
    int v = -1;
    if (live) {
      float4 p = __ldg(pc + i);
      int bx = (int)floorf(p.x / scale) * stride;
      int by = (int)floorf(p.y / scale) * stride;
      int bz = (int)floorf(p.z / scale) * stride;
      int b = (int)p.w;
      v = table_lookup(tkeys, tvals, mask,
                       fnv1a_fold(bx + ((k >> 2) & 1) * stride, by + ((k >> 1) & 1) * stride, bz + (k & 1) * stride, b));
      wk = (v < 0) ? 0.f : ti_weight(p.x, p.y, p.z, scale, k);
    }
    float sum = wk;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    if (live) {
      idx_out[t] = v;
      w_out[t] = __fdiv_rn(wk, sum + 1e-8f);
    }
  }
}

__global__ void p2v_build_kernel(const float4* __restrict__ pc, int64_t n, int stride,
                                 const unsigned long long* __restrict__ tkeys, const int* __restrict__ tvals,
                                 uint32_t mask, int32_t* __restrict__ idx_out, int32_t* __restrict__ cnt, int64_t m) {
  pdl_enter();
  const float scale = (float)stride;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = __ldg(pc + i);
    int bx = (int)floorf(p.x / scale) * stride;
    int by = (int)floorf(p.y / scale) * stride;
    int bz = (int)floorf(p.z / scale) * stride;
    int v = table_lookup(tkeys, tvals, mask, fnv1a_fold(bx, by, bz, (int)p.w));
    idx_out[i] = v;
    if (v >= 0 && v < m) atomicAdd(cnt + v, 1);
  }
}

// ------------------------------------------------------------------ K16 lift
template <int VEC>
__global__ void lift_fwd_kernel(const float* __restrict__ fmap, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                                int B, int C, int H, int W, const int2* __restrict__ rc,
                                const int32_t* __restrict__ bidx, int64_t n, float* __restrict__ out) {
  pdl_enter();
  const int cv = C / VEC;
  int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = t / cv;
    int ch = (int)(t - p * cv) * VEC;
    int2 q = __ldg(rc + p);
    int b = __ldg(bidx + p);
    const bool ok = q.x >= 0 && q.x < H && q.y >= 0 && q.y < W && b >= 0 && b < B;   // out of image -> zeros
    const float* src = fmap + (int64_t)b * sb + (int64_t)q.x * sh + (int64_t)q.y * sw + (int64_t)ch * sc;
    if (VEC == 4) *(float4*)(out + p * C + ch) = ok ? __ldg((const float4*)src) : make_float4(0.f, 0.f, 0.f, 0.f);
    else out[p * C + ch] = ok ? __ldg(src) : 0.f;
  }
}

template <int VEC>
__global__ void lift_bwd_kernel(const float* __restrict__ gout, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                                int B, int C, int H, int W, const int2* __restrict__ rc,
                                const int32_t* __restrict__ bidx, int64_t n, float* __restrict__ gmap) {
  pdl_enter();
  const int cv = C / VEC;
  int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = t / cv;
    int ch = (int)(t - p * cv) * VEC;
    int2 q = __ldg(rc + p);
    int b = __ldg(bidx + p);
    if (!(q.x >= 0 && q.x < H && q.y >= 0 && q.y < W && b >= 0 && b < B)) continue;
    float* dst = gmap + (int64_t)b * sb + (int64_t)q.x * sh + (int64_t)q.y * sw + (int64_t)ch * sc;
    if (VEC == 4) atomicAdd((float4*)dst, __ldg((const float4*)(gout + p * C + ch)));
    else atomicAdd(dst, __ldg(gout + p * C + ch));
  }
}

}  // namespace ft3d

using namespace ft3d;

#define LAUNCH_VEC(kernel, c, n, s, ...)                                                     \
  do {                                                                                       \
    if ((c) % 4 == 0) launch_pdl(kernel<4>, dim3(grid_for((n) * ((c) / 4), 256)), dim3(256), 0, s, __VA_ARGS__); \
    else launch_pdl(kernel<1>, dim3(grid_for((n) * (c), 256)), dim3(256), 0, s, __VA_ARGS__);                    \
  } while (0)

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" {

int ft3d_count(const int32_t* idx, int64_t n, int32_t* cnt_out, int64_t m, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (m > 0) {
    FT3D_REQUIRE(cnt_out != nullptr, "ft3d_count: null output");
    launch_pdl(zero_i32_kernel2, dim3(grid_for(m, 256)), dim3(256), 0, s, cnt_out, m);
  }
  if (n > 0 && m > 0) {
    FT3D_REQUIRE(idx != nullptr, "ft3d_count: null idx");
    launch_pdl(count_kernel, dim3(grid_for(n, 256)), dim3(256), 0, s, idx, n, cnt_out, m);
  }
  return check_launch("ft3d_count");
}

int ft3d_voxelize_fwd(const float* feat, const int32_t* idx, const int32_t* cnt, int64_t n, int64_t m,
                      int32_t c, float* out, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(c > 0, "ft3d_voxelize_fwd: c must be positive");
  if (m > 0) {
    FT3D_REQUIRE(out != nullptr, "ft3d_voxelize_fwd: null output");
    launch_pdl(zero_f32_kernel, dim3(grid_for(m * c / 4 + 1, 256)), dim3(256), 0, s, out, m * c);
  }
  if (n > 0 && m > 0) {
    FT3D_REQUIRE(feat && idx && cnt, "ft3d_voxelize_fwd: null input");
    if (c % 4 == 0 && aligned16(feat) && aligned16(out)) launch_pdl(voxelize_fwd_kernel<4>, dim3(grid_for(n * (c / 4), 256)), dim3(256), 0, s, feat, idx, cnt, n, m, c, out);
    else launch_pdl(voxelize_fwd_kernel<1>, dim3(grid_for(n * c, 256)), dim3(256), 0, s, feat, idx, cnt, n, m, c, out);
  }
  return check_launch("ft3d_voxelize_fwd");
}

int ft3d_voxelize_bwd(const float* gout, const int32_t* idx, const int32_t* cnt, int64_t n, int64_t m,
                      int32_t c, float* gin, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(gout && idx && cnt && gin && c > 0, "ft3d_voxelize_bwd: bad arguments");
  if (c % 4 == 0 && aligned16(gout) && aligned16(gin))
    launch_pdl(voxelize_bwd_kernel<4>, dim3(grid_for(n * (c / 4), 256)), dim3(256), 0, s, gout, idx, cnt, n, m, c, gin);
  else
    launch_pdl(voxelize_bwd_kernel<1>, dim3(grid_for(n * c, 256)), dim3(256), 0, s, gout, idx, cnt, n, m, c, gin);
  return check_launch("ft3d_voxelize_bwd");
}

int ft3d_devoxelize_fwd(const float* feat, const int32_t* idx, const float* w, int64_t n, int64_t m,
                        int32_t c, float* out, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(idx && w && out && c > 0 && (feat || m == 0), "ft3d_devoxelize_fwd: bad arguments");
  if (c % 4 == 0 && aligned16(feat) && aligned16(out))
    launch_pdl(devoxelize_fwd_kernel<4>, dim3(grid_for(n * (c / 4), 256)), dim3(256), 0, s, feat, idx, w, n, m, c, out);
  else
    launch_pdl(devoxelize_fwd_kernel<1>, dim3(grid_for(n * c, 256)), dim3(256), 0, s, feat, idx, w, n, m, c, out);
  return check_launch("ft3d_devoxelize_fwd");
}

int ft3d_segsum_rows(const float* src, const int32_t* entry_row, const float* entry_w, const int32_t* offsets,
                     const int32_t* cnt, int64_t m, int32_t c, float* out, ft3d_stream_t stream) {
  if (m == 0) return FT3D_OK;
  FT3D_REQUIRE(src && entry_row && offsets && out && c > 0 && !(entry_w && cnt), "ft3d_segsum_rows: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (c % 4 == 0 && aligned16(src) && aligned16(out))
    launch_pdl(segsum_rows_kernel<4>, dim3(grid_for(m * (c / 4), 256)), dim3(256), 0, s, src, entry_row, entry_w, offsets, cnt, m, (int)c, out);
  else
    launch_pdl(segsum_rows_kernel<1>, dim3(grid_for(m * c, 256)), dim3(256), 0, s, src, entry_row, entry_w, offsets, cnt, m, (int)c, out);
  return check_launch("ft3d_segsum_rows");
}

int ft3d_devoxelize_bwd(const float* gout, const int32_t* idx, const float* w, int64_t n, int64_t m,
                        int32_t c, float* gfeat, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(c > 0, "ft3d_devoxelize_bwd: c must be positive");
  if (m > 0) {
    FT3D_REQUIRE(gfeat != nullptr, "ft3d_devoxelize_bwd: null output");
    launch_pdl(zero_f32_kernel, dim3(grid_for(m * c / 4 + 1, 256)), dim3(256), 0, s, gfeat, m * c);
  }
  if (n > 0 && m > 0) {
    FT3D_REQUIRE(gout && idx && w, "ft3d_devoxelize_bwd: null input");
    if (c % 4 == 0 && aligned16(gout) && aligned16(gfeat))
      launch_pdl(devoxelize_bwd_kernel<4>, dim3(grid_for(n * (c / 4), 256)), dim3(256), 0, s, gout, idx, w, n, m, c, gfeat);
    else
      launch_pdl(devoxelize_bwd_kernel<1>, dim3(grid_for(n * c, 256)), dim3(256), 0, s, gout, idx, w, n, m, c, gfeat);
  }
  return check_launch("ft3d_devoxelize_bwd");
}

int ft3d_ti_weights(const float* pc, const int64_t* idx, int64_t n, float scale, float* w_out,
                    ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(pc && idx && w_out && scale > 0.f && aligned16(pc), "ft3d_ti_weights: bad arguments");
  launch_pdl(ti_weights_kernel, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)pc, idx, n, scale, w_out);
  return check_launch("ft3d_ti_weights");
}

int ft3d_v2p_build(const float* pc, int64_t n, int32_t stride, const uint64_t* table_keys,
                   const int32_t* table_vals, int64_t cap, int32_t* idx_out, float* w_out,
                   ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(pc && table_keys && table_vals && idx_out && w_out && stride > 0 && aligned16(pc) &&
               cap >= 2 && (cap & (cap - 1)) == 0, "ft3d_v2p_build: bad arguments");
  launch_pdl(v2p_build_kernel, dim3(grid_for(n * 8, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)pc, n, stride, (const unsigned long long*)table_keys, table_vals, (uint32_t)(cap - 1),
      idx_out, w_out);
  return check_launch("ft3d_v2p_build");
}

int ft3d_p2v_build(const float* pc, int64_t n, int32_t stride, const uint64_t* table_keys,
                   const int32_t* table_vals, int64_t cap, int32_t* idx_out, int32_t* cnt_out, int64_t m,
                   ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (m > 0) {
    FT3D_REQUIRE(cnt_out != nullptr, "ft3d_p2v_build: null counts");
    launch_pdl(zero_i32_kernel2, dim3(grid_for(m, 256)), dim3(256), 0, s, cnt_out, m);
  }
  if (n > 0) {
    FT3D_REQUIRE(pc && table_keys && table_vals && idx_out && stride > 0 && aligned16(pc) && cap >= 2 &&
                 (cap & (cap - 1)) == 0, "ft3d_p2v_build: bad arguments");
    launch_pdl(p2v_build_kernel, dim3(grid_for(n, 256)), dim3(256), 0, s, (const float4*)pc, n, stride,
                                                      (const unsigned long long*)table_keys, table_vals,
                                                      (uint32_t)(cap - 1), idx_out, cnt_out, m);
  }
  return check_launch("ft3d_p2v_build");
}

int ft3d_lift_fwd(const float* fmap, int64_t sb, int64_t sc, int64_t sh, int64_t sw, int32_t B, int32_t C,
                  int32_t H, int32_t W, const int32_t* rc, const int32_t* bidx, int64_t n, float* out,
                  ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(fmap && rc && bidx && out && B > 0 && C > 0 && H > 0 && W > 0, "ft3d_lift_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (sc == 1) && (C % 4 == 0) && aligned16(fmap) && aligned16(out) && (sb % 4 == 0) && (sh % 4 == 0) &&
             (sw % 4 == 0);
  if (vec) launch_pdl(lift_fwd_kernel<4>, dim3(grid_for(n * (C / 4), 256)), dim3(256), 0, s, fmap, sb, sc, sh, sw, B, C, H, W,
                                                                        (const int2*)rc, bidx, n, out);
  else launch_pdl(lift_fwd_kernel<1>, dim3(grid_for(n * C, 256)), dim3(256), 0, s, fmap, sb, sc, sh, sw, B, C, H, W, (const int2*)rc,
                                                              bidx, n, out);
  return check_launch("ft3d_lift_fwd");
}

int ft3d_lift_bwd(const float* gout, int64_t sb, int64_t sc, int64_t sh, int64_t sw, int32_t B, int32_t C,
                  int32_t H, int32_t W, const int32_t* rc, const int32_t* bidx, int64_t n, float* gmap,
                  ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(gout && rc && bidx && gmap && B > 0 && C > 0 && H > 0 && W > 0, "ft3d_lift_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (sc == 1) && (C % 4 == 0) && aligned16(gmap) && aligned16(gout) && (sb % 4 == 0) && (sh % 4 == 0) &&
             (sw % 4 == 0);
  if (vec) launch_pdl(lift_bwd_kernel<4>, dim3(grid_for(n * (C / 4), 256)), dim3(256), 0, s, gout, sb, sc, sh, sw, B, C, H, W,
                                                                        (const int2*)rc, bidx, n, gmap);
  else launch_pdl(lift_bwd_kernel<1>, dim3(grid_for(n * C, 256)), dim3(256), 0, s, gout, sb, sc, sh, sw, B, C, H, W, (const int2*)rc,
                                                              bidx, n, gmap);
  return check_launch("ft3d_lift_bwd");
}

}  // extern "C"
