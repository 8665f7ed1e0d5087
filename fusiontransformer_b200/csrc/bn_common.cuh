// Per-channel column reductions shared by the BatchNorm kernels and the convolution's sorted scatter.
//
// Every kernel that reduces over rows uses a 2-D thread block: threadIdx.x = one group of 4 channels (a float4 of a
// row), threadIdx.y = row lane.  A thread keeps its channels for its whole life, accumulates its rows in registers,
// the block folds the row lanes in shared memory and writes ONE partial row per CTA to `partials`; a second, tiny
// launch folds the partial rows in double precision with a fixed-shape tree.  No float atomics: the result is
// independent of scheduling (deterministic), as the reference's cuDNN/ATen batch-norm reductions are.
#pragma once
#include "common.cuh"

namespace ft3d {

constexpr int kColMaxCtas = kNumSMs * 8;        // upper bound on CTAs of a column-reduction launch
constexpr int kColThreads = 256;                // target threads per CTA
constexpr int kColStageFloat4 = kColThreads * 2; // __shared__ float4 staging every column-reduction kernel declares

struct ColGrid {
  dim3 block;
  int grid;
  int rows_per_cta;
};

// cv = channels / 4.  Rows are split into contiguous per-CTA chunks (multiples of the row-lane count).
static inline ColGrid col_grid(int64_t n_rows, int cv) {
  ColGrid g;
  int ry = kColThreads / cv;
  if (ry < 1) ry = 1;
  if (ry > 32) ry = 32;
  g.block = dim3((unsigned)cv, (unsigned)ry, 1);
  int64_t want = (n_rows + ry - 1) / ry;                 // CTAs if each did one row per lane
  const int64_t cap = col_cta_cap();
  int64_t ctas = want < cap ? want : cap;
  if (ctas < 1) ctas = 1;
  int64_t rpc = (n_rows + ctas - 1) / ctas;
  rpc = (rpc + ry - 1) / ry * ry;
  if (rpc < ry) rpc = ry;
  g.rows_per_cta = (int)rpc;
  g.grid = (int)((n_rows + rpc - 1) / rpc);
  if (g.grid < 1) g.grid = 1;
  return g;
}

static inline size_t col_workspace_bytes(int channels) {
  return align_up((size_t)kColMaxCtas * 2 * channels * sizeof(float), 256);
}

// Row count of a capacity-padded activation: the launch covers `n` rows, of which the first *valid_rows are real
// (CUDA-graph replay with static shapes, graph.py).  valid_rows == nullptr: all n rows are real.
__device__ __forceinline__ int64_t effective_rows(int64_t n, const int32_t* __restrict__ valid_rows) {
  if (valid_rows == nullptr) return n;
  const int64_t v = (int64_t)__ldg(valid_rows);
  return v < n ? v : n;
}

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void fma4(float4& a, const float4& b, const float4& c) {
  a.x = fmaf(b.x, c.x, a.x); a.y = fmaf(b.y, c.y, a.y); a.z = fmaf(b.z, c.z, a.z); a.w = fmaf(b.w, c.w, a.w);
}

// BatchNorm statistics are accumulated as SHIFTED sums  s1 = sum (y - p), s2 = sum (y - p)^2  with the per-channel pivot
// p = the layer's running mean as it stands before this batch (0 without running statistics):  mean = p + s1/n,
// var = s2/n - (s1/n)^2.  E[y^2] - mean^2 on raw fp32 sums loses everything once |mean| >> std (mean 100, std 0.1:
// the variance drowns in the rounding of sum y^2); shifted by a pivot within a few std of the mean the subtraction is
// harmless.  Every producer of statistics partial rows (conv_os epilogue and fold, conv_reduce, bn_stats) and the
// finalize use the same pivot pointer; the finalize reads it before it updates the running mean.
__device__ __forceinline__ float4 stat_pivot(const float* __restrict__ pivot, int ch) {
  return pivot != nullptr ? __ldg(reinterpret_cast<const float4*>(pivot + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void stat_add(float4& s1, float4& s2, const float4& v, const float4& pv) {
  const float4 d = make_float4(v.x - pv.x, v.y - pv.y, v.z - pv.z, v.w - pv.w);
  add4(s1, d);
  fma4(s2, d, d);
}

// Fold (s1, s2) over the row lanes of the CTA and publish the CTA's partial row: partials[cta][0][C], [cta][1][C].
// `s_stage` must hold 2 * blockDim.x * blockDim.y float4.
__device__ __forceinline__ void col_publish(float4 s1, float4 s2, float* __restrict__ partials, int channels,
                                            float4* s_stage) {
  const int cv = blockDim.x, ry = blockDim.y;
  const int tx = threadIdx.x, ty = threadIdx.y;
  s_stage[(ty * cv + tx) * 2] = s1;
  s_stage[(ty * cv + tx) * 2 + 1] = s2;
  __syncthreads();
  if (ty == 0) {
    for (int r = 1; r < ry; ++r) {
      add4(s1, s_stage[(r * cv + tx) * 2]);
      add4(s2, s_stage[(r * cv + tx) * 2 + 1]);
    }
    float* p = partials + (size_t)blockIdx.x * 2 * channels;
    *reinterpret_cast<float4*>(p + tx * 4) = s1;
    *reinterpret_cast<float4*>(p + channels + tx * 4) = s2;
  }
}

// Second stage (its own tiny launch, stream-ordered after the main kernel): CTA g owns channels 4g..4g+3, its 256
// threads read the partial rows in parallel (one L2 round trip instead of a serial walk by a "last CTA"), accumulate
// in double and fold with a fixed-shape tree -> the summation order depends only on (nparts, blockDim).
//   MODE 0: BatchNorm training statistics (torch.nn.BatchNorm1d semantics: biased variance for the normalisation,
//           unbiased for running_var); out0 = stat [2,C] = (mean, rstd).
//   MODE 1: BatchNorm backward sums; out0 = red [2,C] = (sum g/n, sum g*xhat/n), out1 = dgamma, out2 = dbeta
//           (`accumulate`: added to what out1/out2 hold -- gradients written straight into the optimizer's arena).
//   MODE 2: plain column sums; out0 [C] = sum over rows (+ out0 when `accumulate`).
template <int MODE>
__global__ void __launch_bounds__(kColThreads)
col_finalize_kernel(const float* __restrict__ partials, int nparts, int channels, int64_t n, float eps, float momentum,
                    float* __restrict__ out0, float* __restrict__ out1, float* __restrict__ out2, int accumulate,
                    const int32_t* __restrict__ valid_rows) {
  pdl_enter();
  __shared__ double s_acc[kColThreads / 32][8];
  n = effective_rows(n, valid_rows);
  if (n < 1) n = 1;
  const int t = threadIdx.x, c0 = blockIdx.x * 4;
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // up to 4 partial rows per thread are requested before the first add: one L2 round trip for <= 1024 CTAs
  for (int b0 = t; b0 < nparts; b0 += 4 * kColThreads) {
    float4 u[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + i * kColThreads;
      u[i] = v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < nparts) {
        const float* p = partials + (size_t)b * 2 * channels + c0;
        u[i] = __ldg(reinterpret_cast<const float4*>(p));
        v[i] = __ldg(reinterpret_cast<const float4*>(p + channels));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[0] += u[i].x; a[1] += u[i].y; a[2] += u[i].z; a[3] += u[i].w;
      a[4] += v[i].x; a[5] += v[i].y; a[6] += v[i].z; a[7] += v[i].w;
    }
  }
  // fixed-shape fold: xor-butterfly inside each warp (no barrier), then the 8 warp totals through shared memory
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a[e] += __shfl_xor_sync(0xffffffffu, a[e], o);
  }
  if ((t & 31) == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[t >> 5][e] = a[e];
  }
  __syncthreads();
  if (t < 8) {
    double tot = 0.0;
    for (int wdx = 0; wdx < kColThreads / 32; ++wdx) tot += s_acc[wdx][t];
    s_acc[0][t] = tot;        // each of the 8 threads overwrites only its own column of row 0 after reading it
  }
  __syncthreads();
  if (t < 4) {
    const int c = c0 + t;
    const double s1 = s_acc[0][t], s2 = s_acc[0][4 + t];
    if (MODE == 0) {
      // shifted sums (stat_add): pivot = the running mean before this batch
      const double dm = s1 / (double)n;
      const double mean = (out1 != nullptr ? (double)out1[c] : 0.0) + dm;
      double var = s2 / (double)n - dm * dm;
      if (var < 0.0) var = 0.0;
      out0[c] = (float)mean;
      out0[channels + c] = (float)(1.0 / sqrt(var + (double)eps));
      if (out1 != nullptr) {             // running_mean / running_var
        const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
        out1[c] = (float)((1.0 - momentum) * (double)out1[c] + (double)momentum * mean);
        out2[c] = (float)((1.0 - momentum) * (double)out2[c] + (double)momentum * unbiased);
      }
    } else if (MODE == 2) {
      out0[c] = (float)s1 + (accumulate ? out0[c] : 0.f);
    } else {
      out0[c] = (float)(s1 / (double)n);
      out0[channels + c] = (float)(s2 / (double)n);
      out1[c] = (float)s2 + (accumulate ? out1[c] : 0.f);   // dgamma
      out2[c] = (float)s1 + (accumulate ? out2[c] : 0.f);   // dbeta
    }
  }
}

}  // namespace ft3d
