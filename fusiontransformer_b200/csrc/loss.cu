// Segmentation loss and metric on the device.
//
// Reference: modules/SemanticTorchpackTrainer.py:70-108 -- F.cross_entropy(logits, labels, weight=class_weights)
// (weighted mean over the non-ignored rows) mixed with the cross-modal term
// F.kl_div(log_softmax(student), softmax(teacher.detach()), 'none').sum(1).mean() as (1-l)*CE + l*KL; and
// models/metric.py:37-58 -- SegIoU: confusion matrix of argmax(logits) against the labels (which the reference
// builds on the CPU after two .cpu() round trips per step).
//
// One pass over the [n, C] logits produces the loss terms AND the gradient: a thread owns a row (C <= 64 values, read
// twice through L1), the weighted-mean denominator comes from a label-only pre-pass, every sum is a fixed-shape
// per-CTA tree followed by a serial double-precision fold (deterministic).  HBM-bound: 2 x 4C bytes per row.
#include "common.cuh"

namespace ft3d {

constexpr int kLossThreads = 256;
constexpr int kLossMaxCtas = 1024;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < kLossThreads / 32; ++i) t += s_red[i];
  return t;                                                // valid in thread 0
}

__device__ __forceinline__ int64_t rows_of(int64_t n, const int32_t* __restrict__ valid_rows) {
  if (valid_rows == nullptr) return n;
  const int64_t v = (int64_t)__ldg(valid_rows);
  return v < n ? v : n;
}

// partial[cta] = sum over the CTA's rows of weight[label] (1 if no weights) for labels != ignore_index
__global__ void __launch_bounds__(kLossThreads)
loss_den_kernel(const int64_t* __restrict__ labels, int64_t n, int C, int64_t ignore_index,
                const float* __restrict__ weight, const int32_t* __restrict__ valid_rows, double* __restrict__ partial) {
  pdl_enter();
  __shared__ double s_red[kLossThreads / 32];
  n = rows_of(n, valid_rows);
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = labels[i];
    if (y != ignore_index && y >= 0 && y < C) acc += weight ? (double)__ldg(weight + y) : 1.0;
  }
  const double t = block_sum(acc, s_red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// grad[i,:] = (1-l)/den * w_i (softmax_i - onehot_i)  +  l/n * (softmax_i - softmax(teacher_i));
// partial[nden + 2*cta + {0,1}] = this CTA's sums of w_i * nll_i and of KL_i
__global__ void __launch_bounds__(kLossThreads)
loss_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t n_cap, int C,
                 int64_t ignore_index, const float* __restrict__ weight, const float* __restrict__ teacher,
                 float lambda_xm, const int32_t* __restrict__ valid_rows, int nden, double* __restrict__ partial,
                 float* __restrict__ grad) {
  pdl_enter();
  __shared__ double s_red[kLossThreads / 32];
  __shared__ double s_den;
  const int64_t n = rows_of(n_cap, valid_rows);
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int i = 0; i < nden; ++i) d += partial[i];        // fixed order: identical in every CTA
    s_den = d;
  }
  __syncthreads();
  const double den = s_den;
  const float ce_scale = den > 0.0 ? (float)((1.0 - (double)lambda_xm) / den) : 0.f;
  const float kl_scale = (teacher != nullptr && n > 0) ? lambda_xm / (float)n : 0.f;
  double a_nll = 0.0, a_kl = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_cap; i += (int64_t)gridDim.x * blockDim.x) {
    float* g = grad + i * C;
    if (i >= n) {                                          // capacity padding: no gradient
      for (int j = 0; j < C; ++j) g[j] = 0.f;
      continue;
    }
    const float* x = logits + i * C;
    float m = -INFINITY;
    for (int j = 0; j < C; ++j) m = fmaxf(m, x[j]);
    float se = 0.f;
    for (int j = 0; j < C; ++j) se += __expf(x[j] - m);
    const float lse = m + __logf(se);
    const int64_t y = labels[i];
    const bool live = (y != ignore_index && y >= 0 && y < C);
    const float w = live ? (weight ? __ldg(weight + y) : 1.f) : 0.f;
    if (live) a_nll += (double)(w * (lse - x[y]));
    float tm = 0.f, tlse = 0.f;
    const float* t = teacher ? teacher + i * C : nullptr;
    if (t != nullptr) {
      tm = -INFINITY;
      for (int j = 0; j < C; ++j) tm = fmaxf(tm, t[j]);
      float tse = 0.f;
      for (int j = 0; j < C; ++j) tse += __expf(t[j] - tm);
      tlse = tm + __logf(tse);
    }
    float kl = 0.f;
    for (int j = 0; j < C; ++j) {
      const float ls = x[j] - lse;                         // log softmax of the student
      const float s = __expf(ls);
      float gj = ce_scale * w * (s - ((live && j == (int)y) ? 1.f : 0.f));
      if (t != nullptr) {
        const float lt = t[j] - tlse;
        const float tp = __expf(lt);
        kl += tp * (lt - ls);
        gj += kl_scale * (s - tp);
      }
      g[j] = gj;
    }
    a_kl += (double)kl;
  }
  const double t0 = block_sum(a_nll, s_red);
  const double t1 = block_sum(a_kl, s_red);
  if (threadIdx.x == 0) {
    partial[nden + 2 * blockIdx.x] = t0;
    partial[nden + 2 * blockIdx.x + 1] = t1;
  }
}

// out[0] = loss, out[1] = CE, out[2] = KL, out[3] = den
__global__ void loss_finalize_kernel(const double* __restrict__ partial, int nden, int nrow_ctas, int64_t n_cap,
                                     const int32_t* __restrict__ valid_rows, float lambda_xm, int has_teacher,
                                     float* __restrict__ out) {
  pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double den = 0.0, nll = 0.0, kl = 0.0;
  for (int i = 0; i < nden; ++i) den += partial[i];
  for (int i = 0; i < nrow_ctas; ++i) {
    nll += partial[nden + 2 * i];
    kl += partial[nden + 2 * i + 1];
  }
  const int64_t n = rows_of(n_cap, valid_rows);
  const double ce = nll / den;                             // 0/0 = NaN when every label is ignored, as torch
  const double klm = (has_teacher && n > 0) ? kl / (double)n : 0.0;
  out[0] = (float)(has_teacher ? (1.0 - (double)lambda_xm) * ce + (double)lambda_xm * klm : ce);
  out[1] = (float)ce;
  out[2] = (float)klm;
  out[3] = (float)den;
}

// mat[label, argmax(logits)] += 1 for labels != ignore_index (int64 counters; integer atomics: order-independent)
__global__ void __launch_bounds__(kLossThreads)
confusion_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t n, int C,
                 int64_t ignore_index, const int32_t* __restrict__ valid_rows, unsigned long long* __restrict__ mat) {
  pdl_enter();
  extern __shared__ unsigned int s_hist[];
  n = rows_of(n, valid_rows);
  for (int t = threadIdx.x; t < C * C; t += blockDim.x) s_hist[t] = 0u;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = labels[i];
    if (y == ignore_index || y < 0 || y >= C) continue;
    const float* x = logits + i * C;
    int best = 0;
    float bv = x[0];
    for (int j = 1; j < C; ++j)
      if (x[j] > bv) bv = x[j], best = j;                  // first maximum, as torch.argmax on the CPU
    atomicAdd(&s_hist[(int)y * C + best], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < C * C; t += blockDim.x)
    if (s_hist[t]) atomicAdd(mat + t, (unsigned long long)s_hist[t]);
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

size_t ft3d_seg_loss_workspace(void) { return (size_t)(3 * kLossMaxCtas) * sizeof(double); }

int ft3d_seg_loss(const float* logits, const int64_t* labels, int64_t n, int32_t num_classes, int64_t ignore_index,
                  const float* class_weight, const float* teacher_logits, float lambda_xm, const int32_t* valid_rows,
                  float* loss_out, float* grad_out, void* workspace, size_t workspace_bytes, ft3d_stream_t stream) {
  FT3D_REQUIRE(logits && labels && loss_out && grad_out && workspace && n > 0 && num_classes >= 2 && num_classes <= 64,
               "ft3d_seg_loss: bad arguments (n=%lld, classes=%d)", (long long)n, num_classes);
  FT3D_REQUIRE(workspace_bytes >= ft3d_seg_loss_workspace() && ((uintptr_t)workspace & 7) == 0,
               "ft3d_seg_loss: workspace too small or misaligned");
  FT3D_REQUIRE(lambda_xm >= 0.f && lambda_xm <= 1.f && (teacher_logits != nullptr || lambda_xm == 0.f),
               "ft3d_seg_loss: lambda_xm must be in [0,1] and needs teacher logits when positive");
  cudaStream_t s = (cudaStream_t)stream;
  int ctas = (int)((n + kLossThreads - 1) / kLossThreads);
  if (ctas > kLossMaxCtas) ctas = kLossMaxCtas;
  double* partial = (double*)workspace;
  launch_pdl(loss_den_kernel, dim3(ctas), dim3(kLossThreads), 0, s, labels, n, num_classes, ignore_index, class_weight,
             valid_rows, partial);
  launch_pdl(loss_rows_kernel, dim3(ctas), dim3(kLossThreads), 0, s, logits, labels, n, num_classes, ignore_index,
             class_weight, lambda_xm > 0.f ? teacher_logits : (const float*)nullptr, lambda_xm, valid_rows, ctas, partial,
             grad_out);
  launch_pdl(loss_finalize_kernel, dim3(1), dim3(32), 0, s, (const double*)partial, ctas, ctas, n, valid_rows, lambda_xm,
             (int)(lambda_xm > 0.f), loss_out);
  return check_launch("ft3d_seg_loss");
}

int ft3d_confusion_update(const float* logits, const int64_t* labels, int64_t n, int32_t num_classes,
                          int64_t ignore_index, const int32_t* valid_rows, int64_t* mat, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(logits && labels && mat && num_classes >= 2 && num_classes <= 64, "ft3d_confusion_update: bad arguments");
  int ctas = (int)((n + kLossThreads - 1) / kLossThreads);
  if (ctas > 2 * kNumSMs) ctas = 2 * kNumSMs;
  launch_pdl(confusion_kernel, dim3(ctas), dim3(kLossThreads), (size_t)num_classes * num_classes * sizeof(unsigned int),
             (cudaStream_t)stream, logits, labels, n, num_classes, ignore_index, valid_rows,
             (unsigned long long*)mat);
  return check_launch("ft3d_confusion_update");
}

}  // extern "C"
