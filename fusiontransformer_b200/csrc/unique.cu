// Sorted-unique of voxel keys (K4 / torch.unique) and the GPU sparse_quantize (a2 / K15).
// Ordering is an *output* of the reference (ascending key), so dedup is sort-based: a stable
// device radix sort (CUB onesweep) of (key,row) pairs, head flags, a prefix sum and one scatter
// pass.  The hand-written kernels around the sort are one element per thread and HBM-bound.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace ft3d {

struct UniqueWs {
  int64_t* keys_sorted;
  int32_t* idx_in;
  int32_t* idx_sorted;
  int32_t* flags;
  int32_t* gid;
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static size_t cub_bytes_for(int64_t n) {
  size_t a = 0, b = 0, c = 0;
  int num = (int)n;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, num, 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, c, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, num, 0, 32);
  cub::DeviceScan::InclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, num);
  size_t m = a > b ? a : b;
  m = m > c ? m : c;
  return align_up(m + 256, 256);
}

static UniqueWs carve(void* ws, int64_t n) {
  UniqueWs w;
  char* p = (char*)ws;
  size_t n8 = align_up((size_t)n * 8, 256), n4 = align_up((size_t)n * 4, 256);
  w.keys_sorted = (int64_t*)p; p += n8;
  w.idx_in = (int32_t*)p; p += n4;
  w.idx_sorted = (int32_t*)p; p += n4;
  w.flags = (int32_t*)p; p += n4;
  w.gid = (int32_t*)p; p += n4;
  w.cub_tmp = p;
  w.cub_bytes = cub_bytes_for(n);
  p += w.cub_bytes;
  w.total = (size_t)(p - (char*)ws);
  return w;
}

__global__ void iota_kernel(int32_t* p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

__global__ void head_flags_kernel(const int64_t* __restrict__ ks, int64_t n, int32_t* __restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flags[i] = (i == 0 || ks[i] != ks[i - 1]) ? 1 : 0;
}

__global__ void unique_scatter_kernel(const int64_t* __restrict__ ks, const int32_t* __restrict__ idx_sorted,
                                      const int32_t* __restrict__ flags, const int32_t* __restrict__ gid_incl,
                                      int64_t n, int64_t* __restrict__ uniq, int32_t* __restrict__ inverse,
                                      int32_t* __restrict__ head_pos, int32_t* __restrict__ first,
                                      int32_t* __restrict__ num_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int g = gid_incl[i] - 1;
    int row = idx_sorted[i];
    inverse[row] = g;
    if (flags[i]) {
      uniq[g] = ks[i];
      first[g] = row;   // stable sort => smallest row of the group
      head_pos[g] = (int32_t)i;
    }
    if (i + 1 == n) *num_out = g + 1;
  }
}

__global__ void unique_counts_kernel(const int32_t* __restrict__ head_pos, const int32_t* __restrict__ num,
                                     int64_t n, int32_t* __restrict__ counts) {
  int m = *num;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < m; g += (int64_t)gridDim.x * blockDim.x)
    counts[g] = (g + 1 < m ? head_pos[g + 1] : (int32_t)n) - head_pos[g];
}

// ------------------------------------------------------------------ quantize
__global__ void quantize_keys_kernel(const int4* __restrict__ coords, int64_t n, uint64_t* __restrict__ keys,
                                     int32_t* __restrict__ idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    keys[i] = fnv1_vec3(c.x, c.y, c.z);
    idx[i] = (int32_t)i;
  }
}

__global__ void gather_scan_kernel(const int4* __restrict__ coords, const int32_t* __restrict__ idx1, int64_t n,
                                   uint32_t* __restrict__ scan1, int32_t* __restrict__ pos) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    scan1[i] = (uint32_t)__ldg(coords + idx1[i]).w;
    pos[i] = (int32_t)i;
  }
}

__global__ void quantize_heads_kernel(const uint64_t* __restrict__ ks1, const int32_t* __restrict__ perm2,
                                      const uint32_t* __restrict__ scan2, int64_t n, int32_t* __restrict__ flags) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    int f = 1;
    if (q > 0) f = (scan2[q] != scan2[q - 1]) || (ks1[perm2[q]] != ks1[perm2[q - 1]]);
    flags[q] = f;
  }
}

__global__ void quantize_emit_heads_kernel(const int32_t* __restrict__ idx1, const int32_t* __restrict__ perm2,
                                           const uint32_t* __restrict__ scan2, const int32_t* __restrict__ flags,
                                           const int32_t* __restrict__ gid_incl, int64_t n,
                                           int32_t* __restrict__ inds, int32_t* __restrict__ scan_first,
                                           int32_t* __restrict__ scan_counts, int32_t* __restrict__ num_out) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    int g = gid_incl[q] - 1;
    if (flags[q]) {
      inds[g] = idx1[perm2[q]];
      atomicAdd(scan_counts + scan2[q], 1);
      if (q == 0 || scan2[q] != scan2[q - 1]) scan_first[scan2[q]] = g;
    }
    if (q + 1 == n) *num_out = g + 1;
  }
}

__global__ void quantize_inverse_kernel(const int32_t* __restrict__ idx1, const int32_t* __restrict__ perm2,
                                        const uint32_t* __restrict__ scan2, const int32_t* __restrict__ gid_incl,
                                        const int32_t* __restrict__ scan_first, int64_t n,
                                        int32_t* __restrict__ inverse) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    inverse[idx1[perm2[q]]] = gid_incl[q] - 1 - scan_first[scan2[q]];
}

__global__ void zero_i32_kernel(int32_t* p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0;
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

size_t ft3d_unique_workspace(int64_t n) {
  if (n <= 0) return 256;
  return align_up((size_t)n * 8, 256) + 4 * align_up((size_t)n * 4, 256) + cub_bytes_for(n) + 256;
}

int ft3d_unique(const int64_t* keys, int64_t n, int64_t* uniq_out, int32_t* inverse_out,
                int32_t* counts_out, int32_t* first_out, int32_t* num_out, void* workspace,
                size_t workspace_bytes, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(num_out != nullptr, "ft3d_unique: null num_out");
  if (n == 0) {
    FT3D_CUDA(cudaMemsetAsync(num_out, 0, sizeof(int32_t), s));
    return FT3D_OK;
  }
  FT3D_REQUIRE(keys && uniq_out && inverse_out && counts_out && first_out && workspace && n < (1LL << 31),
               "ft3d_unique: bad arguments");
  if (workspace_bytes < ft3d_unique_workspace(n)) {
    set_error("ft3d_unique: workspace %zu < %zu", workspace_bytes, ft3d_unique_workspace(n));
    return FT3D_ERR_WORKSPACE;
  }
  UniqueWs w = carve(workspace, n);
  int g = grid_for(n, 256);
  iota_kernel<<<g, 256, 0, s>>>(w.idx_in, n);
  size_t tb = w.cub_bytes;
  FT3D_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, keys, w.keys_sorted, w.idx_in, w.idx_sorted,
                                            (int)n, 0, 64, s));
  head_flags_kernel<<<g, 256, 0, s>>>(w.keys_sorted, n, w.flags);
  tb = w.cub_bytes;
  FT3D_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.flags, w.gid, (int)n, s));
  unique_scatter_kernel<<<g, 256, 0, s>>>(w.keys_sorted, w.idx_sorted, w.flags, w.gid, n, uniq_out, inverse_out,
                                          w.idx_in /* head positions; iota no longer needed */, first_out, num_out);
  unique_counts_kernel<<<g, 256, 0, s>>>(w.idx_in, num_out, n, counts_out);
  return check_launch("ft3d_unique");
}

size_t ft3d_quantize_workspace(int64_t n, int32_t num_scans) {
  if (n <= 0) return 256;
  // keys, keys_sorted (u64) ; idx0, idx1, scan1, scan2, pos, perm2, flags, gid (32-bit) ; scan_first ; cub
  return 2 * align_up((size_t)n * 8, 256) + 8 * align_up((size_t)n * 4, 256) +
         align_up((size_t)(num_scans > 0 ? num_scans : 1) * 4, 256) + cub_bytes_for(n) + 256;
}

int ft3d_quantize(const int32_t* coords, int64_t n, int32_t num_scans, int32_t* inds_out,
                  int32_t* inverse_out, int32_t* scan_counts_out, int32_t* num_unique_out,
                  void* workspace, size_t workspace_bytes, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(num_unique_out && scan_counts_out && num_scans > 0, "ft3d_quantize: bad arguments");
  zero_i32_kernel<<<1, 256, 0, s>>>(scan_counts_out, num_scans);
  if (n == 0) {
    FT3D_CUDA(cudaMemsetAsync(num_unique_out, 0, sizeof(int32_t), s));
    return check_launch("ft3d_quantize");
  }
  FT3D_REQUIRE(coords && inds_out && inverse_out && workspace && n < (1LL << 31), "ft3d_quantize: bad arguments");
  if (workspace_bytes < ft3d_quantize_workspace(n, num_scans)) {
    set_error("ft3d_quantize: workspace %zu < %zu", workspace_bytes, ft3d_quantize_workspace(n, num_scans));
    return FT3D_ERR_WORKSPACE;
  }
  char* p = (char*)workspace;
  size_t n8 = align_up((size_t)n * 8, 256), n4 = align_up((size_t)n * 4, 256);
  uint64_t* keys = (uint64_t*)p; p += n8;
  uint64_t* ks1 = (uint64_t*)p; p += n8;
  int32_t* idx0 = (int32_t*)p; p += n4;
  int32_t* idx1 = (int32_t*)p; p += n4;
  uint32_t* scan1 = (uint32_t*)p; p += n4;
  uint32_t* scan2 = (uint32_t*)p; p += n4;
  int32_t* pos = (int32_t*)p; p += n4;
  int32_t* perm2 = (int32_t*)p; p += n4;
  int32_t* flags = (int32_t*)p; p += n4;
  int32_t* gid = (int32_t*)p; p += n4;
  int32_t* scan_first = (int32_t*)p; p += align_up((size_t)num_scans * 4, 256);
  void* cub_tmp = p;
  size_t cub_bytes = cub_bytes_for(n);

  int g = grid_for(n, 256);
  quantize_keys_kernel<<<g, 256, 0, s>>>((const int4*)coords, n, keys, idx0);
  size_t tb = cub_bytes;
  FT3D_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, keys, ks1, idx0, idx1, (int)n, 0, 64, s));
  gather_scan_kernel<<<g, 256, 0, s>>>((const int4*)coords, idx1, n, scan1, pos);
  int bits = 1;
  while ((1 << bits) < num_scans) ++bits;
  tb = cub_bytes;
  FT3D_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, scan1, scan2, pos, perm2, (int)n, 0, bits, s));
  quantize_heads_kernel<<<g, 256, 0, s>>>(ks1, perm2, scan2, n, flags);
  tb = cub_bytes;
  FT3D_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, tb, flags, gid, (int)n, s));
  quantize_emit_heads_kernel<<<g, 256, 0, s>>>(idx1, perm2, scan2, flags, gid, n, inds_out, scan_first,
                                               scan_counts_out, num_unique_out);
  quantize_inverse_kernel<<<g, 256, 0, s>>>(idx1, perm2, scan2, gid, scan_first, n, inverse_out);
  return check_launch("ft3d_quantize");
}

}  // extern "C"
