// Kernel-map construction (a6/a7, K2+K3+K9 fused).
//
// Layout: the neighbour table is output-stationary and row-major, nbr[n_out][kpad] int32 with
// kpad = 32 (27 offsets) or 8 (2x2x2), -1 = no neighbour.  One row is one 128-byte (or 32-byte)
// line: the map builder writes it with a single coalesced warp store and the convolution kernels
// read a 128-row tile as one contiguous 16 KB block.  The reference's offset-major pair list is
// derived from it by a counted, order-preserving compaction (no atomics => bit-exact order).
#include "common.cuh"

namespace ft3d {

// Warp-cooperative probing: the 32 lanes of a warp own (32/kpad) output voxels x kpad offsets; each
// lane hashes coord+offset[lane%kpad], probes the table and the warp stores whole table rows.
__global__ void kmap_build_kernel(const int4* __restrict__ coords_q, int64_t n_out,
                                  const int32_t* __restrict__ offsets, int K, int kpad,
                                  const unsigned long long* __restrict__ tkeys,
                                  const int* __restrict__ tvals, uint32_t mask,
                                  int32_t* __restrict__ nbr) {
  const int lane = threadIdx.x & 31;
  const int rows_per_warp = 32 / kpad;
  const int k = lane % kpad;
  const int sub = lane / kpad;
  int ox = 0, oy = 0, oz = 0;
  if (k < K) { ox = __ldg(offsets + 3 * k); oy = __ldg(offsets + 3 * k + 1); oz = __ldg(offsets + 3 * k + 2); }
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp0 * rows_per_warp; base < n_out; base += nwarps * rows_per_warp) {
    int64_t row = base + sub;
    int v = -1;
    if (row < n_out && k < K) {
      int4 c = __ldg(coords_q + row);
      v = table_lookup(tkeys, tvals, mask, fnv1a_fold(c.x + ox, c.y + oy, c.z + oz, c.w));
    }
    if (row < n_out) nbr[row * kpad + k] = v;
  }
}

constexpr int kChunkRows = 256;  // rows per CTA in the compaction passes (8 warps x 32 rows)

// pass 1: per-chunk, per-offset valid counts.  smem tile is transposed with +1 padding.
__global__ void __launch_bounds__(kChunkRows)
kmap_count_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int K, int kpad,
                  int32_t* __restrict__ chunk_counts /*[nchunks][32]*/) {
  __shared__ int s_cnt[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kChunkRows + threadIdx.x;
  for (int k = 0; k < K; ++k) {
    int v = (row < n_out) ? __ldg(nbr + row * kpad + k) : -1;
    unsigned b = __ballot_sync(0xffffffffu, v >= 0);
    if (lane == 0) s_cnt[warp][k] = __popc(b);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = 0;
    if (threadIdx.x < K)
      for (int w = 0; w < 8; ++w) t += s_cnt[w][threadIdx.x];
    chunk_counts[(int64_t)blockIdx.x * 32 + threadIdx.x] = t;
  }
}

// pass 2: one warp per offset scans its column of chunk counts; then the K totals are prefixed.
__global__ void __launch_bounds__(1024)
kmap_scan_kernel(int32_t* __restrict__ chunk_counts, int64_t nchunks, int K, int32_t* __restrict__ offsets_out) {
  __shared__ int s_tot[32];
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
  int running = 0;
  if (k < K) {
    for (int64_t c0 = 0; c0 < nchunks; c0 += 32) {
      int64_t c = c0 + lane;
      int v = (c < nchunks) ? chunk_counts[c * 32 + k] : 0;
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (c < nchunks) chunk_counts[c * 32 + k] = running + incl - v;  // exclusive offset inside offset k
      running += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if (lane == 0) s_tot[k] = (k < K) ? running : 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int j = 0; j < K; ++j) { offsets_out[j] = acc; acc += s_tot[j]; }
    offsets_out[K] = acc;
  }
}

// pass 3: ordered emit.
__global__ void __launch_bounds__(kChunkRows)
kmap_emit_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int K, int kpad,
                 const int32_t* __restrict__ chunk_offsets, const int32_t* __restrict__ offsets,
                 int2* __restrict__ pairs, int32_t* __restrict__ ppos) {
  __shared__ int s_cnt[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kChunkRows + threadIdx.x;
  int vals[32];
  unsigned ballots[32];
  #pragma unroll
  for (int k = 0; k < 32; ++k) {
    int v = -1;
    if (k < K && row < n_out) v = __ldg(nbr + row * kpad + k);
    vals[k] = v;
    unsigned b = __ballot_sync(0xffffffffu, v >= 0);
    ballots[k] = b;
    if (lane == 0) s_cnt[warp][k] = __popc(b);
  }
  __syncthreads();
  int filled = 0;                         // ppos rows are stored compacted: valid positions first, ascending offset
  #pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k < K && vals[k] >= 0) {
      int pos = __ldg(offsets + k) + __ldg(chunk_offsets + (int64_t)blockIdx.x * 32 + k);
      for (int w = 0; w < warp; ++w) pos += s_cnt[w][k];
      pos += __popc(ballots[k] & ((1u << lane) - 1u));
      pairs[pos] = make_int2(vals[k], (int)row);
      if (ppos) ppos[row * kpad + filled++] = pos;
    }
  }
  if (ppos && row < n_out)
    for (int j = filled; j < kpad; ++j) ppos[row * kpad + j] = -1;
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void kmap_transpose_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int K, int kpad,
                                      int32_t* __restrict__ nbrT, int64_t n_in) {
  int64_t total = n_out * kpad;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(t % kpad);
    int64_t j = t / kpad;
    int i = __ldg(nbr + t);
    if (k < K && i >= 0 && i < n_in) nbrT[(int64_t)i * kpad + k] = (int32_t)j;
  }
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

int ft3d_kmap_build(const int32_t* coords_q, int64_t n_out, const int32_t* offsets, int32_t K,
                    const uint64_t* table_keys, const int32_t* table_vals, int64_t cap,
                    int32_t* nbr_out, int32_t kpad, ft3d_stream_t stream) {
  if (n_out == 0) return FT3D_OK;
  FT3D_REQUIRE(coords_q && offsets && table_keys && table_vals && nbr_out, "ft3d_kmap_build: null pointer");
  FT3D_REQUIRE(K > 0 && K <= kpad && (kpad == 8 || kpad == 16 || kpad == 32),
               "ft3d_kmap_build: need 0 < K <= kpad and kpad in {8,16,32} (got K=%d kpad=%d)", K, kpad);
  FT3D_REQUIRE(cap >= 2 && (cap & (cap - 1)) == 0, "ft3d_kmap_build: capacity must be a power of two");
  int rows_per_warp = 32 / kpad;
  int64_t warps = (n_out + rows_per_warp - 1) / rows_per_warp;
  kmap_build_kernel<<<grid_for(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      (const int4*)coords_q, n_out, offsets, K, kpad, (const unsigned long long*)table_keys, table_vals,
      (uint32_t)(cap - 1), nbr_out);
  return check_launch("ft3d_kmap_build");
}

size_t ft3d_kmap_pairs_workspace(int64_t n_out, int32_t kpad) {
  (void)kpad;
  int64_t nchunks = (n_out + kChunkRows - 1) / kChunkRows;
  return align_up((size_t)(nchunks > 0 ? nchunks : 1) * 32 * sizeof(int32_t), 256);
}

int ft3d_kmap_pairs(const int32_t* nbr, int64_t n_out, int32_t K, int32_t kpad, int32_t* pairs_out,
                    int32_t* offsets_out, int32_t* ppos_out, void* workspace, size_t workspace_bytes,
                    ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(offsets_out && K > 0 && K <= 32 && K <= kpad, "ft3d_kmap_pairs: bad arguments");
  if (n_out == 0) {
    FT3D_CUDA(cudaMemsetAsync(offsets_out, 0, (K + 1) * sizeof(int32_t), s));
    return FT3D_OK;
  }
  FT3D_REQUIRE(nbr && pairs_out && workspace, "ft3d_kmap_pairs: null pointer");
  if (workspace_bytes < ft3d_kmap_pairs_workspace(n_out, kpad)) {
    set_error("ft3d_kmap_pairs: workspace %zu < %zu", workspace_bytes, ft3d_kmap_pairs_workspace(n_out, kpad));
    return FT3D_ERR_WORKSPACE;
  }
  int64_t nchunks = (n_out + kChunkRows - 1) / kChunkRows;
  int32_t* chunk_counts = (int32_t*)workspace;
  kmap_count_kernel<<<(unsigned)nchunks, kChunkRows, 0, s>>>(nbr, n_out, K, kpad, chunk_counts);
  kmap_scan_kernel<<<1, 1024, 0, s>>>(chunk_counts, nchunks, K, offsets_out);
  kmap_emit_kernel<<<(unsigned)nchunks, kChunkRows, 0, s>>>(nbr, n_out, K, kpad, chunk_counts, offsets_out,
                                                            (int2*)pairs_out, ppos_out);
  return check_launch("ft3d_kmap_pairs");
}

int ft3d_kmap_transpose(const int32_t* nbr, int64_t n_out, int32_t K, int32_t kpad, int32_t* nbrT_out,
                        int64_t n_in, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(K > 0 && K <= kpad, "ft3d_kmap_transpose: bad arguments");
  if (n_in > 0) {
    FT3D_REQUIRE(nbrT_out != nullptr, "ft3d_kmap_transpose: null output");
    fill_i32_kernel<<<grid_for(n_in * kpad, 256), 256, 0, s>>>(nbrT_out, n_in * kpad, -1);
  }
  if (n_out > 0 && n_in > 0) {
    FT3D_REQUIRE(nbr != nullptr, "ft3d_kmap_transpose: null input");
    kmap_transpose_kernel<<<grid_for(n_out * kpad, 256), 256, 0, s>>>(nbr, n_out, K, kpad, nbrT_out, n_in);
  }
  return check_launch("ft3d_kmap_transpose");
}

}  // extern "C"
