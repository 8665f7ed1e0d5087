// Tile schedule of the output-stationary sparse convolution (conv_os.cu).
//
// A LiDAR kernel map is sparse: 2-7 of the 27 offsets are occupied per voxel.  An output-stationary tile of 128
// arbitrary output rows therefore touches nearly every offset (11-27 of 27) and would spend 5-12x the useful gather
// and MMA work.  The schedule built here makes output-stationary tiles dense instead:
//   1. every output row gets its occupancy mask (bit k set <=> the row has a neighbour at offset k);
//   2. rows are sorted by that mask with the bits re-ranked so that the RAREST offsets are the most significant sort
//      digits: rows that share the rare offsets end up in the same tiles and the common offsets (which most rows of a
//      tile use anyway) vary fastest.  Measured on the synthetic scans: 67-82 % of the (row, offset) slots a tile
//      visits are real pairs, against 13-25 % for the natural (hash) row order (tools/tile_occupancy_study.py);
//   3. per tile of 128 sorted rows the union of the masks is the list of "passes"; pass (tile, k) stores the 128
//      gather indices table[row][k] (-1 = absent: the row is zero-filled in the A operand).
// Outputs (all int32):
//   tiles    [T][2]      (first pass, number of passes) of tile t;  T = ceil(n_rows / 128)
//   out_row  [T*128]     output row of tile slot (sorted order), -1 = empty slot
//   pass_k   [P]         offset index of pass p                      (P = sum of passes, *num_pass_out)
//   pass_idx [P][128]    gather row per tile slot or -1
// Everything is integer, deterministic (stable radix sort, no data-dependent atomics order in the result) and built
// once per (map, side) on the geometry stream.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace ft3d {

constexpr int kOsTile = 128;        // rows of one CTA tile; a schedule tile is tile_rows = 128 x (CTAs of a cluster)

// masks + per-offset occupancy counts.  A warp owns 32/kpad rows per iteration; lane -> (row sub, offset k).
__global__ void os_mask_kernel(const int32_t* __restrict__ table, int64_t n_rows, int K, int kpad,
                               uint32_t* __restrict__ mask, int32_t* __restrict__ counts /*[32], zeroed*/) {
  pdl_enter();
  __shared__ int s_cnt[32];
  if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int rows_per_warp = 32 / kpad;
  const int k = lane % kpad, sub = lane / kpad;
  const uint32_t kmask = kpad == 32 ? 0xffffffffu : ((1u << kpad) - 1u);
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int mine = 0;
  for (int64_t base = warp0 * rows_per_warp; base < n_rows; base += nwarps * rows_per_warp) {
    const int64_t row = base + sub;
    const bool hit = row < n_rows && k < K && __ldg(table + row * kpad + k) >= 0;
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    mine += hit ? 1 : 0;
    if (k == 0 && row < n_rows) mask[row] = (b >> (sub * kpad)) & kmask;
  }
  if (mine) atomicAdd(&s_cnt[k], mine);
  __syncthreads();
  if (threadIdx.x < 32 && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, s_cnt[threadIdx.x]);
}

// rank[k] = sort-digit position of offset k: the most frequent offset gets bit 0, the rarest the top bit
// (ties: lower k = lower bit).  One warp.
__global__ void os_rank_kernel(const int32_t* __restrict__ counts, int K, int32_t* __restrict__ rank) {
  pdl_enter();
  const int k = threadIdx.x;
  const int ck = k < K ? counts[k] : -1;
  int r = 0;
  for (int j = 0; j < K; ++j) {
    const int cj = counts[j];
    if (cj > ck || (cj == ck && j < k)) ++r;
  }
  if (k < K) rank[k] = r;
}

__global__ void os_key_kernel(const uint32_t* __restrict__ mask, int64_t n_rows, int K,
                              const int32_t* __restrict__ rank, uint32_t* __restrict__ key, int32_t* __restrict__ row_id) {
  pdl_enter();
  __shared__ int s_rank[32];
  if ((int)threadIdx.x < K) s_rank[threadIdx.x] = rank[threadIdx.x];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t m = mask[i], out = 0;
    while (m) {
      const int k = __ffs(m) - 1;
      m &= m - 1;
      out |= 1u << s_rank[k];
    }
    key[i] = out;
    row_id[i] = (int32_t)i;
  }
}

// one warp per tile: union of the masks of its 128 sorted rows
__global__ void os_tile_union_kernel(const int32_t* __restrict__ sorted_rows, const uint32_t* __restrict__ mask,
                                     int64_t n_rows, int T, int tile_rows, uint32_t* __restrict__ tile_union) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp0; t < T; t += nwarps) {
    uint32_t u = 0;
    for (int i = 0; i < tile_rows / 32; ++i) {
      const int64_t s = t * tile_rows + i * 32 + lane;
      if (s < n_rows) u |= __ldg(mask + __ldg(sorted_rows + s));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) u |= __shfl_xor_sync(0xffffffffu, u, o);
    if (lane == 0) tile_union[t] = u;
  }
}

// units of a tile with n passes: ceil(n / cap), at most 4 (the fold of conv_os keeps 4 partial quads per row in
// registers); the passes are then split evenly
constexpr int kOsMaxChunks = 4;
__host__ __device__ __forceinline__ int os_chunks(int n, int cap) {
  if (n <= cap) return 1;
  const int c = (n + cap - 1) / cap;
  return c < kOsMaxChunks ? c : kOsMaxChunks;
}

// Block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ int os_block_scan(int v, int* s_warp, int* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += u;
    }
    s_warp[lane] = wi - w;              // exclusive prefix of the warp totals
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  const int excl = s_warp[warp] + incl - v;
  *total = s_warp[32];
  __syncthreads();
  return excl;
}

// One CTA: pass prefix over the tiles, the chunk size ("cap") of the work units, then unit / scratch-slot prefixes.
//   tile_info[t] = (first pass, passes, first unit, first scratch slot);  split_idx[t] = rank among the split tiles
//   num_out      = {P passes, U units, S scratch slots, cap, split tiles, 0, 0, 0}
// A tile with more than `cap` passes is split into ceil(passes/cap) units that accumulate disjoint pass ranges; their
// partial tiles meet in scratch slots and the unit that finishes last folds them in chunk order (conv_os.cu).  Heavy
// tiles (the tail of the mask order, whose rows use many different rare offsets) would otherwise set the makespan:
// one CTA walking 23-27 passes while the average CTA has 7.
__global__ void __launch_bounds__(1024)
os_tile_scan_kernel(const uint32_t* __restrict__ tile_union, int T, int grid_ctas, int cap_arg,
                    int4* __restrict__ tile_info, int32_t* __restrict__ split_idx, int32_t* __restrict__ num_out) {
  pdl_enter();
  __shared__ int s_warp[33];
  __shared__ int s_carry[4];
  const int tid = threadIdx.x;
  if (tid < 4) s_carry[tid] = 0;
  __syncthreads();
  for (int base = 0; base < T; base += 1024) {
    const int t = base + tid;
    const int n = t < T ? __popc(tile_union[t]) : 0;
    int tot;
    const int ex = os_block_scan(n, s_warp, &tot);
    if (t < T) tile_info[t].x = s_carry[0] + ex, tile_info[t].y = n;
    __syncthreads();
    if (tid == 0) s_carry[0] += tot;
    __syncthreads();
  }
  const int P = s_carry[0];
  int cap = cap_arg;
  if (cap <= 0) {                       // auto: ~3/4 of the mean number of passes per CTA, within [4, 8]
    const int mean = (P + grid_ctas - 1) / grid_ctas;
    cap = (3 * mean + 3) / 4;
    cap = cap < 4 ? 4 : (cap > 8 ? 8 : cap);
  }
  for (int base = 0; base < T; base += 1024) {
    const int t = base + tid;
    const int n = t < T ? tile_info[t].y : 0;
    const int chunks = t < T ? os_chunks(n, cap) : 0;
    int totu, tots, totn;
    const int exu = os_block_scan(chunks, s_warp, &totu);
    const int exs = os_block_scan(chunks > 1 ? chunks : 0, s_warp, &tots);
    const int exn = os_block_scan(chunks > 1 ? 1 : 0, s_warp, &totn);
    if (t < T) tile_info[t].z = s_carry[1] + exu, tile_info[t].w = s_carry[2] + exs, split_idx[t] = s_carry[3] + exn;
    __syncthreads();
    if (tid == 0) s_carry[1] += totu, s_carry[2] += tots, s_carry[3] += totn;
    __syncthreads();
  }
  if (tid == 0) {
    num_out[0] = P;
    num_out[1] = s_carry[1];
    num_out[2] = s_carry[2];
    num_out[3] = cap;
    num_out[4] = s_carry[3];
    num_out[5] = num_out[6] = num_out[7] = 0;
  }
}

// one CTA (128 threads) per tile: slot r -> output row, per pass the gather index of the slot, and the tile's units
//   unit = {first pass, passes, tile, chunks of the tile, chunk index, first scratch slot of the tile, 0, 0}
__global__ void __launch_bounds__(4 * kOsTile)
os_emit_kernel(const int32_t* __restrict__ table, int kpad, const int32_t* __restrict__ sorted_rows, int64_t n_rows,
               const uint32_t* __restrict__ tile_union, const int4* __restrict__ tile_info,
               const int32_t* __restrict__ split_idx, const int32_t* __restrict__ num, int64_t pass_cap,
               int64_t unit_cap, int64_t split_cap, int32_t* __restrict__ split_tiles, int32_t* __restrict__ out_row,
               int32_t* __restrict__ pass_k, int32_t* __restrict__ pass_idx, int32_t* __restrict__ units,
               uint32_t* __restrict__ unit_key, int32_t* __restrict__ unit_id) {
  pdl_enter();
  const int t = blockIdx.x, r = threadIdx.x, tile_rows = blockDim.x;
  const int64_t s = (int64_t)t * tile_rows + r;
  const int row = s < n_rows ? __ldg(sorted_rows + s) : -1;
  out_row[s] = row;
  uint32_t u = tile_union[t];
  const int4 ti = tile_info[t];
  int64_t p = ti.x;
  while (u) {
    const int k = __ffs(u) - 1;
    u &= u - 1;
    if (p < pass_cap) {                 // the caller sized the arrays from an upper bound; never write past it
      if (r == 0) pass_k[p] = k;
      pass_idx[p * tile_rows + r] = row >= 0 ? __ldg(table + (int64_t)row * kpad + k) : -1;
    }
    ++p;
  }
  const int chunks = os_chunks(ti.y, num[3]);
  if (r == 0 && chunks > 1 && split_idx[t] < split_cap) {
    int32_t* sp = split_tiles + (int64_t)split_idx[t] * 4;          // {tile, units, first scratch slot, 0}
    sp[0] = t; sp[1] = chunks; sp[2] = ti.w; sp[3] = 0;
  }
  if (r < chunks && ti.z + r < unit_cap) {
    // even split: unit c takes floor(n/chunks) passes, the first n % chunks units one more (never empty: chunks <= n)
    const int c = r, q = ti.y / chunks, rem = ti.y % chunks;
    const int np = q + (c < rem ? 1 : 0), p0 = c * q + (c < rem ? c : rem);
    int32_t* un = units + (int64_t)(ti.z + c) * 8;
    un[0] = ti.x + p0; un[1] = np; un[2] = t; un[3] = chunks;
    un[4] = c; un[5] = ti.w; un[6] = 0; un[7] = 0;
    unit_key[ti.z + c] = 255u - (uint32_t)np;                       // longest units first (LPT), stable among equals
    unit_id[ti.z + c] = ti.z + c;
  }
}

__global__ void os_fill_u32_kernel(uint32_t* p, int64_t n, uint32_t v) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// Placement of the work units (one CTA).  conv_os deals unit j * ncl + c to CTA cluster c in round j, so the ORDER of
// the unit array is the schedule.  Units arrive sorted by descending pass count; each round hands its ncl units to the
// clusters in order of their load so far (longest unit -> least loaded cluster): LPT applied round by round.  Dealing
// the sorted list round-robin instead leaves the clusters that drew the long units of round 0 with the long units of
// every later round too (measured makespan 10-12 passes where the mean is 7; this placement: 8).  A last, partial
// round leaves holes (tile = -1, no passes) at the clusters that are already the most loaded; num[1] becomes
// rounds * ncl.  If that does not fit unit_cap the sorted order is kept as it is.
__global__ void __launch_bounds__(256)
os_assign_kernel(const int32_t* __restrict__ units, const int32_t* __restrict__ order, int32_t* __restrict__ num,
                 int64_t unit_cap, int ncl, int32_t* __restrict__ out) {
  pdl_enter();
  __shared__ int s_load[256];
  __shared__ int s_bin[256];
  const int t = threadIdx.x;
  const int U = num[1] < unit_cap ? num[1] : (int)unit_cap;
  const int rounds = (U + ncl - 1) / ncl;
  const int4* src = reinterpret_cast<const int4*>(units);
  int4* dst = reinterpret_cast<int4*>(out);
  if ((int64_t)rounds * ncl > unit_cap || ncl > 256) {
    for (int i = t; i < 2 * U; i += blockDim.x) dst[i] = __ldg(src + (int64_t)order[i >> 1] * 2 + (i & 1));
    return;
  }
  s_load[t] = 0;
  __syncthreads();
  for (int j = 0; j < rounds; ++j) {
    if (t < ncl) {
      const int mine = s_load[t];
      int r = 0;
      for (int b = 0; b < ncl; ++b) {
        const int lb = s_load[b];
        r += (lb < mine || (lb == mine && b < t)) ? 1 : 0;
      }
      s_bin[r] = t;                       // the cluster with the r-th smallest load
    }
    __syncthreads();
    if (t < ncl) {
      const int sidx = j * ncl + t, bin = s_bin[t];
      int4 a = make_int4(0, 0, -1, 1), b = make_int4(0, 0, 0, 0);
      if (sidx < U) {
        const int64_t uid = order[sidx];
        a = __ldg(src + uid * 2);
        b = __ldg(src + uid * 2 + 1);
      }
      dst[(int64_t)(j * ncl + bin) * 2] = a;
      dst[(int64_t)(j * ncl + bin) * 2 + 1] = b;
      s_load[bin] += a.y;                 // one unit per cluster and round: no two threads share a bin
    }
    __syncthreads();
  }
  if (t == 0) num[1] = rounds * ncl;
}

__global__ void os_zero_kernel(int32_t* p, int n) {
  pdl_enter();
  if ((int)threadIdx.x < n) p[threadIdx.x] = 0;
}

struct OsPlanWs {
  uint32_t* mask;
  uint32_t* key;
  uint32_t* key_sorted;
  int32_t* row_id;
  int32_t* sorted_rows;
  uint32_t* tile_union;
  int4* tile_info;
  int32_t* split_idx;
  int32_t* units;        // [unit_cap][8] in tile order
  uint32_t* unit_key;
  uint32_t* unit_key_sorted;
  int32_t* unit_id;
  int32_t* unit_order;
  int32_t* counts;     // [32]
  int32_t* rank;       // [32]
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static OsPlanWs os_carve(void* ws, int64_t n, int64_t unit_cap) {     // sized for the smallest tile (128 rows)
  OsPlanWs w;
  char* p = (char*)ws;
  const size_t n4 = align_up((size_t)(n > 0 ? n : 1) * 4, 256);
  const size_t T = (size_t)((n + kOsTile - 1) / kOsTile + 1);
  const size_t u4 = align_up((size_t)(unit_cap > 0 ? unit_cap : 1) * 4, 256);
  w.mask = (uint32_t*)p; p += n4;
  w.key = (uint32_t*)p; p += n4;
  w.key_sorted = (uint32_t*)p; p += n4;
  w.row_id = (int32_t*)p; p += n4;
  w.sorted_rows = (int32_t*)p; p += n4;
  w.tile_union = (uint32_t*)p; p += align_up(T * 4, 256);
  w.tile_info = (int4*)p; p += align_up(T * 16, 256);
  w.split_idx = (int32_t*)p; p += align_up(T * 4, 256);
  w.units = (int32_t*)p; p += 8 * u4;
  w.unit_key = (uint32_t*)p; p += u4;
  w.unit_key_sorted = (uint32_t*)p; p += u4;
  w.unit_id = (int32_t*)p; p += u4;
  w.unit_order = (int32_t*)p; p += u4;
  w.counts = (int32_t*)p; p += 256;
  w.rank = (int32_t*)p; p += 256;
  size_t c = 0, c2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, c, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)(n > 0 ? n : 1), 0, 32);
  cub::DeviceRadixSort::SortPairs(nullptr, c2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)(unit_cap > 0 ? unit_cap : 1), 0, 32);
  if (c2 > c) c = c2;
  w.cub_tmp = p;
  w.cub_bytes = align_up(c + 256, 256);
  p += w.cub_bytes;
  w.total = (size_t)(p - (char*)ws);
  return w;
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

size_t ft3d_conv_os_plan_workspace(int64_t n_rows, int64_t unit_cap) { return os_carve(nullptr, n_rows, unit_cap).total; }

int ft3d_conv_os_plan(const int32_t* table, int64_t n_rows, int32_t K, int32_t kpad, int32_t tile_rows,
                      int64_t pass_cap, int64_t unit_cap, int32_t chunk_passes, int32_t* units_out,
                      int32_t* split_tiles_out,
                      int32_t* out_row_out, int32_t* pass_k_out, int32_t* pass_idx_out, int32_t* num_out,
                      void* workspace, size_t workspace_bytes, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(num_out != nullptr, "ft3d_conv_os_plan: null num_out");
  if (n_rows == 0) {
    launch_pdl(os_zero_kernel, dim3(1), dim3(32), 0, s, num_out, 8);
    return check_launch("ft3d_conv_os_plan");
  }
  FT3D_REQUIRE(table && units_out && split_tiles_out && out_row_out && pass_k_out && pass_idx_out && workspace,
               "ft3d_conv_os_plan: null argument");
  FT3D_REQUIRE(K > 0 && K <= kpad && (kpad == 8 || kpad == 16 || kpad == 32) && pass_cap > 0 && unit_cap > 0 &&
                   n_rows < (1ll << 31) && chunk_passes >= 0 && chunk_passes <= 32 &&
                   (tile_rows == 128 || tile_rows == 256 || tile_rows == 512),
               "ft3d_conv_os_plan: bad arguments");
  OsPlanWs w = os_carve(workspace, n_rows, unit_cap);
  FT3D_REQUIRE(((uintptr_t)workspace & 255) == 0 && workspace_bytes >= w.total,
               "ft3d_conv_os_plan: workspace too small (%zu < %zu) or not 256-byte aligned", workspace_bytes, w.total);
  const int T = (int)((n_rows + tile_rows - 1) / tile_rows);
  const int clusters = kNumSMs / (tile_rows / kOsTile);            // work units are dealt to CTA clusters
  launch_pdl(os_zero_kernel, dim3(1), dim3(32), 0, s, w.counts, 32);
  launch_pdl(os_mask_kernel, dim3(grid_for(n_rows * kpad, 256)), dim3(256), 0, s, table, n_rows, (int)K, (int)kpad,
             w.mask, w.counts);
  launch_pdl(os_rank_kernel, dim3(1), dim3(32), 0, s, (const int32_t*)w.counts, (int)K, w.rank);
  launch_pdl(os_key_kernel, dim3(grid_for(n_rows, 256)), dim3(256), 0, s, (const uint32_t*)w.mask, n_rows, (int)K,
             (const int32_t*)w.rank, w.key, w.row_id);
  size_t cb = w.cub_bytes;
  FT3D_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, (const uint32_t*)w.key, w.key_sorted,
                                            (const int32_t*)w.row_id, w.sorted_rows, (int)n_rows, 0, (int)K, s));
  launch_pdl(os_tile_union_kernel, dim3(grid_for((int64_t)T * 32, 256)), dim3(256), 0, s,
             (const int32_t*)w.sorted_rows, (const uint32_t*)w.mask, n_rows, T, (int)tile_rows, w.tile_union);
  launch_pdl(os_tile_scan_kernel, dim3(1), dim3(1024), 0, s, (const uint32_t*)w.tile_union, T,
             clusters, (int)chunk_passes, w.tile_info, w.split_idx, num_out);      // conv_os always runs every cluster
  launch_pdl(os_fill_u32_kernel, dim3(grid_for(unit_cap, 256)), dim3(256), 0, s, w.unit_key, unit_cap, 0xFFFFFFFFu);
  launch_pdl(os_emit_kernel, dim3((unsigned)T), dim3((unsigned)tile_rows), 0, s, table, (int)kpad, (const int32_t*)w.sorted_rows,
             n_rows, (const uint32_t*)w.tile_union, (const int4*)w.tile_info, (const int32_t*)w.split_idx,
             (const int32_t*)num_out, pass_cap, unit_cap, (int64_t)T, split_tiles_out, out_row_out, pass_k_out,
             pass_idx_out, w.units, w.unit_key, w.unit_id);
  cb = w.cub_bytes;
  FT3D_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, (const uint32_t*)w.unit_key, w.unit_key_sorted,
                                            (const int32_t*)w.unit_id, w.unit_order, (int)unit_cap, 0, 9, s));
  launch_pdl(os_assign_kernel, dim3(1), dim3(256), 0, s, (const int32_t*)w.units, (const int32_t*)w.unit_order, num_out,
             unit_cap, clusters, units_out);
  return check_launch("ft3d_conv_os_plan");
}

}  // extern "C"
