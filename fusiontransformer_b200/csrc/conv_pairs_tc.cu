// Pair-major sparse convolution on tcgen05 tensor cores: gather -> GEMM -> sorted, atomic-free scatter.
//
// LiDAR kernel maps are sparse (2-8 of the 27 offsets occupied per voxel), so the work is organised by the
// reference's own pair lists: for offset k the L_k (in,out) pairs form dense 128-row GEMM tiles
//     P[p, :] = in[pairs[p].gather, :] @ B_k          (bf16 operands, fp32 accumulation in TMEM)
// and every pair's partial row lands at its position p in the offset-major pair order.  A second, purely
// HBM-bound pass sums the (at most K) partial rows of each output voxel in fixed offset order through the
// pair-position table ppos[row][k] -- a sorted segmented reduction: deterministic, no atomics, each output row
// written once.  MMA work and operand traffic scale with the number of pairs, not with 27 x voxels.
//
// Stage pipeline of one CTA (= one tile of 128 pairs of one offset; red/64 stages):
//   warps 0-3  issue cp.async 16-byte gathers of bf16 feature rows straight into the 128B-swizzled A block and
//              arm the stage's mbarrier with cp.async.mbarrier.arrive.noinc (no register staging, every stage of
//              the tile in flight at once); afterwards they are the epilogue (tcgen05.ld -> P rows);
//   warp 4     streams the pre-packed weight block with cp.async.bulk (TMA unit) onto the same mbarrier;
//   warp 5     one lane issues tcgen05.mma 128 x ncols x 16 and commits.
// The same kernel runs dense GEMMs (k = 1 convolutions) with an identity gather (pairs == nullptr).
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"
#include "bn_common.cuh"

namespace ft3d {
using namespace tc;

constexpr int kPThreads = 192;
constexpr int kPProducers = 128;
constexpr int kPMaxStages = 8;

struct PairsSmemHeader {
  uint64_t full[kPMaxStages];
  uint64_t empty[kPMaxStages];
  uint64_t accum_full;
  uint32_t tmem_base;
  int32_t off[40];          // the map's K+1 prefix offsets
  int32_t idx[kTileRows];
};

__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// tile -> (offset k, [begin,end) in the pair list); tiles never straddle offsets.  `off` is the CTA's shared-memory
// copy of the K+1 prefix offsets (one L2 round trip for the whole CTA instead of one per scanned offset).
__device__ __forceinline__ bool pair_tile(const int32_t* off, int K, int item, int rows, int* k_out,
                                          int* begin, int* end) {
  int acc = 0;
  for (int k = 0; k < K; ++k) {
    int b = off[k], e = off[k + 1];
    int nt = (e - b + rows - 1) / rows;
    if (item < acc + nt) {
      *k_out = k;
      *begin = b + (item - acc) * rows;
      *end = min(*begin + rows, e);
      return true;
    }
    acc += nt;
  }
  return false;
}

// Walks consecutive tiles of a pair list in O(1) per step (pair_tile() scans the K offsets: used once, to seek).
struct TileCursor {
  const int32_t* off;
  int K, k, begin, end;
  __device__ __forceinline__ void seek(const int32_t* off_, int K_, int item) {
    off = off_;
    K = K_;
    if (!pair_tile(off, K, item, kTileRows, &k, &begin, &end)) k = K, begin = end = 0;
  }
  __device__ __forceinline__ void next() {
    int nb = begin + kTileRows;
    if (k < K && nb < off[k + 1]) {
      begin = nb;
      end = min(nb + kTileRows, off[k + 1]);
      return;
    }
    do { ++k; } while (k < K && off[k + 1] <= off[k]);
    if (k >= K) {
      begin = end = 0;
      return;
    }
    begin = off[k];
    end = min(begin + kTileRows, off[k + 1]);
  }
};

// gather one [128 x nchunk*16B] swizzled block of bf16 rows with cp.async (zero-fill for idx < 0)
__device__ __forceinline__ void gather_block_bf16(uint8_t* block, const __nv_bfloat16* __restrict__ src, int row_width,
                                                  int col0, int nchunk, int tid, const int32_t* __restrict__ s_idx) {
  const uint32_t base = smem_u32(block);
  const int total = kTileRows * nchunk;
  const int sh = 31 - __clz(nchunk);
  const bool pow2 = (nchunk & (nchunk - 1)) == 0;          // 2, 4 or 8 chunks for every channel width of the network
  for (int q = tid; q < total; q += kPProducers) {
    const int r = pow2 ? (q >> sh) : (q / nchunk);
    const int c = q - r * nchunk;
    const int g = s_idx[r];
    const __nv_bfloat16* p = src + (g >= 0 ? (int64_t)g * row_width + col0 + c * 8 : 0);
    cp_async_16(base + r * kBlockRowBytes + ((c ^ (r & 7)) << 4), p, g >= 0 ? 16u : 0u);
  }
}

__global__ void __launch_bounds__(kPThreads)
conv_pairs_tc_kernel(const __nv_bfloat16* __restrict__ in, const int2* __restrict__ pairs,
                     const int32_t* __restrict__ off, int K, int gather_col, int64_t n_identity, int red, int ncols,
                     const uint8_t* __restrict__ wpacked, float* __restrict__ P, int nstages, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nkb = (red + 63) / 64;
  const int b_bytes = ncols * kBlockRowBytes;
  const int stage_bytes = kBlockBytes + b_bytes;
  const int staging_bytes = kTileRows * (ncols + 4) * (int)sizeof(float);
  const int ring_bytes = nstages * stage_bytes > staging_bytes ? nstages * stage_bytes : staging_bytes;
  PairsSmemHeader* hdr = (PairsSmemHeader*)(smem + (size_t)((ring_bytes + 127) & ~127));   // ring: nstages <= nkb

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&hdr->full[s], kPProducers + 1);
      mbar_init(&hdr->empty[s], 1);
    }
    mbar_init(&hdr->accum_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();       // launch, barrier and TMEM set-up ran under the predecessor's tail; every global read is below
  int k = 0, begin = 0, end = 0;
  bool live;
  if (pairs != nullptr) {
    if (tid <= K) hdr->off[tid] = __ldg(off + tid);
    __syncthreads();
    live = pair_tile(hdr->off, K, blockIdx.x, kTileRows, &k, &begin, &end);          // uniform per CTA
  } else {
    begin = blockIdx.x * kTileRows;
    end = (int)min((int64_t)begin + kTileRows, n_identity);
    live = begin < end;
  }
  if (live && tid < kPProducers) {
    const int p = begin + tid;
    int g = -1;
    if (p < end) {
      if (pairs != nullptr) {
        int2 pr = __ldg(pairs + p);
        g = gather_col ? pr.y : pr.x;
      } else {
        g = p;
      }
    }
    hdr->idx[tid] = g;
  }
  __syncthreads();
  pdl_trigger();
  const uint32_t tmem_base = hdr->tmem_base;

  if (!live) {
    // no tile for this CTA (the grid is sized from an upper bound of the pair count)
  } else
  if (warp < 4) {
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % nstages;
      if (kb >= nstages) mbar_wait(&hdr->empty[s], ((uint32_t)(kb / nstages) & 1) ^ 1);
      const int width = red - kb * 64 < 64 ? red - kb * 64 : 64;
      gather_block_bf16(smem + (size_t)s * stage_bytes, in, red, kb * 64, width >> 3, tid, hdr->idx);
      cp_async_arrive_noinc(&hdr->full[s]);
    }
    // epilogue: TMEM lane = pair row of the tile.  Each thread stages its row in shared memory (the operand stages
    // are free once accum_full has fired; rows padded by 16 B so that the 8 lanes of a quarter-warp hit 8 different
    // bank groups) and hands it to the TMA unit as ONE bulk store: the P row is ncols*4 contiguous bytes in global
    // memory, written as full sectors instead of 32 scattered 16-byte pieces per warp instruction.
    mbar_wait(&hdr->accum_full, 0);
    tc_fence_after();
    const int p = begin + tid;
    const int row_floats = ncols + 4;
    float* srow = reinterpret_cast<float*>(smem) + (size_t)tid * row_floats;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    if (p < end) {
      fence_proxy_async_smem();            // this thread's st.shared -> visible to the bulk-copy (async proxy) read
      bulk_s2g(P + (int64_t)p * ncols, srow, (uint32_t)ncols * 4u);
      bulk_commit_group();
      bulk_wait_group_read0();             // shared memory must stay intact until the copy engine has read it
    }
  } else if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % nstages;
        if (kb >= nstages) mbar_wait(&hdr->empty[s], ((uint32_t)(kb / nstages) & 1) ^ 1);
        mbar_arrive_expect_tx(&hdr->full[s], (uint32_t)b_bytes);
        bulk_g2s(smem + (size_t)s * stage_bytes + kBlockBytes, wpacked + ((size_t)k * nkb + kb) * b_bytes,
                 (uint32_t)b_bytes, &hdr->full[s]);
      }
    }
  } else {
    if (lane == 0) {
      const int nchunks = ncols > 256 ? 2 : 1;
      const int ncw = ncols / nchunks;
      const uint32_t idesc = umma_idesc_bf16(128, ncw, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % nstages;
        mbar_wait(&hdr->full[s], (uint32_t)(kb / nstages) & 1);
        fence_proxy_async_smem();          // cp.async (generic proxy) data -> tensor-core (async proxy) reads
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + kBlockBytes;
        const int ksteps = (red - kb * 64 < 64 ? red - kb * 64 : 64) >> 4;
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t da = smem_desc_sw128(a_addr + kk * 32, 16, 1024);
          for (int c = 0; c < nchunks; ++c) {
            const uint64_t db = smem_desc_sw128(b_addr + c * ncw * kBlockRowBytes + kk * 32, 16, 1024);
            umma_bf16(tmem_base + (uint32_t)(c * ncw), da, db, idesc, (kb | kk) != 0);
          }
        }
        if (kb + nstages < nkb) umma_commit(&hdr->empty[s]);
      }
      umma_commit(&hdr->accum_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------ persistent variant
// Same arithmetic, organised for reuse and overlap: each CTA owns a contiguous run of tiles (so consecutive tiles
// mostly share their offset k), keeps B_k resident in shared memory until k changes, streams the gathered A blocks
// through a ring that runs across tile boundaries, and double-buffers the TMEM accumulator so that the gathers and
// MMAs of tile t+1 proceed under the epilogue of tile t.  Warp roles (320 threads):
//   0-3  gather producers   cp.async rows -> A ring slot -> full_a[slot]
//   4    weight loader      cp.async.bulk B_k (all k-blocks) -> b_full, after b_free of the previous offset
//   5    MMA issuer         tcgen05.mma into accumulator buffer (tile & 1); commits empty_a[slot], acc_full[buf], b_free
//   6-9  epilogue           tcgen05.ld -> per-warp padded staging -> coalesced 128-byte row segments -> acc_empty[buf]
constexpr int kV3Threads = 320;
constexpr int kV3MaxSlots = 8;
constexpr int kV3StageFloats = 32 * 36;        // one warp: 32 rows x (32 + 4 pad) floats

struct V3Header {
  uint64_t full_a[kV3MaxSlots];
  uint64_t empty_a[kV3MaxSlots];
  uint64_t b_full, b_free;
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int32_t off[40];
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kV3Threads)
conv_pairs_tc_v3_kernel(const __nv_bfloat16* __restrict__ in, const int2* __restrict__ pairs,
                        const int32_t* __restrict__ off, int K, int gather_col, int64_t n_identity, int red, int ncols,
                        const uint8_t* __restrict__ wpacked, float* __restrict__ P, int nslots, int tcols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nkb = (red + 63) / 64;
  const int b_bytes = ncols * kBlockRowBytes;
  uint8_t* b_region = smem;
  uint8_t* a_ring = smem + (size_t)nkb * b_bytes;
  float* staging = reinterpret_cast<float*>(a_ring + (size_t)nslots * kBlockBytes);
  V3Header* hdr = reinterpret_cast<V3Header*>(reinterpret_cast<uint8_t*>(staging) + 4 * kV3StageFloats * sizeof(float));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&hdr->full_a[s], kPProducers);
      mbar_init(&hdr->empty_a[s], 1);
    }
    mbar_init(&hdr->b_full, 1);
    mbar_init(&hdr->b_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hdr->acc_full[b], 1);
      mbar_init(&hdr->acc_empty[b], kPProducers);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)(2 * tcols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();       // launch, barrier and TMEM set-up ran under the predecessor's tail; every global read is below
  // ---- schedule: this CTA's run of tiles
  int T;
  if (pairs != nullptr) {
    if (tid <= K) hdr->off[tid] = __ldg(off + tid);
    __syncthreads();
    T = 0;
    for (int k = 0; k < K; ++k) T += (hdr->off[k + 1] - hdr->off[k] + kTileRows - 1) / kTileRows;
  } else {
    T = (int)((n_identity + kTileRows - 1) / kTileRows);
    if (tid == 0) hdr->off[0] = 0, hdr->off[1] = (int32_t)n_identity;     // identity gather = one offset of n rows
    K = 1;
    __syncthreads();
  }
  pdl_trigger();
  const int chunk = (T + (int)gridDim.x - 1) / (int)gridDim.x;
  const int g0 = (int)blockIdx.x * chunk;
  const int g1 = g0 + chunk < T ? g0 + chunk : T;          // g0 >= g1: no tiles for this CTA (uniform)
  const uint32_t tmem_base = hdr->tmem_base;

  if (g0 >= g1) {
  } else
  if (warp < 4) {
    // ------------------------------------------------------------------ gather producers
    // Thread t owns 16-byte chunk c = t & 7 of the tile rows r_j = (t >> 3) + 16 j, j = 0..7; its 8 gather indices live
    // in registers and are requested one tile ahead (see the wgrad kernel below: the shared-memory index staging of
    // the first version cost a dependent shared-memory load per copy).
    const int c = tid & 7, r0 = tid >> 3;
    const uint32_t dst_t = (uint32_t)(r0 * kBlockRowBytes) + (uint32_t)((c ^ (r0 & 7)) << 4);
    auto load_rows = [&](const TileCursor& tc_, bool live, int (&gi)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = tc_.begin + r0 + 16 * j;
        gi[j] = -1;
        if (live && p < tc_.end) {
          if (pairs != nullptr) {
            const int2 pr = __ldg(pairs + p);
            gi[j] = gather_col ? pr.y : pr.x;
          } else {
            gi[j] = p;
          }
        }
      }
    };
    uint32_t cnt = 0;
    TileCursor cur;
    cur.seek(hdr->off, K, g0);
    int gi[8], gn[8];
    load_rows(cur, true, gi);
    const uint32_t ring0 = smem_u32(a_ring) + dst_t;
    const __nv_bfloat16* src0 = in + c * 8;
    for (int g = g0; g < g1; ++g) {
      cur.next();
      load_rows(cur, g + 1 < g1, gn);
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb, ++cnt) {
        const int slot = (int)(cnt % (uint32_t)nslots);
        const uint32_t use = cnt / (uint32_t)nslots;
        if (use > 0) mbar_wait(&hdr->empty_a[slot], (use & 1) ^ 1);
        if (c * 8 < red - kb * 64) {
          const uint32_t dst = ring0 + (uint32_t)slot * (uint32_t)kBlockBytes;
          const __nv_bfloat16* src = src0 + kb * 64;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            cp_async_16(dst + (uint32_t)(16 * j * kBlockRowBytes), src + (gi[j] >= 0 ? (int64_t)gi[j] * red : 0),
                        gi[j] >= 0 ? 16u : 0u);
        }
        cp_async_arrive_noinc(&hdr->full_a[slot]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) gi[j] = gn[j];
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      int cur_k = -1;
      uint32_t nb = 0;
      TileCursor cur;
      cur.seek(hdr->off, K, g0);
      for (int g = g0; g < g1; ++g, cur.next()) {
        const int k = cur.k;
        if (k == cur_k) continue;
        if (nb > 0) mbar_wait(&hdr->b_free, (nb - 1) & 1);     // every MMA that read the previous B_k has completed
        mbar_arrive_expect_tx(&hdr->b_full, (uint32_t)(nkb * b_bytes));
        for (int kb = 0; kb < nkb; ++kb)
          bulk_g2s(b_region + (size_t)kb * b_bytes, wpacked + ((size_t)k * nkb + kb) * b_bytes, (uint32_t)b_bytes,
                   &hdr->b_full);
        cur_k = k;
        ++nb;
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    {
      const uint32_t idesc = umma_idesc_bf16(128, ncols, 0, 0);
      const uint64_t dhi = smem_desc_sw128(0, 16, 1024);
      const uint32_t a_ring0 = smem_u32(a_ring), b_region0 = smem_u32(b_region);
      int cur_k = -1;
      uint32_t nb = 0, cnt = 0;
      TileCursor cur;
      cur.seek(hdr->off, K, g0);
      int k = cur.k;
      for (int g = g0; g < g1; ++g) {
        const int it = g - g0, buf = it & 1;
        const uint32_t ub = (uint32_t)it >> 1;
        if (ub > 0) mbar_wait(&hdr->acc_empty[buf], (ub & 1) ^ 1);   // epilogue has drained this accumulator
        if (k != cur_k) {
          mbar_wait(&hdr->b_full, nb & 1);
          ++nb;
          cur_k = k;
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * tcols);
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int slot = (int)(cnt % (uint32_t)nslots);
          const uint32_t use = cnt / (uint32_t)nslots;
          mbar_wait(&hdr->full_a[slot], use & 1);
          fence_proxy_async_smem();
          tc_fence_after();
          const uint64_t da0 = dhi | (uint64_t)(((a_ring0 + (uint32_t)slot * (uint32_t)kBlockBytes) >> 4) & 0x3FFF);
          const uint64_t db0 = dhi | (uint64_t)(((b_region0 + (uint32_t)kb * (uint32_t)b_bytes) >> 4) & 0x3FFF);
          const int ksteps = (red - kb * 64 < 64 ? red - kb * 64 : 64) >> 4;
          if (elect_one_sync()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              if (kk < ksteps) umma_bf16(tmem_d, da0 + 2 * kk, db0 + 2 * kk, idesc, (kb | kk) != 0);
            umma_commit(&hdr->empty_a[slot]);
          }
          __syncwarp();
        }
        int nk = k;
        bool last_of_k = false;
        if (g + 1 < g1) {
          cur.next();
          nk = cur.k;
          last_of_k = nk != k;
        }
        if (elect_one_sync()) {
          umma_commit(&hdr->acc_full[buf]);
          if (last_of_k) umma_commit(&hdr->b_free);
        }
        __syncwarp();
        k = nk;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM quadrant = warp % 4)
    const int q = warp & 3;
    float* st = staging + (size_t)(warp - 6) * kV3StageFloats;
    TileCursor cur;
    cur.seek(hdr->off, K, g0);
    for (int g = g0; g < g1; ++g, cur.next()) {
      const int begin = cur.begin, end = cur.end;
      const int it = g - g0, buf = it & 1;
      mbar_wait(&hdr->acc_full[buf], ((uint32_t)it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * tcols) + ((uint32_t)(q * 32) << 16);
      const int row0 = begin + q * 32;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(st + lane * 36 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        __syncwarp();
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {                   // a warp store = 4 rows x 128 contiguous bytes
          const int r = i8 * 4 + (lane >> 3);
          const int p = row0 + r;
          if (p < end)
            *reinterpret_cast<float4*>(P + (int64_t)p * ncols + c0 + (lane & 7) * 4) =
                *reinterpret_cast<const float4*>(st + r * 36 + (lane & 7) * 4);
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&hdr->acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)(2 * tcols));
}

// ------------------------------------------------------------------------------------------------ sorted scatter
// out[row,:] = sum of P[ppos[row,j],:] over the row's pair positions (stored compacted in ascending offset order and
// terminated by -1).  2-D block: threadIdx.x = 4 channels, threadIdx.y = row lane; each CTA owns a contiguous chunk
// of rows.  With STATS the per-channel sum / sum of squares of the rows just written are folded per CTA and a tiny
// second launch turns them into the BatchNorm statistics of this layer (bn_common.cuh): the activation is not re-read.
template <bool STATS, int RB>
__global__ void __launch_bounds__(kColThreads)
conv_reduce_kernel(const float* __restrict__ P, const int32_t* __restrict__ ppos, int64_t n_rows, int kpad, int ncols,
                   int rows_per_cta, float* __restrict__ out, float* __restrict__ partials,
                   const float* __restrict__ pivot, const int32_t* __restrict__ valid_rows) {
  pdl_enter();
  __shared__ float4 s_stage[STATS ? kColStageFloat4 : 1];
  const int ch = threadIdx.x * 4;
  const float4 pv = STATS ? stat_pivot(pivot, ch) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t n_stat = STATS ? effective_rows(n_rows, valid_rows) : 0;    // padding rows stay out of the shifted sums
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row1 = row0 + rows_per_cta < n_rows ? row0 + rows_per_cta : n_rows;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const int nq = kpad >> 2;
  // A thread walks its rows four at a time: the four position quads are requested together, then all sixteen
  // candidate partial rows (absent slots re-read partial row 0 and are discarded), so that ~16 independent 16-byte
  // loads are in flight per thread instead of a serial  position -> partial -> next row  chain.  The additions keep
  // the ascending-offset order, so the result is bit-identical to the one-row-at-a-time loop.
  const int64_t rstep = blockDim.y;
  for (int64_t rbase = row0 + threadIdx.y; rbase < row1; rbase += RB * rstep) {
    int4 pp[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t row = rbase + i * rstep;
      pp[i] = row < row1 ? __ldg(reinterpret_cast<const int4*>(ppos + row * kpad)) : make_int4(-1, -1, -1, -1);
    }
    float4 v[RB][4];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      v[i][0] = __ldg(reinterpret_cast<const float4*>(P + (int64_t)max(pp[i].x, 0) * ncols + ch));
      v[i][1] = __ldg(reinterpret_cast<const float4*>(P + (int64_t)max(pp[i].y, 0) * ncols + ch));
      v[i][2] = __ldg(reinterpret_cast<const float4*>(P + (int64_t)max(pp[i].z, 0) * ncols + ch));
      v[i][3] = __ldg(reinterpret_cast<const float4*>(P + (int64_t)max(pp[i].w, 0) * ncols + ch));
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t row = rbase + i * rstep;
      if (row >= row1) break;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pp[i].x >= 0) add4(acc, v[i][0]);
      if (pp[i].y >= 0) add4(acc, v[i][1]);
      if (pp[i].z >= 0) add4(acc, v[i][2]);
      if (pp[i].w >= 0) {
        add4(acc, v[i][3]);
        const int4* pr = reinterpret_cast<const int4*>(ppos + row * kpad);
        for (int q = 1; q < nq; ++q) {                     // rows with more than four pairs (deep levels)
          const int4 p4 = __ldg(pr + q);
          if (p4.x < 0) break;
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(P + (int64_t)p4.x * ncols + ch));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(P + (int64_t)max(p4.y, 0) * ncols + ch));
          const float4 a2 = __ldg(reinterpret_cast<const float4*>(P + (int64_t)max(p4.z, 0) * ncols + ch));
          const float4 a3 = __ldg(reinterpret_cast<const float4*>(P + (int64_t)max(p4.w, 0) * ncols + ch));
          add4(acc, a0);
          if (p4.y < 0) break;
          add4(acc, a1);
          if (p4.z < 0) break;
          add4(acc, a2);
          if (p4.w < 0) break;
          add4(acc, a3);
        }
      }
      *reinterpret_cast<float4*>(out + row * ncols + ch) = acc;
      if (STATS && row < n_stat) stat_add(s1, s2, acc, pv);
    }
  }
  if (STATS) col_publish(s1, s2, partials, ncols, s_stage);
}

// pposT[gather-side row][k] = pair position, for the role-swapped use of a map (dgrad, transposed conv)
__global__ void pair_positions_kernel(const int2* __restrict__ pairs, const int32_t* __restrict__ off, int K, int kpad,
                                      int col, int32_t* __restrict__ ppos) {
  pdl_enter();
  __shared__ int32_t s_off[40];
  if ((int)threadIdx.x <= K) s_off[threadIdx.x] = __ldg(off + threadIdx.x);
  __syncthreads();
  const int total = s_off[K];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    int k = 0;
    while (k + 1 < K && p >= s_off[k + 1]) ++k;
    const int2 pr = __ldg(pairs + p);
    ppos[(int64_t)(col ? pr.y : pr.x) * kpad + k] = p;
  }
}

// in-place row compaction: valid positions to the front (ascending offset order kept), -1 behind
__global__ void compact_rows_kernel(int32_t* ppos, int64_t n_rows, int kpad) {
  pdl_enter();
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < n_rows; row += (int64_t)gridDim.x * blockDim.x) {
    int4* pr = reinterpret_cast<int4*>(ppos + row * kpad);
    int v[32];
    const int nq = kpad >> 2;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int4 t = q < nq ? pr[q] : make_int4(-1, -1, -1, -1);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    int32_t* dst = ppos + row * kpad;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (v[k] >= 0) dst[cnt++] = v[k];
    for (int j = cnt; j < kpad; ++j) dst[j] = -1;
  }
}

__global__ void fill_i32_kernel_p(int32_t* p, int64_t n, int32_t v) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread
__global__ void to_bf16_kernel(const float* __restrict__ src, int64_t n, __nv_bfloat16* __restrict__ dst) {
  pdl_enter();
  const int64_t n8 = n >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n8 << 3; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
}

// ------------------------------------------------------------------------------------------------ wgrad (bf16 inputs)
// The unit line [0, U) of the persistent wgrad kernel (unit u = Cin block u / T, tile u % T) is cut at the CTA run
// boundaries (multiples of `chunk`) and at the starts of the non-empty (Cin block, offset) groups; the pieces between
// cuts are the SEGMENTS, each flushed exactly once.  Index of the segment that starts at unit `pos` = number of
// distinct cut positions in (0, pos].  `off` = the K+1 pair prefix offsets (shared-memory copy).
__device__ __forceinline__ int wg_segment_index(const int32_t* off, int K, int T, int MB, int chunk, int pos) {
  int n = pos / chunk;
  for (int mb = 0; mb < MB; ++mb) {
    int tp = 0;
    for (int k = 0; k < K; ++k) {
      const int nt = (off[k + 1] - off[k] + kTileRows - 1) / kTileRows;
      if (nt > 0) {
        const int p = mb * T + tp;
        if (p > 0 && p <= pos && p % chunk != 0) ++n;      // a group start on a CTA boundary is the same cut
      }
      tp += nt;
    }
  }
  return n;
}

// Second stage of the deterministic persistent weight gradient: gw[k][Cin block rows][:] (+)= sum of the group's
// segments in unit order.  blockIdx.x = (Cin block, offset), blockIdx.y = row chunk, threadIdx.x = 4 columns.
__global__ void __launch_bounds__(kColThreads)
conv_wgrad_fold2_kernel(const float* __restrict__ partial, const int32_t* __restrict__ off_g, int K, int64_t n_identity,
                        int wgrad_grid, int MB, int cin, int cout, int accumulate, float* __restrict__ gw) {
  pdl_enter();
  __shared__ int32_t s_off[40];
  if (off_g != nullptr) {
    if ((int)(threadIdx.y * blockDim.x + threadIdx.x) <= K) s_off[threadIdx.y * blockDim.x + threadIdx.x] =
        __ldg(off_g + threadIdx.y * blockDim.x + threadIdx.x);
  } else if (threadIdx.x == 0 && threadIdx.y == 0) {
    s_off[0] = 0;
    s_off[1] = (int32_t)n_identity;
  }
  __syncthreads();
  const int mb = blockIdx.x / K, k = blockIdx.x % K;
  int T = 0, tp = 0, nt = 0;
  for (int j = 0; j < K; ++j) {
    const int n = (s_off[j + 1] - s_off[j] + kTileRows - 1) / kTileRows;
    if (j < k) tp += n;
    if (j == k) nt = n;
    T += n;
  }
  const int m_valid = cin - mb * 128 < 128 ? cin - mb * 128 : 128;
  const int ch = threadIdx.x * 4;
  int first = 0, count = 0;
  if (nt > 0) {
    const int U = T * MB;
    const int chunk = (U + wgrad_grid - 1) / wgrad_grid;
    const int gs = mb * T + tp, ge = gs + nt;
    first = wg_segment_index(s_off, K, T, MB, chunk, gs);
    count = 1 + (ge - 1) / chunk - gs / chunk;
  }
  const size_t sstride = (size_t)kTileRows * cout;
  for (int r = blockIdx.y * blockDim.y + threadIdx.y; r < m_valid; r += gridDim.y * blockDim.y) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = partial + ((size_t)first * kTileRows + r) * cout + ch;
    int i = 0;
    for (; i + 4 <= count; i += 4) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(p + (size_t)i * sstride));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(i + 1) * sstride));
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(i + 2) * sstride));
      const float4 a3 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(i + 3) * sstride));
      add4(acc, a0); add4(acc, a1); add4(acc, a2); add4(acc, a3);
    }
    for (; i < count; ++i) add4(acc, __ldg(reinterpret_cast<const float4*>(p + (size_t)i * sstride)));
    float4* dst = reinterpret_cast<float4*>(gw + ((int64_t)k * cin + mb * 128 + r) * cout + ch);
    if (accumulate) {
      const float4 old = *dst;
      acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
    }
    *dst = acc;
  }
}

constexpr int kWg2MaxSlots = 6;

struct Wg2Header {
  uint64_t full[kWg2MaxSlots];
  uint64_t empty[kWg2MaxSlots];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int32_t off[40];
};

__global__ void __launch_bounds__(kV3Threads)
conv_wgrad_pairs_tc_v2_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                              const int2* __restrict__ pairs, const int32_t* __restrict__ off, int K, int ca,
                              int64_t n_identity, int cin, int cout, float* __restrict__ gw, int nslots, int tcols,
                              int a_blocks, float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nb_blocks = (cout + 63) / 64;
  const int stage_bytes = (a_blocks + nb_blocks) * kBlockBytes;
  float* staging = reinterpret_cast<float*>(smem + (size_t)nslots * stage_bytes);
  Wg2Header* hdr = reinterpret_cast<Wg2Header*>(reinterpret_cast<uint8_t*>(staging) + 4 * kV3StageFloats * sizeof(float));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&hdr->full[s], kPProducers);
      mbar_init(&hdr->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->acc_full[i], 1);
      mbar_init(&hdr->acc_empty[i], kPProducers);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)(2 * tcols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();       // set-up ran under the predecessor's tail; every global read is below
  int T;
  if (pairs != nullptr) {
    if (tid <= K) hdr->off[tid] = __ldg(off + tid);
    __syncthreads();
    T = 0;
    for (int k = 0; k < K; ++k) T += (hdr->off[k + 1] - hdr->off[k] + kTileRows - 1) / kTileRows;
  } else {
    T = (int)((n_identity + kTileRows - 1) / kTileRows);
    if (tid == 0) hdr->off[0] = 0, hdr->off[1] = (int32_t)n_identity;
    K = 1;
    __syncthreads();
  }
  pdl_trigger();
  const int MB = (cin + 127) / 128;
  const int U = T * MB;
  const int chunk = (U + (int)gridDim.x - 1) / (int)gridDim.x;
  const int u0 = (int)blockIdx.x * chunk;
  const int u1 = u0 + chunk < U ? u0 + chunk : U;          // u0 >= u1: nothing for this CTA (uniform)
  const uint32_t tmem_base = hdr->tmem_base;

  // unit u = (Cin block u / T, tile u % T); walked with an O(1) cursor
  struct UnitCursor {
    TileCursor c;
    int mb, t, T;
    __device__ __forceinline__ void seek(const int32_t* off_, int K_, int T_, int u) {
      T = T_;
      mb = u / T;
      t = u - mb * T;
      c.seek(off_, K_, t);
    }
    __device__ __forceinline__ void next() {
      if (++t == T) {
        t = 0;
        ++mb;
        c.seek(c.off, c.K, 0);
      } else {
        c.next();
      }
    }
  };

  if (u0 >= u1) {
  } else
  if (warp < 4) {
    // ------------------------------------------------------------------ gather producers
    // Thread t owns 16-byte chunk c = t & 7 of the tile rows r_j = (t >> 3) + 16 j, j = 0..7, in EVERY block of a stage
    // (8 lanes = one 128-byte row segment, as before).  The 8 pair entries of those rows are read straight from the
    // pair list into registers -- requested one unit ahead, so the L2 round trip of the indices hides behind the
    // previous unit's gathers -- and a block is 8 cp.async per thread: one 64-bit multiply-add and the copy.  (The
    // first version staged the tile's indices in shared memory and walked a runtime-length loop: shared-memory load ->
    // address -> copy, ~180 cycles per copy and thread, 11 B/cycle/SM -- ten times slower than the tensor pipe.)
    const int c = tid & 7, r0 = tid >> 3;
    const uint32_t dst_t = (uint32_t)(r0 * kBlockRowBytes) + (uint32_t)((c ^ (r0 & 7)) << 4);   // (r0 + 16 j) & 7 == r0 & 7
    auto load_rows = [&](const UnitCursor& cur, bool live, int (&ia)[8], int (&ib)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = cur.c.begin + r0 + 16 * j;
        ia[j] = ib[j] = -1;
        if (live && p < cur.c.end) {
          if (pairs != nullptr) {
            const int2 pr = __ldg(pairs + p);
            ia[j] = ca ? pr.y : pr.x;
            ib[j] = ca ? pr.x : pr.y;
          } else {
            ia[j] = ib[j] = p;
          }
        }
      }
    };
    UnitCursor uc;
    uc.seek(hdr->off, K, T, u0);
    int ia[8], ib[8], na[8], nb[8];
    load_rows(uc, true, ia, ib);
    for (int u = u0; u < u1; ++u) {
      const int mb = uc.mb;
      uc.next();
      load_rows(uc, u + 1 < u1, na, nb);                   // next unit's rows: in flight under this unit's gathers
      const int it = u - u0;
      const int slot = it % nslots;
      const uint32_t use = (uint32_t)(it / nslots);
      if (use > 0) mbar_wait(&hdr->empty[slot], (use & 1) ^ 1);
      const uint32_t st = smem_u32(smem + (size_t)slot * stage_bytes) + dst_t;
      const int m_valid = cin - mb * 128 < 128 ? cin - mb * 128 : 128;
      for (int blk = 0; blk * 64 < m_valid; ++blk) {
        const int width = m_valid - blk * 64 < 64 ? m_valid - blk * 64 : 64;
        if (c * 8 < width) {
          const __nv_bfloat16* src = a + mb * 128 + blk * 64 + c * 8;
          const uint32_t dst = st + (uint32_t)(blk * kBlockBytes);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            cp_async_16(dst + (uint32_t)(16 * j * kBlockRowBytes), src + (ia[j] >= 0 ? (int64_t)ia[j] * cin : 0),
                        ia[j] >= 0 ? 16u : 0u);
        }
      }
      for (int blk = 0; blk < nb_blocks; ++blk) {
        const int width = cout - blk * 64 < 64 ? cout - blk * 64 : 64;
        if (c * 8 < width) {
          const __nv_bfloat16* src = b + blk * 64 + c * 8;
          const uint32_t dst = st + (uint32_t)((a_blocks + blk) * kBlockBytes);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            cp_async_16(dst + (uint32_t)(16 * j * kBlockRowBytes), src + (ib[j] >= 0 ? (int64_t)ib[j] * cout : 0),
                        ib[j] >= 0 ? 16u : 0u);
        }
      }
      cp_async_arrive_noinc(&hdr->full[slot]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ia[j] = na[j], ib[j] = nb[j];
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the units (uniform control flow keeps descriptors and barrier addresses in uniform
    // registers) and one elected lane issues; from inside `if (lane == 0)` every tcgen05 instruction is wrapped in an
    // elect / broadcast loop of ~30 dependent instructions (measured 165 ns per MMA in conv_os.cu).
    {
      const uint32_t idesc = umma_idesc_bf16(128, cout, 1, 1);
      const uint64_t dhi = smem_desc_sw128(0, kBlockBytes, 1024);
      const uint32_t smem0 = smem_u32(smem);
      int group = -1, pk = -1, pmb = -1;
      UnitCursor uc;
      uc.seek(hdr->off, K, T, u0);
      for (int u = u0; u < u1; ++u, uc.next()) {
        const int mb = uc.mb, k = uc.c.k;
        const bool fresh = (k != pk || mb != pmb);
        if (fresh) {
          if (group >= 0 && elect_one_sync()) umma_commit(&hdr->acc_full[group & 1]);   // previous group -> flush warps
          __syncwarp();
          ++group;
          const uint32_t ub = (uint32_t)group >> 1;
          if (ub > 0) mbar_wait(&hdr->acc_empty[group & 1], (ub & 1) ^ 1);
          pk = k;
          pmb = mb;
        }
        const int it = u - u0;
        const int slot = it % nslots;
        mbar_wait(&hdr->full[slot], (uint32_t)(it / nslots) & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)((group & 1) * tcols);
        const uint32_t a_addr = smem0 + (uint32_t)slot * (uint32_t)stage_bytes;
        const uint64_t da0 = dhi | (uint64_t)((a_addr >> 4) & 0x3FFF);
        const uint64_t db0 = dhi | (uint64_t)(((a_addr + (uint32_t)(a_blocks * kBlockBytes)) >> 4) & 0x3FFF);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kTileRows / 16; ++kk)       // 16 pairs = 16 rows of 128 bytes = 128 descriptor units
            umma_bf16(tmem_d, da0 + (uint64_t)(kk * 16 * kBlockRowBytes >> 4), db0 + (uint64_t)(kk * 16 * kBlockRowBytes >> 4),
                      idesc, (!fresh || kk != 0) ? 1u : 0u);
          umma_commit(&hdr->empty[slot]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&hdr->acc_full[group & 1]);
      __syncwarp();
    }
  } else if (warp >= 6) {
    // ------------------------------------------------------------------ flush
    const int q = warp & 3;
    float* st = staging + (size_t)(warp - 6) * kV3StageFloats;
    int group = -1, pk = -1, pmb = -1;
    const int seg0 = partial != nullptr ? wg_segment_index(hdr->off, K, T, MB, chunk, u0) : 0;
    UnitCursor uc;
    uc.seek(hdr->off, K, T, u0);
    for (int u = u0; u <= u1; ++u) {
      int mb = -1, k = -1;
      if (u < u1) {
        mb = uc.mb;
        k = uc.c.k;
        uc.next();
      }
      if (k == pk && mb == pmb) continue;
      if (group >= 0) {                                    // group (pk, pmb) is complete
        const int buf = group & 1;
        mbar_wait(&hdr->acc_full[buf], ((uint32_t)group >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(buf * tcols) + ((uint32_t)(q * 32) << 16);
        const int row0 = pmb * 128 + q * 32;
        // deterministic mode: this segment (= the part of an (offset, Cin block) group that lies in this CTA's run) owns
        // slot seg0 + group of `partial`; conv_wgrad_fold2_kernel adds the segments of a group in order
        float* pslot = partial != nullptr ? partial + ((size_t)(seg0 + group) * kTileRows + q * 32) * cout : nullptr;
        for (int c0 = 0; c0 < cout; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(st + lane * 36 + j) = make_float4(
                __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          __syncwarp();
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = i8 * 4 + (lane >> 3);
            const int row = row0 + r;
            const float4 val = *reinterpret_cast<const float4*>(st + r * 36 + (lane & 7) * 4);
            if (pslot != nullptr)
              *reinterpret_cast<float4*>(pslot + (size_t)r * cout + c0 + (lane & 7) * 4) = val;
            else if (row < cin)
              atomicAdd(reinterpret_cast<float4*>(gw + ((int64_t)pk * cin + row) * cout + c0 + (lane & 7) * 4), val);
          }
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(&hdr->acc_empty[buf]);
      }
      ++group;
      pk = k;
      pmb = mb;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)(2 * tcols));
}

static int tmem_cols_pow2(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

int ft3d_to_bf16(const float* src, int64_t n, void* dst, ft3d_stream_t stream) {
  if (n == 0) return FT3D_OK;
  FT3D_REQUIRE(src && dst && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "ft3d_to_bf16: bad arguments");
  launch_pdl(to_bf16_kernel, dim3(grid_for((n >> 3) + 1, 256)), dim3(256), 0, (cudaStream_t)stream, src, n, (__nv_bfloat16*)dst);
  return check_launch("ft3d_to_bf16");
}

int ft3d_kmap_pair_positions(const int32_t* pairs, const int32_t* pair_offsets, int32_t K, int32_t kpad, int32_t col,
                             int64_t n_rows, int64_t max_pairs, int32_t* ppos_out, ft3d_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FT3D_REQUIRE(K > 0 && K <= kpad && (col == 0 || col == 1), "ft3d_kmap_pair_positions: bad arguments");
  if (n_rows > 0) {
    FT3D_REQUIRE(ppos_out != nullptr, "ft3d_kmap_pair_positions: null output");
    launch_pdl(fill_i32_kernel_p, dim3(grid_for(n_rows * kpad, 256)), dim3(256), 0, s, ppos_out, n_rows * kpad, -1);
  }
  if (n_rows > 0 && max_pairs > 0) {
    FT3D_REQUIRE(pairs && pair_offsets, "ft3d_kmap_pair_positions: null input");
    launch_pdl(pair_positions_kernel, dim3(grid_for(max_pairs, 256)), dim3(256), 0, s, (const int2*)pairs, pair_offsets, K, kpad, col,
                                                                   ppos_out);
    launch_pdl(compact_rows_kernel, dim3(grid_for(n_rows, 128)), dim3(128), 0, s, ppos_out, n_rows, kpad);
  }
  return check_launch("ft3d_kmap_pair_positions");
}

int ft3d_conv_pairs_tc(const void* in_bf16, const int32_t* pairs, const int32_t* pair_offsets, int32_t K,
                       int32_t gather_col, int64_t max_pairs, int32_t red, int32_t ncols, const void* wpacked,
                       float* partial_out, ft3d_stream_t stream) {
  if (max_pairs == 0) return FT3D_OK;
  FT3D_REQUIRE(in_bf16 && wpacked && partial_out && K > 0 && K <= 32, "ft3d_conv_pairs_tc: bad arguments");
  FT3D_REQUIRE((pairs == nullptr) == (pair_offsets == nullptr), "ft3d_conv_pairs_tc: pairs and pair_offsets go together");
  FT3D_REQUIRE(pairs != nullptr || K == 1, "ft3d_conv_pairs_tc: identity gather needs K == 1");
  FT3D_REQUIRE(red >= 16 && red % 16 == 0 && red <= 512 && ncols >= 32 && ncols % 32 == 0 &&
                   (ncols <= 256 || ncols == 384),
               "ft3d_conv_pairs_tc: unsupported shape red=%d ncols=%d", red, ncols);
  FT3D_REQUIRE(((uintptr_t)in_bf16 & 15) == 0 && ((uintptr_t)partial_out & 15) == 0 && ((uintptr_t)wpacked & 15) == 0,
               "ft3d_conv_pairs_tc: pointers must be 16-byte aligned");
  const int nkb = (red + 63) / 64;
  // ---- persistent weight-stationary variant whenever B_k and >= 2 ring slots fit (everything but 384-wide operands)
  {
    const int b_total = nkb * ncols * tc::kBlockRowBytes;
    const int fixed = b_total + 4 * kV3StageFloats * (int)sizeof(float) + (int)sizeof(V3Header) + 1024;
    int nslots = (226 * 1024 - fixed) / tc::kBlockBytes;
    static int use_v3 = -1;
    if (use_v3 < 0) {
      const char* e = getenv("FT3D_PAIRS_PERSISTENT");
      use_v3 = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    if (use_v3 && ncols <= 256 && nslots >= 2) {
      const int want = 2 * nkb < 3 ? 3 : 2 * nkb;            // two tiles of A blocks in flight
      if (nslots > want) nslots = want;
      if (nslots > kV3MaxSlots) nslots = kV3MaxSlots;
      // prefer two resident CTAs when that costs no ring depth below one full tile
      if (fixed + nslots * tc::kBlockBytes > 113 * 1024 && fixed + nkb * tc::kBlockBytes <= 113 * 1024 && nkb >= 2)
        nslots = (113 * 1024 - fixed) / tc::kBlockBytes;
      const int smem_bytes = fixed + nslots * tc::kBlockBytes;
      int ctas_per_sm = (227 * 1024) / smem_bytes;                 // shared memory
      const int by_tmem = 512 / (2 * tmem_cols_pow2(ncols));        // two accumulator buffers per CTA
      if (ctas_per_sm > by_tmem) ctas_per_sm = by_tmem;
      if (ctas_per_sm > 4) ctas_per_sm = 4;
      if (ctas_per_sm < 1) ctas_per_sm = 1;
      static int configured3 = 0;
      if (!configured3) {
        FT3D_CUDA(cudaFuncSetAttribute(conv_pairs_tc_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured3 = 1;
      }
      const int64_t tiles = (max_pairs + tc::kTileRows - 1) / tc::kTileRows + (pairs ? K : 0);
      const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
      const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
      launch_pdl(conv_pairs_tc_v3_kernel, dim3(grid), dim3(kV3Threads), smem_bytes, (cudaStream_t)stream, (const __nv_bfloat16*)in_bf16, (const int2*)pairs, pair_offsets, K, gather_col, max_pairs, red, ncols,
          (const uint8_t*)wpacked, partial_out, nslots, tmem_cols_pow2(ncols));
      return check_launch("ft3d_conv_pairs_tc");
    }
  }
  const int stage_bytes = tc::kBlockBytes + ncols * tc::kBlockRowBytes;
  const int tail = (int)sizeof(PairsSmemHeader) + 1024 + 128;
  int nstages = (226 * 1024 - tail) / stage_bytes;
  if (nstages > nkb) nstages = nkb;
  FT3D_REQUIRE(nstages >= 1, "ft3d_conv_pairs_tc: red=%d ncols=%d does not fit shared memory", red, ncols);
  const int staging = tc::kTileRows * (ncols + 4) * (int)sizeof(float);     // epilogue rows reuse the operand stages
  const int ring = nstages * stage_bytes > staging ? nstages * stage_bytes : staging;
  FT3D_REQUIRE(ring + tail <= 227 * 1024, "ft3d_conv_pairs_tc: ncols=%d does not fit shared memory", ncols);
  const int smem_bytes = ring + tail;
  static int configured = 0;
  if (!configured) {
    FT3D_CUDA(cudaFuncSetAttribute(conv_pairs_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 1;
  }
  const int64_t tiles = (max_pairs + tc::kTileRows - 1) / tc::kTileRows + (pairs ? K : 0);
  launch_pdl(conv_pairs_tc_kernel, dim3((unsigned)tiles), dim3(kPThreads), smem_bytes, (cudaStream_t)stream, (const __nv_bfloat16*)in_bf16, (const int2*)pairs, pair_offsets, K, gather_col, max_pairs, red, ncols,
      (const uint8_t*)wpacked, partial_out, nstages, tmem_cols_pow2(ncols));
  return check_launch("ft3d_conv_pairs_tc");
}

static int launch_reduce(const float* partial, const int32_t* ppos, int64_t n_rows, int32_t kpad, int32_t ncols,
                         float* out, bool stats, float eps, float momentum, float* stat, float* running_mean,
                         float* running_var, const int32_t* valid_rows, void* workspace, size_t workspace_bytes,
                         cudaStream_t s, const char* what) {
  // `partial` must hold at least one row even when the map has no pair: absent slots re-read row 0 and discard it
  FT3D_REQUIRE(partial && ppos && out && (kpad == 8 || kpad == 16 || kpad == 32) && ncols >= 4 && ncols % 4 == 0 &&
                   ncols <= 1024,
               "%s: bad arguments", what);
  FT3D_REQUIRE(((uintptr_t)partial & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)ppos & 15) == 0,
               "%s: pointers must be 16-byte aligned", what);
  ColGrid g = col_grid(n_rows, ncols / 4);
  if (stats) {
    FT3D_REQUIRE(stat && workspace && ((uintptr_t)workspace & 15) == 0 &&
                     workspace_bytes >= col_workspace_bytes(ncols) && (running_mean == nullptr) == (running_var == nullptr),
                 "%s: statistics need stat and a workspace of ft3d_bn_workspace(ncols) bytes", what);
    switch (row_batch()) {
      case 1: launch_pdl(conv_reduce_kernel<true, 1>, dim3(g.grid), dim3(g.block), 0, s, partial, ppos, n_rows, kpad, ncols, g.rows_per_cta, out,
                                                        (float*)workspace, (const float*)running_mean, valid_rows); break;
      case 2: launch_pdl(conv_reduce_kernel<true, 2>, dim3(g.grid), dim3(g.block), 0, s, partial, ppos, n_rows, kpad, ncols, g.rows_per_cta, out,
                                                        (float*)workspace, (const float*)running_mean, valid_rows); break;
      default: launch_pdl(conv_reduce_kernel<true, 4>, dim3(g.grid), dim3(g.block), 0, s, partial, ppos, n_rows, kpad, ncols, g.rows_per_cta, out,
                                                        (float*)workspace, (const float*)running_mean, valid_rows); break;
    }
    launch_pdl(col_finalize_kernel<0>, dim3(ncols / 4), dim3(kColThreads), 0, s, (const float*)workspace, g.grid, ncols, n_rows, eps, momentum,
                                                             stat, running_mean, running_var, 0, valid_rows);
  } else {
    switch (row_batch()) {
      case 1: launch_pdl(conv_reduce_kernel<false, 1>, dim3(g.grid), dim3(g.block), 0, s, partial, ppos, n_rows, kpad, ncols, g.rows_per_cta, out, nullptr, nullptr, nullptr); break;
      case 2: launch_pdl(conv_reduce_kernel<false, 2>, dim3(g.grid), dim3(g.block), 0, s, partial, ppos, n_rows, kpad, ncols, g.rows_per_cta, out, nullptr, nullptr, nullptr); break;
      default: launch_pdl(conv_reduce_kernel<false, 4>, dim3(g.grid), dim3(g.block), 0, s, partial, ppos, n_rows, kpad, ncols, g.rows_per_cta, out, nullptr, nullptr, nullptr); break;
    }
  }
  return check_launch(what);
}

int ft3d_conv_reduce(const float* partial, const int32_t* ppos, int64_t n_rows, int32_t kpad, int32_t ncols,
                     float* out, ft3d_stream_t stream) {
  if (n_rows == 0) return FT3D_OK;
  return launch_reduce(partial, ppos, n_rows, kpad, ncols, out, false, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr,
                       nullptr, 0, (cudaStream_t)stream, "ft3d_conv_reduce");
}

int ft3d_conv_reduce_bn(const float* partial, const int32_t* ppos, int64_t n_rows, int32_t kpad, int32_t ncols,
                        float* out, float eps, float momentum, float* stat, float* running_mean, float* running_var,
                        const int32_t* valid_rows, void* workspace, size_t workspace_bytes, ft3d_stream_t stream) {
  FT3D_REQUIRE(n_rows > 0, "ft3d_conv_reduce_bn: BatchNorm statistics need at least one row");
  return launch_reduce(partial, ppos, n_rows, kpad, ncols, out, true, eps, momentum, stat, running_mean, running_var,
                       valid_rows, workspace, workspace_bytes, (cudaStream_t)stream, "ft3d_conv_reduce_bn");
}

// launch geometry of the persistent wgrad kernel (shared by the atomic and the deterministic entry points)
struct Wg2Launch {
  int a_blocks, nslots, tcols, smem_bytes;
  unsigned grid;
};

static int wg2_launch(int cin, int cout, int K, int64_t max_pairs, bool has_pairs, Wg2Launch* out) {
  const int nb_blocks = (cout + 63) / 64;
  out->a_blocks = cin <= 64 ? 1 : 2;
  const int stage = (out->a_blocks + nb_blocks) * tc::kBlockBytes;
  const int fixed = 4 * kV3StageFloats * (int)sizeof(float) + (int)sizeof(Wg2Header) + 1024;
  out->tcols = tmem_cols_pow2(cout);
  // two resident CTAs (3-deep rings) when shared memory and TMEM (2 accumulators each) allow, else one CTA
  int ctas_per_sm = 1, nslots = (226 * 1024 - fixed) / stage;
  if (fixed + 3 * stage <= 113 * 1024 && 4 * out->tcols <= 512) {
    ctas_per_sm = 2;
    nslots = (113 * 1024 - fixed) / stage;
  }
  if (nslots > kWg2MaxSlots) nslots = kWg2MaxSlots;
  if (nslots < 2) return 1;
  out->nslots = nslots;
  out->smem_bytes = fixed + nslots * stage;
  const int64_t tiles = (max_pairs + tc::kTileRows - 1) / tc::kTileRows + (has_pairs ? K : 0);
  const int64_t units = tiles * ((cin + 127) / 128);
  const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  out->grid = (unsigned)(units < cap ? units : cap);
  return 0;
}

size_t ft3d_conv_wgrad_det_workspace(int32_t K, int32_t cin, int32_t cout, int64_t max_pairs, int32_t has_pairs) {
  Wg2Launch l;
  if (wg2_launch(cin, cout, K, max_pairs, has_pairs != 0, &l)) return 0;
  const size_t slots = (size_t)l.grid + (size_t)K * ((cin + 127) / 128) + 1;
  return align_up(slots * tc::kTileRows * (size_t)cout * sizeof(float), 256);
}

int ft3d_conv_wgrad_pairs_tc_det(const void* a_bf16, const void* b_bf16, const int32_t* pairs,
                                 const int32_t* pair_offsets, int32_t K, int32_t ca, int32_t cin, int32_t cout,
                                 int64_t max_pairs, float* gw, int32_t accumulate, void* workspace,
                                 size_t workspace_bytes, ft3d_stream_t stream) {
  if (max_pairs == 0) return FT3D_OK;
  FT3D_REQUIRE(a_bf16 && b_bf16 && gw && K > 0 && K <= 32 && workspace, "ft3d_conv_wgrad_pairs_tc_det: bad arguments");
  FT3D_REQUIRE((pairs == nullptr) == (pair_offsets == nullptr) && (pairs != nullptr || K == 1),
               "ft3d_conv_wgrad_pairs_tc_det: identity gather needs K == 1 and no offsets");
  FT3D_REQUIRE(cin >= 16 && cin % 16 == 0 && cin <= 512 && cout >= 32 && cout % 32 == 0 && cout <= 256,
               "ft3d_conv_wgrad_pairs_tc_det: unsupported shape cin=%d cout=%d", cin, cout);
  FT3D_REQUIRE(((uintptr_t)a_bf16 & 15) == 0 && ((uintptr_t)b_bf16 & 15) == 0 && ((uintptr_t)gw & 15) == 0 &&
                   ((uintptr_t)workspace & 255) == 0,
               "ft3d_conv_wgrad_pairs_tc_det: pointers must be 16-byte aligned (workspace 256)");
  FT3D_REQUIRE(workspace_bytes >= ft3d_conv_wgrad_det_workspace(K, cin, cout, max_pairs, pairs != nullptr),
               "ft3d_conv_wgrad_pairs_tc_det: workspace smaller than ft3d_conv_wgrad_det_workspace(...)");
  Wg2Launch l;
  FT3D_REQUIRE(wg2_launch(cin, cout, K, max_pairs, pairs != nullptr, &l) == 0,
               "ft3d_conv_wgrad_pairs_tc_det: tile does not fit shared memory");
  static int configured = 0;
  if (!configured) {
    FT3D_CUDA(cudaFuncSetAttribute(conv_wgrad_pairs_tc_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 1;
  }
  const int MB = (cin + 127) / 128;
  cudaStream_t s = (cudaStream_t)stream;
  launch_pdl(conv_wgrad_pairs_tc_v2_kernel, dim3(l.grid), dim3(kV3Threads), l.smem_bytes, s,
             (const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16, (const int2*)pairs, pair_offsets, K, ca,
             max_pairs, cin, cout, gw, l.nslots, l.tcols, l.a_blocks, (float*)workspace);
  const int cv = cout / 4;
  int ry = kColThreads / cv;
  if (ry < 1) ry = 1;
  if (ry > 32) ry = 32;
  const int ychunks = (128 + ry - 1) / ry < 4 ? (128 + ry - 1) / ry : 4;
  launch_pdl(conv_wgrad_fold2_kernel, dim3((unsigned)(K * MB), (unsigned)ychunks), dim3(cv, ry), 0, s,
             (const float*)workspace, pair_offsets, (int)K, max_pairs, (int)l.grid, MB, (int)cin, (int)cout,
             (int)accumulate, gw);
  return check_launch("ft3d_conv_wgrad_pairs_tc_det");
}

int ft3d_conv_wgrad_pairs_tc(const void* a_bf16, const void* b_bf16, const int32_t* pairs,
                             const int32_t* pair_offsets, int32_t K, int32_t ca, int32_t cin, int32_t cout,
                             int64_t max_pairs, float* gw, ft3d_stream_t stream) {
  if (max_pairs == 0) return FT3D_OK;
  FT3D_REQUIRE(a_bf16 && b_bf16 && gw && K > 0, "ft3d_conv_wgrad_pairs_tc: bad arguments");
  FT3D_REQUIRE((pairs == nullptr) == (pair_offsets == nullptr) && (pairs != nullptr || K == 1),
               "ft3d_conv_wgrad_pairs_tc: identity gather needs K == 1 and no offsets");
  FT3D_REQUIRE(cin >= 16 && cin % 16 == 0 && cin <= 512 && cout >= 32 && cout % 32 == 0 && cout <= 256,
               "ft3d_conv_wgrad_pairs_tc: unsupported shape cin=%d cout=%d", cin, cout);
  FT3D_REQUIRE(((uintptr_t)a_bf16 & 15) == 0 && ((uintptr_t)b_bf16 & 15) == 0 && ((uintptr_t)gw & 15) == 0,
               "ft3d_conv_wgrad_pairs_tc: pointers must be 16-byte aligned");
  Wg2Launch l;
  FT3D_REQUIRE(wg2_launch(cin, cout, K, max_pairs, pairs != nullptr, &l) == 0,
               "ft3d_conv_wgrad_pairs_tc: tile does not fit shared memory");
  static int configured2 = 0;
  if (!configured2) {
    FT3D_CUDA(cudaFuncSetAttribute(conv_wgrad_pairs_tc_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured2 = 1;
  }
  launch_pdl(conv_wgrad_pairs_tc_v2_kernel, dim3(l.grid), dim3(kV3Threads), l.smem_bytes, (cudaStream_t)stream,
             (const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16, (const int2*)pairs, pair_offsets, K, ca,
             max_pairs, cin, cout, gw, l.nslots, l.tcols, l.a_blocks, (float*)nullptr);
  return check_launch("ft3d_conv_wgrad_pairs_tc");
}

}  // extern "C"
