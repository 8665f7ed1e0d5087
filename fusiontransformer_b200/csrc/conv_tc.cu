// Sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// forward / dgrad / transposed conv (a8-a10):  out[j,:] = sum_k in[nbr[j,k],:] @ B_k
//   One CTA owns a tile of 128 output rows and keeps its [128 x ncols] fp32 accumulator in TMEM across all
//   kernel offsets, so every output row is written exactly once: no scatter, no atomics, deterministic.
//   Per (offset k, 64-wide reduction block) pipeline stage:
//     * 4 producer warps gather the 128 neighbour rows (fp32 -> bf16) into a 128B-swizzled K-major A block,
//     * 1 loader warp streams the pre-packed bf16 weight block B_k with one cp.async.bulk (TMA unit),
//     * 1 issuer warp (one elected lane) issues tcgen05.mma 128 x ncols x 16 and commits to the stage's
//       "empty" mbarrier, releasing the slot to the producers.
//   Offsets for which no row of the tile has a neighbour are skipped by all roles (27-bit tile mask).
//   After the last commit the producer warps become the epilogue: tcgen05.ld -> registers -> global.
//
// wgrad (a9):  gW[k] += sum over pairs p of offset k of  a[pa(p),:]^T  b[pb(p),:]
//   One CTA owns (offset k, a slice of that offset's pair list, a 128-wide block of Cin).  The gathered rows are
//   the reduction dimension, so both operands are MN-major views of the same swizzled blocks; the
//   [128 x Cout] accumulator lives in TMEM and is added to gW with vector red.global.add.
#include "common.cuh"
#include "tc_common.cuh"

namespace ft3d {
using namespace tc;

constexpr int kProducerThreads = 128;  // warps 0-3: gather, later epilogue (TMEM lane quadrant = warp id)
constexpr int kLoaderWarp = 4;         // weight-block bulk loads
constexpr int kIssuerWarp = 5;         // tcgen05.mma
constexpr int kConvThreads = 192;
constexpr int kMaxStages = 6;

static int tmem_cols_for(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// ------------------------------------------------------------------------------------------------ weight packing
// image[(k*nkb + kb)][n][64] bf16, 128B-swizzled rows; element (n, r) of B_k is W[k][r][n] (forward) or W[k][n][r]
// (w_transposed: dgrad).  One thread per 16-byte chunk.
__global__ void pack_weights_kernel(const float* __restrict__ w, int K, int cin, int cout, int w_transposed,
                                    uint4* __restrict__ img) {
  const int red = w_transposed ? cout : cin;
  const int ncols = w_transposed ? cin : cout;
  const int nkb = (red + 63) / 64;
  const int64_t total = (int64_t)K * nkb * ncols * 8;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int pc = (int)(t & 7);               // physical chunk inside the 128-byte row
    int64_t rowid = t >> 3;
    int n = (int)(rowid % ncols);
    int64_t blk = rowid / ncols;
    int kb = (int)(blk % nkb);
    int k = (int)(blk / nkb);
    int c = pc ^ (n & 7);                // logical chunk stored at this physical position
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int r = kb * 64 + c * 8 + e;
      float x = 0.f;
      if (r < red) x = w_transposed ? __ldg(w + ((int64_t)k * cin + n) * cout + r) : __ldg(w + ((int64_t)k * cin + r) * cout + n);
      v[e] = x;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    img[t] = o;
  }
}

// All weight images of a model in ONE launch (they are re-packed after every optimizer step): `desc` lists, per
// image, the fp32 source, the destination and the first 16-byte chunk it owns in the launch's flat chunk index.
struct PackDesc {
  const float* w;
  uint4* img;
  int32_t K, cin, cout, w_transposed;
  int64_t chunk_begin;
};

__global__ void pack_weights_multi_kernel(const PackDesc* __restrict__ desc, int n_desc, int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_desc - 1;             // last descriptor with chunk_begin <= t
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(&desc[mid].chunk_begin) <= t) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = desc[lo];
    const int64_t u = t - d.chunk_begin;
    const int red = d.w_transposed ? d.cout : d.cin;
    const int ncols = d.w_transposed ? d.cin : d.cout;
    const int nkb = (red + 63) / 64;
    const int pc = (int)(u & 7);
    const int64_t rowid = u >> 3;
    const int n = (int)(rowid % ncols);
    const int64_t blk = rowid / ncols;
    const int kb = (int)(blk % nkb);
    const int k = (int)(blk / nkb);
    const int c = pc ^ (n & 7);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int r = kb * 64 + c * 8 + e;
      float x = 0.f;
      if (r < red)
        x = d.w_transposed ? __ldg(d.w + ((int64_t)k * d.cin + n) * d.cout + r)
                           : __ldg(d.w + ((int64_t)k * d.cin + r) * d.cout + n);
      v[e] = x;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    d.img[u] = o;
  }
}

// ------------------------------------------------------------------------------------------------ forward / dgrad
struct ConvSmemHeader {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t accum_full;
  uint32_t tmem_base;
  uint32_t kmask;
};

__global__ void __launch_bounds__(kConvThreads, 2)
conv_gather_tc_kernel(const float* __restrict__ in, const int32_t* __restrict__ nbr, int64_t n_out, int K, int kpad,
                      int kflip, int red, int ncols, const uint8_t* __restrict__ wpacked, float* __restrict__ out,
                      int nstages, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16 KB + B ncols*128)] 1024-aligned, then the nbr tile, then the header
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_bytes = ncols * kBlockRowBytes;
  const int stage_bytes = kBlockBytes + b_bytes;
  int32_t* s_nbr = (int32_t*)(smem + (size_t)nstages * stage_bytes);
  ConvSmemHeader* hdr = (ConvSmemHeader*)(s_nbr + kTileRows * kpad);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTileRows;
  const int nkb = (red + 63) / 64;

  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&hdr->full[s], kProducerThreads + 1);
      mbar_init(&hdr->empty[s], 1);
    }
    mbar_init(&hdr->accum_full, 1);
    hdr->kmask = 0;
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)tmem_cols);
  // neighbour tile -> smem (coalesced); rows past n_out read as "no neighbour"
  {
    const int total = kTileRows * kpad;
    const int64_t valid = (n_out - row0 < kTileRows ? n_out - row0 : kTileRows) * kpad;
    const int32_t* src = nbr + row0 * kpad;
    for (int t = tid; t < total; t += kConvThreads) s_nbr[t] = (t < valid) ? __ldg(src + t) : -1;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < kProducerThreads) {
    uint32_t m = 0;
    for (int k = 0; k < K; ++k) {
      unsigned b = __ballot_sync(0xffffffffu, s_nbr[tid * kpad + k] >= 0);
      if (b) m |= 1u << k;
    }
    if (lane == 0 && m) atomicOr(&hdr->kmask, m);
  }
  __syncthreads();
  const uint32_t kmask = hdr->kmask;
  const uint32_t tmem_base = hdr->tmem_base;
  const int nact = __popc(kmask);
  const int niter = nact * nkb;

  if (warp < 4) {
    // ------------------------------------------------ producers
    uint32_t rem = kmask;
    int it = 0;
    for (int a = 0; a < nact; ++a) {
      const int k = __ffs(rem) - 1;
      rem &= rem - 1;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % nstages;
        const uint32_t round = (uint32_t)(it / nstages);
        mbar_wait(&hdr->empty[s], (round & 1) ^ 1);
        uint8_t* a_blk = smem + (size_t)s * stage_bytes;
        const int width = red - kb * 64 < 64 ? red - kb * 64 : 64;
        fill_block_f32<kProducerThreads>(a_blk, in, red, kb * 64, width >> 3, tid,
                                         [&](int r) { return (int64_t)s_nbr[r * kpad + k]; });
        fence_proxy_async_smem();
        mbar_arrive(&hdr->full[s]);
      }
    }
    // ------------------------------------------------ epilogue: TMEM lane = tile row
    const int64_t row = row0 + tid;
    float* orow = out + row * ncols;
    if (niter > 0) {
      mbar_wait(&hdr->accum_full, 0);
      tc_fence_after();
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row < n_out) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    } else if (row < n_out) {
      for (int c0 = 0; c0 < ncols; c0 += 4) *reinterpret_cast<float4*>(orow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (warp == kLoaderWarp) {
    // ------------------------------------------------ weight loader (one lane drives the TMA unit)
    if (lane == 0) {
      uint32_t rem = kmask;
      int it = 0;
      for (int a = 0; a < nact; ++a) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1;
        const int kw = kflip ? (K - 1 - k) : k;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % nstages;
          const uint32_t round = (uint32_t)(it / nstages);
          mbar_wait(&hdr->empty[s], (round & 1) ^ 1);
          uint8_t* b_blk = smem + (size_t)s * stage_bytes + kBlockBytes;
          mbar_arrive_expect_tx(&hdr->full[s], (uint32_t)b_bytes);
          bulk_g2s(b_blk, wpacked + ((size_t)kw * nkb + kb) * b_bytes, (uint32_t)b_bytes, &hdr->full[s]);
        }
      }
    }
  } else {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const int nchunks = ncols > 256 ? 2 : 1;
      const int ncw = ncols / nchunks;  // 192 when ncols == 384
      const uint32_t idesc = umma_idesc_bf16(128, ncw, 0, 0);
      int it = 0;
      for (int a = 0; a < nact; ++a) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % nstages;
          const uint32_t round = (uint32_t)(it / nstages);
          mbar_wait(&hdr->full[s], round & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t b_addr = a_addr + kBlockBytes;
          const int ksteps = (red - kb * 64 < 64 ? red - kb * 64 : 64) >> 4;
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t da = smem_desc_sw128(a_addr + kk * 32, 16, 1024);
            for (int c = 0; c < nchunks; ++c) {
              const uint64_t db = smem_desc_sw128(b_addr + c * ncw * kBlockRowBytes + kk * 32, 16, 1024);
              umma_bf16(tmem_base + (uint32_t)(c * ncw), da, db, idesc, (it | kk) != 0);
            }
          }
          umma_commit(&hdr->empty[s]);
        }
      }
      if (niter > 0) umma_commit(&hdr->accum_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------ wgrad
constexpr int kWgradPairsPerCta = 1024;  // pairs reduced by one CTA (8 stages of 128)

struct WgradSmemHeader {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t accum_full;
  uint32_t tmem_base;
  int32_t pa[kMaxStages][kTileRows];
  int32_t pb[kMaxStages][kTileRows];
};

// work item -> (offset k, pair range) from the device-side prefix array
__device__ __forceinline__ bool wgrad_tc_item(const int32_t* __restrict__ off, int K, int item, int* k_out, int* begin,
                                              int* end) {
  int acc = 0;
  for (int k = 0; k < K; ++k) {
    int b = __ldg(off + k), e = __ldg(off + k + 1);
    int nc = (e - b + kWgradPairsPerCta - 1) / kWgradPairsPerCta;
    if (item < acc + nc) {
      *k_out = k;
      *begin = b + (item - acc) * kWgradPairsPerCta;
      *end = min(*begin + kWgradPairsPerCta, e);
      return true;
    }
    acc += nc;
  }
  return false;
}

__global__ void __launch_bounds__(kConvThreads, 2)
conv_wgrad_tc_kernel(const float* __restrict__ a, const float* __restrict__ b, const int2* __restrict__ pairs,
                     const int32_t* __restrict__ off, int K, int ca, int cin, int cout, float* __restrict__ gw,
                     int nstages, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nb_blocks = (cout + 63) / 64;                      // 64-channel column blocks of grad_out
  const int stage_bytes = (2 + nb_blocks) * kBlockBytes;       // A: two blocks (128 channels of Cin), B: cout
  WgradSmemHeader* hdr = (WgradSmemHeader*)(smem + (size_t)nstages * stage_bytes);

  int k, begin, end;
  if (!wgrad_tc_item(off, K, blockIdx.x, &k, &begin, &end)) return;   // uniform per CTA
  const int mb = blockIdx.y;                                    // 128-wide block of Cin
  const int m_valid = cin - mb * 128 < 128 ? cin - mb * 128 : 128;
  const int npairs = end - begin;
  const int niter = (npairs + kTileRows - 1) / kTileRows;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&hdr->full[s], kProducerThreads);
      mbar_init(&hdr->empty[s], 1);
    }
    mbar_init(&hdr->accum_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  if (warp < 4) {
    for (int it = 0; it < niter; ++it) {
      const int s = it % nstages;
      const uint32_t round = (uint32_t)(it / nstages);
      mbar_wait(&hdr->empty[s], (round & 1) ^ 1);
      // pair indices of this stage (rows past the slice end gather nothing => zero rows)
      {
        const int p = begin + it * kTileRows + tid;
        int ia = -1, ib = -1;
        if (p < end) {
          int2 pr = __ldg(pairs + p);
          ia = ca ? pr.y : pr.x;
          ib = ca ? pr.x : pr.y;
        }
        hdr->pa[s][tid] = ia;
        hdr->pb[s][tid] = ib;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // producers only
      uint8_t* st = smem + (size_t)s * stage_bytes;
      for (int blk = 0; blk * 64 < m_valid; ++blk) {
        const int width = m_valid - blk * 64 < 64 ? m_valid - blk * 64 : 64;
        fill_block_f32<kProducerThreads>(st + blk * kBlockBytes, a, cin, mb * 128 + blk * 64, width >> 3, tid,
                                         [&](int r) { return (int64_t)hdr->pa[s][r]; });
      }
      for (int blk = 0; blk < nb_blocks; ++blk) {
        const int width = cout - blk * 64 < 64 ? cout - blk * 64 : 64;
        fill_block_f32<kProducerThreads>(st + (2 + blk) * kBlockBytes, b, cout, blk * 64, width >> 3, tid,
                                         [&](int r) { return (int64_t)hdr->pb[s][r]; });
      }
      fence_proxy_async_smem();
      mbar_arrive(&hdr->full[s]);
    }
    // epilogue: TMEM lane = Cin index inside the block, columns = Cout
    mbar_wait(&hdr->accum_full, 0);
    tc_fence_after();
    float* grow = gw + ((int64_t)k * cin + mb * 128 + tid) * cout;
    for (int c0 = 0; c0 < cout; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (tid < m_valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          atomicAdd(reinterpret_cast<float4*>(grow + c0 + j),
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3])));
      }
    }
  } else if (warp == kIssuerWarp) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, cout, 1, 1);
      for (int it = 0; it < niter; ++it) {
        const int s = it % nstages;
        const uint32_t round = (uint32_t)(it / nstages);
        mbar_wait(&hdr->full[s], round & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + 2 * kBlockBytes;
        for (int kk = 0; kk < kTileRows / 16; ++kk) {   // 16 gathered rows per MMA
          const uint64_t da = smem_desc_sw128(a_addr + kk * 16 * kBlockRowBytes, kBlockBytes, 1024);
          const uint64_t db = smem_desc_sw128(b_addr + kk * 16 * kBlockRowBytes, kBlockBytes, 1024);
          umma_bf16(tmem_base, da, db, idesc, (it | kk) != 0);
        }
        umma_commit(&hdr->empty[s]);
      }
      umma_commit(&hdr->accum_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

static bool tc_shape_ok(int red, int ncols) {
  return red >= 16 && red <= 512 && red % 16 == 0 && ncols >= 32 && ncols <= 384 && ncols % 32 == 0 &&
         (ncols <= 256 || ncols == 384);
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

size_t ft3d_conv_packed_bytes(int32_t K, int32_t red, int32_t ncols) {
  if (K <= 0 || red <= 0 || ncols <= 0) return 0;
  return (size_t)K * ((red + 63) / 64) * ncols * tc::kBlockRowBytes;
}

int ft3d_conv_pack_weights(const float* w, int32_t K, int32_t cin, int32_t cout, int32_t w_transposed,
                           void* wpacked, ft3d_stream_t stream) {
  FT3D_REQUIRE(w && wpacked && K > 0 && cin > 0 && cout > 0, "ft3d_conv_pack_weights: bad arguments");
  const int red = w_transposed ? cout : cin, ncols = w_transposed ? cin : cout;
  int64_t total = (int64_t)K * ((red + 63) / 64) * ncols * 8;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, K, cin, cout, w_transposed,
                                                                              (uint4*)wpacked);
  return check_launch("ft3d_conv_pack_weights");
}

size_t ft3d_conv_pack_desc_bytes(void) { return sizeof(PackDesc); }

int ft3d_conv_pack_weights_multi(const void* desc, int32_t n_desc, int64_t total_chunks, ft3d_stream_t stream) {
  if (n_desc == 0 || total_chunks == 0) return FT3D_OK;
  FT3D_REQUIRE(desc && n_desc > 0 && total_chunks > 0, "ft3d_conv_pack_weights_multi: bad arguments");
  pack_weights_multi_kernel<<<grid_for(total_chunks, 256), 256, 0, (cudaStream_t)stream>>>((const PackDesc*)desc, n_desc,
                                                                                           total_chunks);
  return check_launch("ft3d_conv_pack_weights_multi");
}

int ft3d_conv_gather_tc(const float* in, const int32_t* nbr, int64_t n_out, int32_t K, int32_t kpad,
                        int32_t kflip, int32_t red, int32_t ncols, const void* wpacked, float* out,
                        ft3d_stream_t stream) {
  if (n_out == 0) return FT3D_OK;
  FT3D_REQUIRE(in && nbr && wpacked && out, "ft3d_conv_gather_tc: null pointer");
  FT3D_REQUIRE(K > 0 && K <= 32 && K <= kpad && kpad <= 32, "ft3d_conv_gather_tc: bad K=%d kpad=%d", K, kpad);
  FT3D_REQUIRE(tc_shape_ok(red, ncols), "ft3d_conv_gather_tc: unsupported shape red=%d ncols=%d", red, ncols);
  FT3D_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)wpacked & 15) == 0,
               "ft3d_conv_gather_tc: pointers must be 16-byte aligned");
  const int stage_bytes = tc::kBlockBytes + ncols * tc::kBlockRowBytes;
  const int tail = tc::kTileRows * kpad * 4 + (int)sizeof(ConvSmemHeader) + 1024;
  // as many stages as fit while leaving room for a second resident CTA when the tile is small
  int budget = (ncols <= 128) ? 110 * 1024 : 226 * 1024;
  int nstages = (budget - tail) / stage_bytes;
  if (nstages > kMaxStages) nstages = kMaxStages;
  FT3D_REQUIRE(nstages >= 2, "ft3d_conv_gather_tc: tile does not fit shared memory");
  const int smem_bytes = nstages * stage_bytes + tail;
  static int configured = 0;
  if (!configured) {
    FT3D_CUDA(cudaFuncSetAttribute(conv_gather_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 1;
  }
  const int64_t tiles = (n_out + tc::kTileRows - 1) / tc::kTileRows;
  conv_gather_tc_kernel<<<(unsigned)tiles, kConvThreads, smem_bytes, (cudaStream_t)stream>>>(
      in, nbr, n_out, K, kpad, kflip, red, ncols, (const uint8_t*)wpacked, out, nstages, tmem_cols_for(ncols));
  return check_launch("ft3d_conv_gather_tc");
}

int ft3d_conv_wgrad_tc(const float* a, const float* b, const int32_t* pairs, const int32_t* pair_offsets,
                       int32_t K, int32_t ca, int32_t cin, int32_t cout, int64_t max_pairs, float* gw,
                       ft3d_stream_t stream) {
  if (max_pairs == 0) return FT3D_OK;
  FT3D_REQUIRE(a && b && pairs && pair_offsets && gw && K > 0, "ft3d_conv_wgrad_tc: bad arguments");
  FT3D_REQUIRE(cin >= 16 && cin % 16 == 0 && cin <= 512 && cout >= 32 && cout % 32 == 0 && cout <= 256,
               "ft3d_conv_wgrad_tc: unsupported shape cin=%d cout=%d", cin, cout);
  FT3D_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)gw & 15) == 0,
               "ft3d_conv_wgrad_tc: pointers must be 16-byte aligned");
  const int nb_blocks = (cout + 63) / 64;
  const int stage_bytes = (2 + nb_blocks) * tc::kBlockBytes;
  const int tail = (int)sizeof(WgradSmemHeader) + 1024;
  int budget = (stage_bytes * 2 + tail <= 110 * 1024) ? 110 * 1024 : 226 * 1024;
  int nstages = (budget - tail) / stage_bytes;
  if (nstages > kMaxStages) nstages = kMaxStages;
  FT3D_REQUIRE(nstages >= 2, "ft3d_conv_wgrad_tc: tile does not fit shared memory");
  const int smem_bytes = nstages * stage_bytes + tail;
  static int configured = 0;
  if (!configured) {
    FT3D_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 1;
  }
  const int64_t items = (max_pairs + kWgradPairsPerCta - 1) / kWgradPairsPerCta + K;
  dim3 grid((unsigned)items, (unsigned)((cin + 127) / 128));
  conv_wgrad_tc_kernel<<<grid, kConvThreads, smem_bytes, (cudaStream_t)stream>>>(
      a, b, (const int2*)pairs, pair_offsets, K, ca, cin, cout, gw, nstages, tmem_cols_for(cout));
  return check_launch("ft3d_conv_wgrad_tc");
}

}  // extern "C"
