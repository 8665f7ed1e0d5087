// Output-stationary sparse convolution on tcgen05 tensor cores, with the BatchNorm statistics in the epilogue.
//
//     out[j,:] = sum_k  in[ nbr_k(j), : ] @ B_k          bf16 operands, fp32 accumulation in TMEM
//
// A CTA owns tiles of 128 output rows taken from the mask-sorted schedule of os_plan.cu: for every offset k that at
// least one row of the tile uses (a "pass") it gathers the 128 neighbour rows (absent neighbours are zero rows),
// multiplies them with B_k and ACCUMULATES IN TMEM across passes in ascending k -- the reference's own summation
// order.  A finished tile is read from TMEM once and every output row is written exactly once, complete: there is no
// partial-row buffer, no second reduction pass, no atomics, and the result does not depend on scheduling.
// While the rows pass through the epilogue their per-channel sum and sum of squares are folded per CTA, and the last
// CTA to finish turns the partials into this layer's BatchNorm training statistics (mean, rstd, running stats): the
// activation is never re-read for its statistics and no finalize kernel is launched.
//
// Warp roles ((NPW + 6) warps, one persistent CTA per SM, units dealt round by round; NPW = 4 or 8 gather warps):
//   0..NPW-1  TMA gather warps: each issues 32/NPW x cp.async.bulk.tensor.2d ... tile::gather4 per stage (4 rows x 128 B,
//             hardware 128B swizzle, row index -1 => zero fill); together one 128 x 64 A block;
//   NPW       weight loader: arms the stage's mbarrier with the byte count and issues the cp.async.bulk of B_k;
//             (LDGSTS mode, FT3D_OS_GATHER=ldgsts, NPW = 4: warps 0-3 issue 16-byte cp.async row pieces instead)
//   NPW+1     MMA issuer: tcgen05.mma 128 x ncols x 16 into accumulator buffer (unit & 1); tcgen05.commit frees ring
//             slots and hands the accumulator to the epilogue;
//   NPW+2..5  epilogue: tcgen05.ld -> per-warp staging transpose -> 128-byte row segments to out[row] + statistics.
// The A/B ring runs across pass and unit boundaries; two TMEM accumulators (ncols <= 256) let the gathers and MMAs
// of unit u+1 run under the epilogue of unit u.
//
// Every role is ONE warp walking the schedule serially, so the instructions a role spends per stage bound the kernel
// long before the tensor pipe or L2 do (a stage is only 4 MMAs of 16-128 cycles each).  The loops are therefore kept
// to a few dozen instructions per stage: ring slot / phase are counters (no division), the four k-steps of a stage
// are ONE asm block (descriptor low words advance by +2, predicates instead of branches), everything that only the
// timing experiments need (FT3D_OS_DEBUG, the trace stamps) is compiled into a separate DBG instantiation.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"
#include "bn_common.cuh"

namespace ft3d {
using namespace tc;

constexpr int kOsMaxSlots = 10;
constexpr int kOsDefaultGatherWarps = 8;       // TMA gather warps per CTA unless FT3D_OS_PRODUCERS says otherwise
constexpr int kOsStageFloats = 32 * 36;        // one epilogue warp: 32 rows x (32 + 4 pad) floats
constexpr int kOsProducers = 128;              // LDGSTS mode: warps 0-3
constexpr int kOsAhead = 4;                    // passes whose indices are in flight in a TMA gather warp

struct OsHeader {
  uint64_t full[kOsMaxSlots];
  uint64_t empty[kOsMaxSlots];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t pfull[kOsMaxSlots];                 // CTA pairs: rank 1's half of the stage has landed (remote arrive)
  uint32_t tmem_base;
};

struct OsArgs {
  const __nv_bfloat16* in;
  const int4* units;          // [U][2] int4 = {first pass, passes, tile, chunks, chunk, scratch base, -, -}
  const int32_t* num;         // {P, U, S, cap}
  const int32_t* out_row;
  const int32_t* pass_k;
  const int32_t* pass_idx;
  const uint8_t* wpacked;
  float* out;
  float* partials;            // statistics: [gridDim.x][2][ncols]; nullptr = no statistics
  float* scratch;             // [S][128][ncols] partial tiles of split tiles (folded by conv_os_fold_kernel)
  const int32_t* valid_rows;
  const float* pivot;         // statistics: per-channel shift of the sums (the running mean; bn_common.cuh), nullable
  unsigned long long* trace;  // nullable: per CTA 8 x u64 (globaltimer stamps and counts), tools/conv_os_probe.py
  int64_t n_out;
  int unit_cap, K, kflip, red, ncols, nslots, tcols, nbuf;
  int cs, tile_rows;          // CTAs per cluster (1, 2, 4) and rows of a schedule tile (128 * cs)
  int dbg;                    // FT3D_OS_DEBUG, timing experiments only (results are then meaningless): 1 = no MMAs,
                              // 2 = no A gathers, 4 = no B copies (FT3D_OS_DEBUG=.. tools/conv_os_probe.py);
                              // 16 = alternate the accumulator of consecutive MMAs (are dependent MMAs serialised?)
                              // 8 = launch conv_os_kernel only, no fold / finalize (bench.py times the kernel alone
                              // on repeats of a call whose complete form has already run)
};

__device__ __forceinline__ void os_cp_async_16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void os_cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void os_named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned long long os_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// TMA gather4: rows {r0..r3} x 64 bf16 starting at column `col` of the 2-D tensor -> 4 consecutive 128-byte rows of
// a 128B-swizzled block at dst (negative / out-of-range rows are filled with zeros; the transaction always counts
// 512 bytes on the mbarrier).
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const CUtensorMap* tmap, int col, int r0, int r1, int r2,
                                            int r3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst_smem),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

// ---- thread-block clusters: the CTAs of a cluster own the 128-row slices of ONE schedule tile of 128*cs rows and walk
// the same pass list in lockstep, so the weight block B_k of a pass is fetched from L2 once per cluster (rank 0 issues
// a multicast bulk copy that lands at the same shared-memory offset in every CTA and completes on every CTA's own
// mbarrier) instead of once per 128 rows.  B is 47-65 % of the kernel's L2 traffic for the 128-256-wide layers.
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// arrives on the mbarrier at the same offset in every CTA of cta_mask once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA pairs (cta_group::2): the two CTAs of a cluster issue ONE tcgen05.mma of M = 256 per k-step.  Each CTA keeps
// its own 128 gathered A rows and only HALF of the weight block (N/2 columns) in shared memory; the tensor cores of the
// pair read both halves, so a weight block enters each SM once per 256 output rows instead of once per 128 -- the
// multicast above saves L2 reads but not SM ingress, which is what bounds the wide layers.  Rank 0 issues the MMAs and
// commits to both CTAs' barriers; rank 1's MMA warp only relays "my half of stage s has landed" to rank 0.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {   // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {     // pairs with a remote arrive
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAITC_DONE;\n"
      "bra WAITC_LOOP;\n"
      "WAITC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// walks the passes of this CTA's work units u = first, first + step, ... (units without passes are skipped); the
// record of the following unit is requested one unit ahead so that crossing a unit boundary costs no L2 round trip
struct OsPassIter {
  const int4* units;
  int U, u, step, begin, n, q, nb, nn;
  __device__ __forceinline__ void fetch_next() {
    nb = nn = 0;
    if (u + step < U) {
      const int4 un = __ldg(units + 2 * (int64_t)(u + step));
      nb = un.x; nn = un.y;
    }
  }
  __device__ __forceinline__ void init(const int4* units_, int U_, int first, int step_) {
    units = units_; U = U_; u = first; step = step_; begin = n = q = nb = nn = 0;
    if (u < U) {
      const int4 un = __ldg(units + 2 * (int64_t)u);
      begin = un.x; n = un.y;
      fetch_next();
      while (u < U && n == 0) advance();
    }
  }
  __device__ __forceinline__ void advance() {
    u += step; begin = nb; n = nn; q = 0;
    if (u < U) fetch_next();
  }
  __device__ __forceinline__ bool valid() const { return u < U; }
  __device__ __forceinline__ int pass() const { return begin + q; }
  __device__ __forceinline__ void next() {
    if (++q >= n) {
      advance();
      while (u < U && n == 0) advance();
    }
  }
};

// ---- barrier operations on shared-window addresses (the role loops keep &full[0] / &empty[0] as 32-bit bases)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITA_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAITA_DONE;\n"
      "bra WAITA_LOOP;\n"
      "WAITA_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_a(uint32_t bar, uint32_t parity) {   // pairs with a remote arrive
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITCA_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAITCA_DONE;\n"
      "bra WAITCA_LOOP;\n"
      "WAITCA_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                     uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_gather4_a(uint32_t dst_smem, const CUtensorMap* tmap, int col, int r0, int r1, int r2,
                                              int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst_smem),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void umma_commit_a(uint32_t bar, int cs, uint16_t cmask) {
  if (PAIR)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cmask)
                 : "memory");
  else if (cs == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cmask)
                 : "memory");
}

// One stage = one 64-wide k-block = up to four k-steps of 16: D (+)= A[128 x 16] * B[n x 16]^T, k-step kk reads the
// operands 32 bytes (2 descriptor units) further along the swizzled rows.  a_lo / b_lo are the LOW words of the
// shared-memory descriptors (start >> 4 | LBO field), desc_hi the constant high word; `acc_first` = 0 overwrites the
// accumulator with the first k-step (first stage of a unit).  One straight-line block, the k-steps beyond `ksteps`
// predicated off: between two MMAs the issuing warp executes two adds and two moves on uniform registers.
#define FT3D_OS_KBLOCK_ASM(GROUP)                                                      \
  asm volatile(                                                                        \
      "{\n"                                                                            \
      ".reg .pred p0, pt, q0, q1, q2, q3;\n"                                           \
      ".reg .b64 da, db;\n"                                                            \
      ".reg .b32 al, bl;\n"                                                            \
      "setp.ne.b32 p0, %5, 0;\n"                                                       \
      "setp.ge.u32 pt, %6, 0;\n"                                                       \
      "setp.gt.u32 q0, %6, 0;\n"                                                       \
      "setp.gt.u32 q1, %6, 1;\n"                                                       \
      "setp.gt.u32 q2, %6, 2;\n"                                                       \
      "setp.gt.u32 q3, %6, 3;\n"                                                       \
      "mov.b64 da, {%1, %3};\n"                                                        \
      "mov.b64 db, {%2, %3};\n"                                                        \
      "@q0 tcgen05.mma.cta_group::" GROUP ".kind::f16 [%0], da, db, %4, p0;\n"         \
      "add.u32 al, %1, 2;\n"                                                           \
      "add.u32 bl, %2, 2;\n"                                                           \
      "mov.b64 da, {al, %3};\n"                                                        \
      "mov.b64 db, {bl, %3};\n"                                                        \
      "@q1 tcgen05.mma.cta_group::" GROUP ".kind::f16 [%0], da, db, %4, pt;\n"         \
      "add.u32 al, %1, 4;\n"                                                           \
      "add.u32 bl, %2, 4;\n"                                                           \
      "mov.b64 da, {al, %3};\n"                                                        \
      "mov.b64 db, {bl, %3};\n"                                                        \
      "@q2 tcgen05.mma.cta_group::" GROUP ".kind::f16 [%0], da, db, %4, pt;\n"         \
      "add.u32 al, %1, 6;\n"                                                           \
      "add.u32 bl, %2, 6;\n"                                                           \
      "mov.b64 da, {al, %3};\n"                                                        \
      "mov.b64 db, {bl, %3};\n"                                                        \
      "@q3 tcgen05.mma.cta_group::" GROUP ".kind::f16 [%0], da, db, %4, pt;\n"         \
      "}\n" ::"r"(tmem_d),                                                             \
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc_first), "r"(ksteps)      \
      : "memory")
template <bool PAIR>
__device__ __forceinline__ void umma_bf16_kblock(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                 uint32_t idesc, uint32_t acc_first, uint32_t ksteps) {
  if (PAIR) FT3D_OS_KBLOCK_ASM("2");
  else FT3D_OS_KBLOCK_ASM("1");
}
#undef FT3D_OS_KBLOCK_ASM

// ring position of a role: slot index and the number of completed trips round the ring (-> barrier parities)
struct OsRing {
  uint32_t slot, round, nslots;
  __device__ __forceinline__ void init(int n) { slot = 0; round = 0; nslots = (uint32_t)n; }
  __device__ __forceinline__ void advance() {
    if (++slot == nslots) { slot = 0; ++round; }
  }
};

template <bool TMA, bool PAIR, int NPW, bool DBG>
__global__ void __launch_bounds__((NPW + 6) * 32)
conv_os_kernel(const __grid_constant__ CUtensorMap tmap, const OsArgs a) {
  static_assert(TMA || !PAIR, "CTA pairs use the TMA gather");
  static_assert(NPW == 4 || NPW == 8, "4 or 8 gather warps");
  static_assert(TMA || NPW == 4, "LDGSTS mode has four producer warps");
  constexpr int kLoaderWarp = NPW, kMmaWarp = NPW + 1, kEpiWarp0 = NPW + 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int red = a.red, ncols = a.ncols, nslots = a.nslots;
  const int nkb = (red + 63) / 64;
  const int b_bytes = ncols * kBlockRowBytes;
  const int b_stage = PAIR ? b_bytes / 2 : b_bytes;                        // a CTA of a pair holds half the columns
  const int stage_bytes = kBlockBytes + b_stage;
  float* staging = reinterpret_cast<float*>(smem + (size_t)nslots * stage_bytes);
  float* sstat = staging + 4 * kOsStageFloats;                                   // [4 warps][2][ncols]
  OsHeader* hdr = reinterpret_cast<OsHeader*>(sstat + 4 * 2 * ncols);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = a.cs;
  const int rank = cs > 1 ? (int)cluster_rank() : 0;
  const int step = (int)gridDim.x / cs, first = (int)blockIdx.x / cs;      // units are dealt to clusters
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  const int row_base = rank * kTileRows;                                   // this CTA's slice of a schedule tile
  const int dbg = DBG ? a.dbg : 0;

  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&hdr->full[s], TMA ? 1 : kOsProducers + 1);
      mbar_init(&hdr->empty[s], PAIR ? 1u : (uint32_t)cs);                 // every MMA issuer of the cluster has consumed it
      mbar_init(&hdr->pfull[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hdr->acc_full[b], 1);
      mbar_init(&hdr->acc_empty[b], PAIR ? 256 : 128);                     // pair: both CTAs' epilogues free rank 0's MMA
    }
    mbar_fence_init();
  }
  if (PAIR) {
    __syncthreads();
    cluster_sync_all();             // both CTAs are resident and their barriers initialised before the paired allocation
    if (warp == 0) tmem_alloc_pair(&hdr->tmem_base, (uint32_t)(a.nbuf * a.tcols));
  } else if (warp == 0) {
    tmem_alloc(&hdr->tmem_base, (uint32_t)(a.nbuf * a.tcols));
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // peers' barriers are initialised before anything is multicast to them
  tc_fence_after();
  pdl_enter();      // launch, barrier and TMEM set-up ran under the predecessor's tail; every global read is below
  const uint32_t tmem_base = hdr->tmem_base;
  int U = __ldg(a.num + 1);
  if (U > a.unit_cap) U = a.unit_cap;
  unsigned long long* tr = (DBG && a.trace != nullptr) ? a.trace + (size_t)blockIdx.x * 8 : nullptr;
  if (DBG && tr != nullptr && tid == 0) tr[0] = os_now();
  // per-stage stamps of CTA 0 behind the per-CTA records: [stage][4] = {issued, landed, MMAs committed, -} for the first
  // 512 stages, then [unit][2] = {accumulator ready, rows written} from offset 2048 (tools/conv_os_probe.py --fine)
  unsigned long long* fine = (DBG && a.trace != nullptr && blockIdx.x == 0) ? a.trace + (size_t)kNumSMs * 8 : nullptr;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(&hdr->full[0]), empty0 = smem_u32(&hdr->empty[0]);

  if (TMA && warp < NPW) {
    // ------------------------------------------------------------------ gather warps: TMA gather4 (A)
    // A 128 x 64 block is 32 gather4 loads; a warp issues them one by one (the operands of a TMA instruction are
    // warp-uniform), so the NPW gather warps each take 32/NPW of them: lanes 0..G-1 of warp w hold the indices of rows
    // (w*G + lane)*4 .. +3.  The loads complete on the stage's mbarrier, which the weight loader arms with the byte
    // count (a transaction count may run negative until then: only the loader's arrival can complete the phase).
    // The gather indices of a pass are requested kOsAhead passes before they are needed: a narrow layer issues a pass
    // in ~0.2 us, far less than a loaded L2 round trip.
    constexpr int G = 32 / NPW;
    OsPassIter it;
    it.init(a.units, U, first, step);
    const int sub = lane & (G - 1);
    int4 rq[kOsAhead];
    int queued = 0;
#pragma unroll
    for (int d = 0; d < kOsAhead; ++d) {
      rq[d] = make_int4(-1, -1, -1, -1);
      if (it.valid()) {
        rq[d] = __ldg(reinterpret_cast<const int4*>(a.pass_idx + (int64_t)it.pass() * a.tile_rows + row_base) + warp * G + sub);
        it.next();
        ++queued;
      }
    }
    OsRing ring;
    ring.init(nslots);
    const uint32_t dst0 = smem0 + (uint32_t)(warp * G) * 512u;
    uint32_t cnt = 0, npass = 0;
    while (queued > 0) {
#pragma unroll
      for (int d = 0; d < kOsAhead; ++d) {
        if (queued == 0) break;
        const int4 r4 = rq[d];
        --queued;
        if (DBG) ++npass;
        if (it.valid()) {                                    // refill the slot just consumed: pass + kOsAhead
          rq[d] = __ldg(reinterpret_cast<const int4*>(a.pass_idx + (int64_t)it.pass() * a.tile_rows + row_base) + warp * G + sub);
          it.next();
          ++queued;
        }
        // every lane walks the loop (uniform control flow: the TMA operands are built in uniform registers) and one
        // elected lane issues; the row indices come from lanes 0..G-1
        int rr[G][4];
#pragma unroll
        for (int i = 0; i < G; ++i) {
          rr[i][0] = __shfl_sync(0xffffffffu, r4.x, i); rr[i][1] = __shfl_sync(0xffffffffu, r4.y, i);
          rr[i][2] = __shfl_sync(0xffffffffu, r4.z, i); rr[i][3] = __shfl_sync(0xffffffffu, r4.w, i);
        }
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
          if (ring.round > 0) mbar_wait_a(empty0 + 8u * ring.slot, (ring.round & 1) ^ 1);
          if (DBG && fine != nullptr && tid == 0 && cnt < 512) fine[cnt * 4] = os_now();
          if (DBG) ++cnt;
          if (!(DBG && (dbg & 2)) && elect_one_sync()) {
            const uint32_t dst = dst0 + ring.slot * (uint32_t)stage_bytes;
            const uint32_t bar = full0 + 8u * ring.slot;
#pragma unroll
            for (int i = 0; i < G; ++i)
              tma_gather4_a(dst + (uint32_t)i * 512u, &tmap, kb * 64, rr[i][0], rr[i][1], rr[i][2], rr[i][3], bar);
          }
          __syncwarp();
          ring.advance();
        }
      }
    }
    if (DBG && tr != nullptr && tid == 0) tr[1] = os_now(), tr[5] = npass;
  } else if (TMA && warp == kLoaderWarp) {
    // ------------------------------------------------------------------ weight loader: expect_tx + bulk copy of B_k
    OsPassIter it;
    it.init(a.units, U, first, step);
    int kq[kOsAhead];
    int queued = 0;
#pragma unroll
    for (int d = 0; d < kOsAhead; ++d) {
      kq[d] = 0;
      if (it.valid()) {
        kq[d] = __ldg(a.pass_k + it.pass());
        it.next();
        ++queued;
      }
    }
    OsRing ring;
    ring.init(nslots);
    const uint32_t tx = (uint32_t)(((DBG && (dbg & 2)) ? 0 : kBlockBytes) + ((DBG && (dbg & 4)) ? 0 : b_stage));
    const int nsp = ncols > 256 ? 2 : 1;
    const uint32_t half = (uint32_t)((ncols / nsp / 2) * kBlockRowBytes);   // pairs: this CTA's columns of one MMA's N range
    while (queued > 0) {
#pragma unroll
      for (int d = 0; d < kOsAhead; ++d) {
        if (queued == 0) break;
        const int k = a.kflip ? a.K - 1 - kq[d] : kq[d];
        --queued;
        if (it.valid()) {
          kq[d] = __ldg(a.pass_k + it.pass());
          it.next();
          ++queued;
        }
        const uint8_t* bsrc = a.wpacked + (size_t)k * nkb * b_bytes;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb, bsrc += b_bytes) {
          if (ring.round > 0) mbar_wait_a(empty0 + 8u * ring.slot, (ring.round & 1) ^ 1);
          if (elect_one_sync()) {
            const uint32_t bar = full0 + 8u * ring.slot;
            const uint32_t dst = smem0 + ring.slot * (uint32_t)stage_bytes + (uint32_t)kBlockBytes;
            mbar_expect_tx_a(bar, tx);
            if (DBG && (dbg & 4)) {
            } else if (PAIR) {        // columns [c*ncw + rank*ncw/2, +ncw/2) of each MMA's N range: this CTA's half
              for (int c = 0; c < nsp; ++c)
                bulk_g2s_a(dst + (uint32_t)c * half, bsrc + (size_t)(2 * c + rank) * half, half, bar);
            } else if (cs == 1) {
              bulk_g2s_a(dst, bsrc, (uint32_t)b_bytes, bar);
            } else if (rank == 0) {   // one L2 read for the whole cluster; every CTA's full[slot] gets its complete_tx
              bulk_g2s_multicast_a(dst, bsrc, (uint32_t)b_bytes, bar, cmask);
            }
          }
          __syncwarp();
          ring.advance();
        }
      }
    }
  } else if (!TMA && warp < 4) {
    // ------------------------------------------------------------------ producers: 16-byte cp.async row pieces
    // Thread t owns 16-byte chunk c = t & 7 of the tile rows r_j = (t >> 3) + 16 j, j = 0..7 (8 lanes = one 128-byte row
    // segment).  Its 8 gather indices of a pass are read straight from the schedule into registers, kOsAhead passes
    // ahead; a k-block is then 8 cp.async per thread -- one 64-bit multiply-add and the copy each.  (Staging the
    // indices in shared memory and walking a runtime-length loop cost a dependent shared-memory load per copy.)
    OsPassIter it;
    it.init(a.units, U, first, step);
    const int c = tid & 7, r0 = tid >> 3;
    const uint32_t dst_t = (uint32_t)(r0 * kBlockRowBytes) + (uint32_t)((c ^ (r0 & 7)) << 4);   // (r0 + 16 j) & 7 == r0 & 7
    int rq[kOsAhead][8];
    int queued = 0;
#pragma unroll
    for (int d = 0; d < kOsAhead; ++d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) rq[d][j] = -1;
      if (it.valid()) {
        const int32_t* pi = a.pass_idx + (int64_t)it.pass() * a.tile_rows + row_base + r0;
#pragma unroll
        for (int j = 0; j < 8; ++j) rq[d][j] = __ldg(pi + 16 * j);
        it.next();
        ++queued;
      }
    }
    OsRing ring;
    ring.init(nslots);
    const __nv_bfloat16* src0 = a.in + c * 8;
    uint32_t np = 0, cnt = 0;
    while (queued > 0) {
#pragma unroll
      for (int d = 0; d < kOsAhead; ++d) {
        if (queued == 0) break;
        int g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = rq[d][j];
        --queued;
        if (DBG) ++np;
        if (it.valid()) {
          const int32_t* pi = a.pass_idx + (int64_t)it.pass() * a.tile_rows + row_base + r0;
#pragma unroll
          for (int j = 0; j < 8; ++j) rq[d][j] = __ldg(pi + 16 * j);
          it.next();
          ++queued;
        }
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
          if (ring.round > 0) mbar_wait_a(empty0 + 8u * ring.slot, (ring.round & 1) ^ 1);
          if (DBG && fine != nullptr && tid == 0 && cnt < 512) fine[cnt * 4] = os_now();
          if (DBG) ++cnt;
          if (c * 8 < red - kb * 64) {
            const uint32_t dst = smem0 + ring.slot * (uint32_t)stage_bytes + dst_t;
            const __nv_bfloat16* src = src0 + kb * 64;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              os_cp_async_16(dst + (uint32_t)(16 * j * kBlockRowBytes), src + (g[j] >= 0 ? (int64_t)g[j] * red : 0),
                             g[j] >= 0 ? 16u : 0u);
          }
          os_cp_async_arrive_noinc(&hdr->full[ring.slot]);
          ring.advance();
        }
      }
    }
    if (DBG && tr != nullptr && tid == 0) tr[1] = os_now(), tr[5] = np;
  } else if (!TMA && warp == kLoaderWarp) {
    // ------------------------------------------------------------------ weight loader (LDGSTS mode)
    if (lane == 0) {
      OsPassIter it;
      it.init(a.units, U, first, step);
      uint32_t cnt = 0;
      for (; it.valid(); it.next()) {
        int k = __ldg(a.pass_k + it.pass());
        if (a.kflip) k = a.K - 1 - k;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int slot = (int)(cnt % (uint32_t)nslots);
          const uint32_t use = cnt / (uint32_t)nslots;
          if (use > 0) mbar_wait(&hdr->empty[slot], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&hdr->full[slot], (uint32_t)b_bytes);
          bulk_g2s(smem + (size_t)slot * stage_bytes + kBlockBytes, a.wpacked + ((size_t)k * nkb + kb) * b_bytes,
                   (uint32_t)b_bytes, &hdr->full[slot]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (PAIR && rank != 0) {
      // rank 1 of a pair: tell rank 0's MMA issuer when this CTA's half of a stage has landed
      if (lane == 0) {
        OsRing ring;
        ring.init(nslots);
        int np_n = first < U ? __ldg(a.units + 2 * (int64_t)first).y : 0;
#pragma unroll 1
        for (int u = first; u < U; u += step) {
          const int np = np_n;
          if (u + step < U) np_n = __ldg(a.units + 2 * (int64_t)(u + step)).y;
#pragma unroll 1
          for (int q = 0; q < np * nkb; ++q) {
            mbar_wait_a(full0 + 8u * ring.slot, ring.round & 1);
            mbar_arrive_remote(&hdr->pfull[ring.slot], 0);
            ring.advance();
          }
        }
      }
    } else {
      // All 32 lanes walk the schedule together and ONE elected lane issues: with uniform control flow the shared-
      // memory descriptors, TMEM addresses and barrier addresses live in uniform registers.  (Issued from inside an
      // `if (lane == 0)` region the compiler wraps every tcgen05 instruction in an elect/broadcast loop: ~30 dependent
      // instructions per MMA, measured 165 ns per MMA whatever its N.)
      const int nsplit = ncols > 256 ? 2 : 1;
      const int ncw = ncols / nsplit;
      const int b_rows = PAIR ? ncw / 2 : ncw;             // rows of one MMA's B operand held by this CTA
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, ncw, 0, 0);
      const uint64_t dfull = smem_desc_sw128(0, 16, 1024); // descriptor without its start address
      const uint32_t desc_hi = (uint32_t)(dfull >> 32), desc_lo0 = (uint32_t)dfull;
      const uint32_t c_step = (uint32_t)(b_rows * kBlockRowBytes) >> 4;
      const uint32_t pfull0 = smem_u32(&hdr->pfull[0]);
      const uint32_t last_ksteps = (uint32_t)((red - (nkb - 1) * 64) >> 4);   // k-steps of the last (maybe partial) k-block
      OsRing ring;
      ring.init(nslots);
      uint32_t cnt = 0, use_acc = 0;
      int np_n = first < U ? __ldg(a.units + 2 * (int64_t)first).y : 0;
#pragma unroll 1
      for (int u = first; u < U; u += step) {
        const int np = np_n;
        if (u + step < U) np_n = __ldg(a.units + 2 * (int64_t)(u + step)).y;
        if (np == 0) continue;
        const int buf = (int)(use_acc % (uint32_t)a.nbuf);
        const uint32_t ub = use_acc / (uint32_t)a.nbuf;
        if (ub > 0) {                                                  // the epilogue has drained this accumulator
          if (PAIR) mbar_wait_cluster(&hdr->acc_empty[buf], (ub & 1) ^ 1);
          else mbar_wait(&hdr->acc_empty[buf], (ub & 1) ^ 1);
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * a.tcols);
        uint32_t acc = 0;                                              // the unit's first k-step overwrites
#pragma unroll 1
        for (int q = 0; q < np; ++q) {
#pragma unroll 1
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait_a(full0 + 8u * ring.slot, ring.round & 1);
            if (PAIR) mbar_wait_cluster_a(pfull0 + 8u * ring.slot, ring.round & 1);
            if (!TMA) fence_proxy_async_smem();            // cp.async (generic proxy) data -> tensor-core reads
            tc_fence_after();
            if (DBG && fine != nullptr && lane == 0 && cnt < 512) fine[cnt * 4 + 1] = os_now();
            const uint32_t a_lo = desc_lo0 | ((smem0 + ring.slot * (uint32_t)stage_bytes) >> 4);
            const uint32_t b_lo = a_lo + (uint32_t)(kBlockBytes >> 4);
            uint32_t ksteps = kb == nkb - 1 ? last_ksteps : 4u;
            if (DBG && (dbg & 1)) ksteps = 0;
            if (elect_one_sync()) {
              umma_bf16_kblock<PAIR>(tmem_d, a_lo, b_lo, desc_hi, idesc, acc, ksteps);
              if (nsplit == 2) umma_bf16_kblock<PAIR>(tmem_d + (uint32_t)ncw, a_lo, b_lo + c_step, desc_hi, idesc, acc, ksteps);
              umma_commit_a<PAIR>(empty0 + 8u * ring.slot, cs, cmask);
            }
            __syncwarp();
            acc = 1;
            if (DBG && fine != nullptr && lane == 0 && cnt < 512) fine[cnt * 4 + 2] = os_now();
            if (DBG) ++cnt;
            ring.advance();
          }
        }
        if (elect_one_sync()) umma_commit_a<PAIR>(smem_u32(&hdr->acc_full[buf]), 1, cmask);
        __syncwarp();
        ++use_acc;
      }
      if (DBG && tr != nullptr && lane == 0) tr[2] = os_now();
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue (TMEM quadrant = warp % 4)
    const int q = warp & 3, te = tid - kEpiWarp0 * 32;
    float* st = staging + (size_t)(warp - kEpiWarp0) * kOsStageFloats;
    float* ws1 = sstat + (size_t)(warp - kEpiWarp0) * 2 * ncols;
    float* ws2 = ws1 + ncols;
    const bool stats = a.partials != nullptr;
    if (stats)
      for (int c = lane; c < 2 * ncols; c += 32) ws1[c] = 0.f;
    // rows beyond the real row count of a capacity-padded output stay exactly zero (graph.py)
    if (a.valid_rows != nullptr) {
      const int64_t v = (int64_t)__ldg(a.valid_rows);
      const int c4 = ncols >> 2;
      for (int64_t row = v + first; row < a.n_out; row += step)
        for (int c = te; c < c4; c += 128)
          *reinterpret_cast<float4*>(a.out + row * ncols + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    uint32_t use_acc = 0, nunits = 0, nsplit_units = 0;
    int4 u0_n = make_int4(0, 0, 0, 1), u1_n = make_int4(0, 0, 0, 0);
    if (first < U) u0_n = __ldg(a.units + 2 * (int64_t)first), u1_n = __ldg(a.units + 2 * (int64_t)first + 1);
#pragma unroll 1
    for (int u = first; u < U; u += step, ++nunits) {
      const int4 u0 = u0_n, u1 = u1_n;                     // {first pass, passes, tile, chunks}, {chunk, scratch base}
      if (u + step < U) u0_n = __ldg(a.units + 2 * (int64_t)(u + step)), u1_n = __ldg(a.units + 2 * (int64_t)(u + step) + 1);
      const int tile = u0.z, chunks = u0.w;
      if (tile < 0) continue;                              // hole of the unit placement (os_plan.cu)
      const int my_row = __ldg(a.out_row + (int64_t)tile * a.tile_rows + row_base + q * 32 + lane);
      if (u0.y == 0) {                                     // rows without any neighbour: the result is zero
        if (my_row >= 0)
          for (int c = 0; c < ncols; c += 4)
            *reinterpret_cast<float4*>(a.out + (int64_t)my_row * ncols + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      const int buf = (int)(use_acc % (uint32_t)a.nbuf);
      mbar_wait(&hdr->acc_full[buf], (use_acc / (uint32_t)a.nbuf) & 1);
      tc_fence_after();
      if (fine != nullptr && te == 0 && use_acc < 64) fine[2048 + use_acc * 2] = os_now();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * a.tcols) + ((uint32_t)(q * 32) << 16);
      // split tile: this unit's partial rows go to its scratch slot, [slot][tile row][ncols]
      float* part = chunks > 1 ? a.scratch + ((size_t)(u1.y + u1.x) * a.tile_rows + row_base + q * 32) * ncols : nullptr;
      nsplit_units += chunks > 1 ? 1u : 0u;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(st + lane * 36 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        __syncwarp();
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
        const float4 pv = stats ? stat_pivot(a.pivot, c0 + (lane & 7) * 4) : s1;
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {                   // a warp store = 4 rows x 128 contiguous bytes
          const int r = i8 * 4 + (lane >> 3);
          const int row = __shfl_sync(0xffffffffu, my_row, r);
          const float4 val = *reinterpret_cast<const float4*>(st + r * 36 + (lane & 7) * 4);
          if (part != nullptr) {
            *reinterpret_cast<float4*>(part + (size_t)r * ncols + c0 + (lane & 7) * 4) = val;
          } else if (row >= 0) {
            *reinterpret_cast<float4*>(a.out + (int64_t)row * ncols + c0 + (lane & 7) * 4) = val;
            stat_add(s1, s2, val, pv);
          }
        }
        if (stats && part == nullptr) {                    // fold the 4 row groups (lane >> 3), lanes 0-7 keep the sums
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
            s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
            s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
            s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
          }
          if (lane < 8) {
            float4* p1 = reinterpret_cast<float4*>(ws1 + c0 + lane * 4);
            float4* p2 = reinterpret_cast<float4*>(ws2 + c0 + lane * 4);
            float4 x1 = *p1, x2 = *p2;
            x1.x += s1.x; x1.y += s1.y; x1.z += s1.z; x1.w += s1.w;
            x2.x += s2.x; x2.y += s2.y; x2.z += s2.z; x2.w += s2.w;
            *p1 = x1; *p2 = x2;
          }
        }
        __syncwarp();
      }
      if (fine != nullptr && te == 0 && use_acc < 64) fine[2048 + use_acc * 2 + 1] = os_now();
      tc_fence_before();
      if (PAIR) mbar_arrive_remote(&hdr->acc_empty[buf], 0);
      else mbar_arrive(&hdr->acc_empty[buf]);
      ++use_acc;
    }
    if (tr != nullptr && te == 0) tr[3] = os_now(), tr[6] = nunits | ((unsigned long long)nsplit_units << 32);
    if (stats) {
      // CTA partial row = sum of the four warps in fixed order; col_finalize_kernel (bn_common.cuh, its own tiny
      // launch: channels/4 CTAs read the <= 148 partial rows in ONE L2 round trip) folds them in double.  A "last CTA
      // folds everything" tail was measured at 7.5 us per layer: 37 dependent L2 round trips on one SM.
      os_named_bar_sync(2, 128);
      float* prow = a.partials + (size_t)blockIdx.x * 2 * ncols;
      for (int c = te; c < 2 * ncols; c += 128)
        prow[c] = (sstat[c] + sstat[2 * ncols + c]) + (sstat[4 * ncols + c] + sstat[6 * ncols + c]);
    }
    if (tr != nullptr && te == 0) tr[4] = os_now();
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 0) {
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)(a.nbuf * a.tcols));
    else tmem_dealloc(tmem_base, (uint32_t)(a.nbuf * a.tcols));
  }
}

// Fold of the split tiles: out[row] = sum of the tile's unit partials IN UNIT ORDER (ascending offsets), one CTA per
// split tile.  All loads of a thread are independent (rows x units), so the kernel is one or two L2 round trips long;
// it touches only the 8-15 % of the rows that live in heavy tiles.  With STATS each CTA adds its partial statistics row
// behind conv_os's (CTAs without a tile publish zeros, so the finalize reads a fixed number of rows).
template <bool STATS>
__global__ void __launch_bounds__(kColThreads)
conv_os_fold_kernel(const int4* __restrict__ split_tiles, const int32_t* __restrict__ num,
                    const int32_t* __restrict__ out_row, const float* __restrict__ scratch, int ncols, int tile_rows,
                    float* __restrict__ out, float* __restrict__ partials, const float* __restrict__ pivot) {
  pdl_enter();
  __shared__ float4 s_stage[STATS ? kColStageFloat4 : 1];
  __shared__ int s_rows[4 * kTileRows];
  const int ns = __ldg(num + 4);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  // blockIdx.y = slice of the tile's rows: the fold is a handful of independent loads per thread, so more CTAs per
  // tile shorten it to ~one L2 round trip
  const int rows_per_slice = tile_rows / (int)gridDim.y;
  const int slice0 = (int)blockIdx.y * rows_per_slice;
  if ((int)blockIdx.x < ns) {
    const int4 sp = __ldg(split_tiles + blockIdx.x);       // {tile, units, first scratch slot, -}
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int r = tid; r < rows_per_slice; r += blockDim.x * blockDim.y)
      s_rows[r] = __ldg(out_row + (int64_t)sp.x * tile_rows + slice0 + r);
    __syncthreads();
    const int ch = threadIdx.x * 4;
    const float4 pv = STATS ? stat_pivot(pivot, ch) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p0 = scratch + (size_t)sp.z * tile_rows * ncols + ch;
    const size_t cstride = (size_t)tile_rows * ncols;
    constexpr int RB = 4;
    for (int r0 = threadIdx.y; r0 < rows_per_slice; r0 += RB * blockDim.y) {
      float4 x[RB][4];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = r0 + i * blockDim.y;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          x[i][c] = (r < rows_per_slice && c < sp.y)
                        ? __ldg(reinterpret_cast<const float4*>(p0 + c * cstride + (size_t)(slice0 + r) * ncols))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = r0 + i * blockDim.y;
        if (r >= rows_per_slice) break;
        const int row = s_rows[r];
        if (row < 0) continue;
        float4 acc = x[i][0];
#pragma unroll
        for (int c = 1; c < 4; ++c)
          if (c < sp.y) add4(acc, x[i][c]);
        *reinterpret_cast<float4*>(out + (int64_t)row * ncols + ch) = acc;
        if (STATS) stat_add(s1, s2, acc, pv);
      }
    }
  }
  if (STATS) col_publish(s1, s2, partials + (size_t)blockIdx.y * gridDim.x * 2 * ncols, ncols, s_stage);
}

static int os_tmem_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor [n_rows][red], box = 64 columns x 1 row, 128B swizzle: the descriptor of the gather4 loads
static int make_row_tmap(CUtensorMap* tm, const void* base, int64_t n_rows, int red) {
  EncodeTiledFn fn = encode_fn();
  FT3D_REQUIRE(fn != nullptr, "ft3d_conv_os: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)red, (cuuint64_t)(n_rows > 0 ? n_rows : 1)};
  cuuint64_t strides[1] = {(cuuint64_t)red * 2};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FT3D_REQUIRE(r == CUDA_SUCCESS, "ft3d_conv_os: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return FT3D_OK;
}

// 1 = TMA gather4, 0 = 16-byte cp.async (LDGSTS).  FT3D_OS_GATHER=tma|ldgsts forces one; by default the narrower layers
// (red <= 192) take the cp.async path and the wide ones the TMA path: the TMA unit accepts one gathered row per ~7.5
// cycles whatever its width (measured: a 128-row block every ~0.5 us, tools/conv_os_probe.py --fine), which bounds
// the layers whose stages are small; 128 threads issuing 8 cp.async each from register-held indices are 10-25 %
// faster there and equal on the 256-384-wide layers, whose stages are bounded by the weight blocks (read per call).
static int os_gather_mode(int red, int cs) {
  const char* e = getenv("FT3D_OS_GATHER");
  if (e != nullptr && (e[0] == 'l' || e[0] == 'L')) return 0;
  if (e != nullptr && (e[0] == 't' || e[0] == 'T')) return 1;
  return (cs > 1 || red > 192) ? 1 : 0;            // cluster schedules need the TMA path
}

static bool os_pair_mode() {       // 256-row tiles: cta_group::2 pairs (default) or, FT3D_OS_PAIR=0, two CTAs + B multicast
  const char* e = getenv("FT3D_OS_PAIR");
  return e == nullptr || e[0] != '0';
}

static int os_gather_warps() {    // FT3D_OS_PRODUCERS = 4 | 8 TMA gather warps per CTA (read per call)
  const char* e = getenv("FT3D_OS_PRODUCERS");
  if (e != nullptr && e[0] == '4') return 4;
  if (e != nullptr && e[0] == '8') return 8;
  return kOsDefaultGatherWarps;
}

typedef void (*OsKernelFn)(const CUtensorMap, const OsArgs);
struct OsVariant {
  bool tma, pair;
  int npw;
  bool dbg;
  OsKernelFn fn;
};
#define FT3D_OS_V(T, P, N, D) {T, P, N, D, conv_os_kernel<T, P, N, D>}
static const OsVariant kOsVariants[] = {
    FT3D_OS_V(true, false, 4, false),  FT3D_OS_V(true, false, 8, false),  FT3D_OS_V(true, true, 4, false),
    FT3D_OS_V(true, true, 8, false),   FT3D_OS_V(false, false, 4, false), FT3D_OS_V(true, false, 4, true),
    FT3D_OS_V(true, false, 8, true),   FT3D_OS_V(true, true, 4, true),    FT3D_OS_V(true, true, 8, true),
    FT3D_OS_V(false, false, 4, true),
};
#undef FT3D_OS_V
static const OsVariant* os_variant(bool tma, bool pair, int npw, bool dbg) {
  for (const OsVariant& v : kOsVariants)
    if (v.tma == tma && v.pair == pair && v.npw == npw && v.dbg == dbg) return &v;
  return nullptr;
}
static cudaError_t os_configure() {          // once per process: opt in to 227 KB of dynamic shared memory
  static cudaError_t state = cudaErrorUnknown;
  if (state == cudaErrorUnknown) {
    state = cudaSuccess;
    for (const OsVariant& v : kOsVariants) {
      cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) state = e;
    }
  }
  return state;
}

constexpr int kOsMaxCtas = kNumSMs;

constexpr int kOsFoldSlices = 4;                 // CTAs per split tile in the fold kernel
static int64_t os_fold_ctas(int64_t scratch_slots) { return (scratch_slots + 1) / 2; }     // a split tile has >= 2 units
static size_t os_partials_bytes(int ncols, int64_t scratch_slots) {
  return align_up((size_t)(kOsMaxCtas + kOsFoldSlices * os_fold_ctas(scratch_slots)) * 2 * (size_t)ncols * sizeof(float), 256);
}

}  // namespace ft3d

using namespace ft3d;

extern "C" {

size_t ft3d_conv_os_workspace(int32_t ncols, int64_t scratch_slots, int32_t tile_rows) {
  return os_partials_bytes(ncols, scratch_slots) +
         align_up((size_t)scratch_slots * (size_t)tile_rows * (size_t)ncols * sizeof(float), 256);
}

int ft3d_conv_os(const void* in_bf16, int64_t n_in, const int32_t* units, const int32_t* split_tiles,
                 const int32_t* num, const int32_t* out_row, const int32_t* pass_k, const int32_t* pass_idx,
                 int64_t unit_cap, int64_t tiles, int32_t tile_rows, int64_t scratch_slots, int32_t K, int32_t kflip,
                 int32_t red,
                 int32_t ncols, const void* wpacked, float* out, int64_t n_out, const int32_t* valid_rows, float eps,
                 float momentum, float* stat, float* running_mean, float* running_var, void* workspace,
                 size_t workspace_bytes, void* trace, ft3d_stream_t stream) {
  if (tiles == 0 || n_out == 0) return FT3D_OK;
  FT3D_REQUIRE(in_bf16 && units && num && out_row && pass_k && pass_idx && wpacked && out && K > 0 && K <= 32 &&
                   n_in > 0 && unit_cap >= tiles && scratch_slots >= 0 && (scratch_slots == 0 || split_tiles) &&
                   (tile_rows == 128 || tile_rows == 256 || tile_rows == 512),
               "ft3d_conv_os: bad arguments");
  FT3D_REQUIRE(red >= 16 && red % 16 == 0 && red <= 512 && ncols >= 32 && ncols % 32 == 0 &&
                   (ncols <= 256 || ncols == 384),
               "ft3d_conv_os: unsupported shape red=%d ncols=%d", red, ncols);
  FT3D_REQUIRE(((uintptr_t)in_bf16 & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)wpacked & 15) == 0 &&
                   ((uintptr_t)pass_idx & 15) == 0 && ((uintptr_t)units & 15) == 0 && ((uintptr_t)split_tiles & 15) == 0,
               "ft3d_conv_os: pointers must be 16-byte aligned");
  const bool stats = stat != nullptr;
  FT3D_REQUIRE((!stats && scratch_slots == 0) ||
                   (workspace && ((uintptr_t)workspace & 255) == 0 &&
                    workspace_bytes >= ft3d_conv_os_workspace(ncols, scratch_slots, tile_rows)),
               "ft3d_conv_os: needs a 256-byte aligned workspace of ft3d_conv_os_workspace(ncols, slots, tile_rows) bytes");
  FT3D_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "ft3d_conv_os: running stats go together");
  OsArgs a;
  a.in = (const __nv_bfloat16*)in_bf16;
  a.units = (const int4*)units;
  a.num = num;
  a.out_row = out_row;
  a.pass_k = pass_k;
  a.pass_idx = pass_idx;
  a.wpacked = (const uint8_t*)wpacked;
  a.out = out;
  a.partials = stats ? (float*)workspace : nullptr;
  a.scratch = workspace ? (float*)((char*)workspace + os_partials_bytes(ncols, scratch_slots)) : nullptr;
  a.valid_rows = valid_rows;
  a.pivot = stats ? running_mean : nullptr;
  a.trace = (unsigned long long*)trace;
  a.n_out = n_out;
  a.unit_cap = (int)unit_cap; a.K = K; a.kflip = kflip; a.red = red; a.ncols = ncols;
  a.tcols = os_tmem_cols(ncols);
  a.nbuf = 2 * a.tcols <= 512 ? 2 : 1;
  a.cs = tile_rows / tc::kTileRows;
  a.tile_rows = tile_rows;
  {
    const char* e = getenv("FT3D_OS_DEBUG");
    a.dbg = e ? atoi(e) : 0;
  }
  const bool tma = os_gather_mode(red, a.cs) == 1;
  const bool pair = tma && a.cs == 2 && os_pair_mode();      // 256-row tiles: one cta_group::2 MMA per CTA pair
  const int stage_bytes = tc::kBlockBytes + ncols * tc::kBlockRowBytes / (pair ? 2 : 1);
  const int fixed = 4 * kOsStageFloats * (int)sizeof(float) + 8 * ncols * (int)sizeof(float) + (int)sizeof(OsHeader) + 1024 + 64;
  int nslots = (227 * 1024 - fixed) / stage_bytes;
  if (nslots > kOsMaxSlots) nslots = kOsMaxSlots;
  FT3D_REQUIRE(nslots >= 2, "ft3d_conv_os: red=%d ncols=%d does not fit shared memory", red, ncols);
  a.nslots = nslots;
  const int smem_bytes = fixed + nslots * stage_bytes;
  // every cluster runs: the unit array is laid out for kNumSMs / cs clusters (os_assign_kernel), and a map with fewer
  // tiles than SMs still has enough units (split tiles) to occupy them
  const unsigned grid = (unsigned)((kOsMaxCtas / a.cs) * a.cs);
  CUtensorMap tm;
  FT3D_REQUIRE(tma || a.cs == 1, "ft3d_conv_os: cluster schedules (tile_rows > 128) need the TMA gather mode");
  if (tma) {
    int rc = make_row_tmap(&tm, in_bf16, n_in, red);
    if (rc) return rc;
  } else {
    memset(&tm, 0, sizeof(tm));
  }
  const int npw = tma ? os_gather_warps() : 4;
  const bool dbgk = (a.dbg & ~8) != 0 || a.trace != nullptr;     // timing experiments / trace: the DBG instantiation
  const OsVariant* v = os_variant(tma, pair, npw, dbgk);
  FT3D_REQUIRE(v != nullptr, "ft3d_conv_os: no kernel variant");
  FT3D_CUDA(os_configure());
  cudaStream_t s = (cudaStream_t)stream;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((unsigned)((npw + 6) * 32));
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (a.cs > 1) {                  // thread-block clusters: CTA pairs (cta_group::2) or B multicast
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)a.cs;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaLaunchKernelEx(&cfg, v->fn, tm, a);
  int nparts = (int)grid;
  if (scratch_slots > 0 && !(a.dbg & 8)) {       // the schedule has split tiles: fold their unit partials
    const int fgrid = (int)os_fold_ctas(scratch_slots);
    const int cv = ncols / 4;
    int ry = kColThreads / cv;
    if (ry < 1) ry = 1;
    if (ry > 32) ry = 32;
    float* fparts = stats ? a.partials + (size_t)grid * 2 * ncols : nullptr;
    if (stats)
      launch_pdl(conv_os_fold_kernel<true>, dim3(fgrid, kOsFoldSlices), dim3(cv, ry), 0, s, (const int4*)split_tiles, num,
                 out_row, (const float*)a.scratch, (int)ncols, (int)tile_rows, out, fparts, a.pivot);
    else
      launch_pdl(conv_os_fold_kernel<false>, dim3(fgrid, kOsFoldSlices), dim3(cv, ry), 0, s, (const int4*)split_tiles,
                 num, out_row, (const float*)a.scratch, (int)ncols, (int)tile_rows, out, fparts, a.pivot);
    nparts += fgrid * kOsFoldSlices;
  }
  if (stats && !(a.dbg & 8))
    launch_pdl(col_finalize_kernel<0>, dim3(ncols / 4), dim3(kColThreads), 0, s, (const float*)a.partials, nparts,
               (int)ncols, n_out, eps, momentum, stat, running_mean, running_var, 0, valid_rows);
  return check_launch("ft3d_conv_os");
}

}  // extern "C"
