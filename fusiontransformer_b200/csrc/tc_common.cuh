// tcgen05 / TMEM / mbarrier / bulk-copy primitives for sm_100a (inline PTX; no CUTLASS dependency).
//
// Shared-memory operand layout used by every tensor-core kernel in libft3d: a "block" is
// [rows x 128 bytes] = rows x 64 bf16, 128-byte swizzled (16-byte chunk c of row r is stored at chunk
// c ^ (r & 7)); blocks are 1024-byte aligned.  Read as a K-major operand (rows = M/N index, the 64
// elements = reduction index) it is the canonical SWIZZLE_128B K-major layout with SBO = 1024; read as an
// MN-major operand (rows = reduction index, the 64 elements = M/N index) it is the canonical
// SWIZZLE_128B MN-major layout with SBO = 1024 (8-row groups) and LBO = the distance between blocks.
// So one gather routine feeds the forward/dgrad GEMMs (K-major A) and the wgrad GEMM (MN-major A and B).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace ft3d {
namespace tc {

constexpr int kBlockRowBytes = 128;                     // one swizzle row: 64 bf16
constexpr int kTileRows = 128;                          // UMMA M
constexpr int kBlockBytes = kTileRows * kBlockRowBytes; // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// true in exactly one lane of the (fully active) warp.  tcgen05 / TMA instructions take their operands from uniform
// registers: issued under `if (elect_one_sync())` inside warp-uniform control flow they cost a few instructions, issued
// under `if (lane == 0)` the compiler wraps each one in an elect / broadcast loop.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- 1-D bulk copy global -> shared (TMA unit, UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 1-D bulk copy shared -> global (bulk_group completion): the source must not be overwritten before wait_group.read
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: arrives on the mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 operands, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of warp w reads TMEM lane (w%4)*32 + t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// 64-bit shared-memory matrix descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | base_offset=0 [49,52) | layout [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 32-bit instruction descriptor, kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn_major & 1) << 15) | ((uint32_t)(b_mn_major & 1) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------- gathered operand blocks
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Fill one [128 rows x nchunk*16 B] swizzled block from fp32 rows: block row r <- src[row_idx(r)][col0 .. col0+8*nchunk)
// (row index < 0 => zeros).  Executed by `nthreads` producer threads (tid in [0,nthreads)); all global loads of a
// thread are issued before the first conversion so that up to 2*ITER 16-byte requests are in flight per thread.
template <int NTHREADS, typename RowIdx>
__device__ __forceinline__ void fill_block_f32(uint8_t* block, const float* __restrict__ src, int row_width, int col0,
                                               int nchunk, int tid, RowIdx row_idx) {
  const int total = kTileRows * nchunk;
  constexpr int ITER = 4;
  for (int base = 0; base < total; base += NTHREADS * ITER) {
    float4 lo[ITER], hi[ITER];
    int dst[ITER];
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
      int q = base + i * NTHREADS + tid;
      dst[i] = -1;
      lo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      hi[i] = lo[i];
      if (q < total) {
        int r = q / nchunk;
        int c = q - r * nchunk;
        dst[i] = r * kBlockRowBytes + ((c ^ (r & 7)) << 4);
        int64_t g = row_idx(r);
        if (g >= 0) {
          const float4* p = reinterpret_cast<const float4*>(src + g * row_width + col0 + c * 8);
          lo[i] = __ldg(p);
          hi[i] = __ldg(p + 1);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
      if (dst[i] >= 0) {
        uint4 v;
        v.x = pack_bf16x2(lo[i].x, lo[i].y);
        v.y = pack_bf16x2(lo[i].z, lo[i].w);
        v.z = pack_bf16x2(hi[i].x, hi[i].y);
        v.w = pack_bf16x2(hi[i].z, hi[i].w);
        *reinterpret_cast<uint4*>(block + dst[i]) = v;
      }
    }
  }
}

}  // namespace tc
}  // namespace ft3d
