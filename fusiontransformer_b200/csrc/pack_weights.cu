// bf16 weight images of the tcgen05 convolution kernels (conv_os.cu, conv_pairs_tc.cu), sm_100a.
//
// The fp32 parameters keep the reference's checkpoint layout [K, Cin, Cout]; the tensor-core kernels stream B_k with
// cp.async.bulk from a per-layer packed image that is rewritten after every optimizer step (one launch per model:
// ft3d_conv_pack_weights_multi).
#include "common.cuh"
#include "tc_common.cuh"

namespace ft3d {
using namespace tc;

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

}  // extern "C"
