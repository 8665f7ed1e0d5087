/* libft3d -- C ABI of the B200-native 3D-branch hot path of FusionTransformer.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + the caller's CUDA
 * stream, never allocates or frees device memory, never synchronises the host (except where
 * stated) and returns 0 on success.  On failure it returns a non-zero code and
 * ft3d_last_error() holds a thread-local message.  All work is enqueued on `stream`.
 *
 * The reference (aliabdelkader/FusionTransformer) reaches this arithmetic through the
 * torchsparse v1.1.0 Python API (docker/Dockerfile:33); each function names the reference
 * call site (file:line under /root/reference/FusionTransformer) it serves and the upstream
 * operator it replaces.
 */
#ifndef FT3D_H_
#define FT3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ft3d_stream_t; /* cudaStream_t */

#define FT3D_VERSION 100
#define FT3D_KPAD_K3 32 /* row pitch (int32) of a 27-offset neighbour table */
#define FT3D_KPAD_K2 8  /* row pitch (int32) of an 8-offset neighbour table  */

int ft3d_version(void);
const char* ft3d_last_error(void);

/* ---- K1/K2  spf.sphash  (models/utils.py:19,46-52,74-79) --------------------------------- */
/* out[i] = fold60(FNV1a64(x,y,z,b)) for coords int32 [n,4]. */
int ft3d_hash(const int32_t* coords, int64_t n, int64_t* out, ft3d_stream_t stream);
/* out[k*n+i] = hash(x+dx_k, y+dy_k, z+dz_k, b); offsets int32 [K,3]. */
int ft3d_kernel_hash(const int32_t* coords, int64_t n, const int32_t* offsets, int32_t K,
                     int64_t* out, ft3d_stream_t stream);

/* ---- a1  augment_and_scale_3d + cast + bounds (data/utils/augmentation_3d.py:43-46,
 *          data/semantic_kitti/semantic_kitti_dataloader.py:220,225), no augmentation -------- */
/* points f32 [n,3]; scan_id int32 [n] in [0,num_scans); min_ws f32 [num_scans*3] scratch.
 * coords_out int32 [n,4] = (trunc(p*scale - min_scan(p*scale)), scan); keep_out[i] = all in [0,full_scale). */
int ft3d_scale_coords(const float* points, const int32_t* scan_id, int64_t n, int32_t num_scans,
                      float scale, int32_t full_scale, int32_t* coords_out, uint8_t* keep_out,
                      float* min_ws, ft3d_stream_t stream);
/* The same with the augmentation branch (data/utils/augmentation_3d.py:22-41,48-51).  The caller draws the random
 * numbers on the host in the reference's order (a dozen per scan) and passes them: rot f32 [num_scans,3,3] (nullable:
 * no rotation; p' = p . rot evaluated as numpy's float32 sgemm does), transl_u f64 [num_scans,3] (nullable: the
 * uniform factors of the random translation; offset = clip(full_scale - max - 0.001, 0) * u in float64, added into the
 * float32 coordinate).  ws f32 [num_scans*6] scratch (per-scan minima and maxima). */
int ft3d_augment_scale_coords(const float* points, const int32_t* scan_id, int64_t n, int32_t num_scans,
                              float scale, int32_t full_scale, const float* rot, const double* transl_u,
                              int32_t* coords_out, uint8_t* keep_out, float* ws, ft3d_stream_t stream);

/* ---- a2 / K15  torchsparse.utils.sparse_quantize (semantic_kitti_dataloader.py:231) ------- */
/* coords int32 [n,4] (x,y,z,scan), scans contiguous and ascending.  Per scan: unique 64-bit
 * FNV-1 keys in ascending order.  inds_out[g] (g < *num_unique_out) = global row of the first
 * occurrence of unique voxel g (groups ordered by scan, then key); inverse_out[i] = rank of
 * row i's voxel inside its scan; scan_counts_out[s] = unique voxels of scan s. */
size_t ft3d_quantize_workspace(int64_t n, int32_t num_scans);
int ft3d_quantize(const int32_t* coords, int64_t n, int32_t num_scans, int32_t* inds_out,
                  int32_t* inverse_out, int32_t* scan_counts_out, int32_t* num_unique_out,
                  void* workspace, size_t workspace_bytes, ft3d_stream_t stream);

/* ---- K4  torch.unique(sorted, return_inverse, return_counts) on hashes
 *          (models/utils.py:20; torchsparse spdownsample) ----------------------------------- */
/* uniq_out/counts_out/first_out hold *num_out valid entries; first_out = smallest row index of
 * each group (stable). */
size_t ft3d_unique_workspace(int64_t n);
int ft3d_unique(const int64_t* keys, int64_t n, int64_t* uniq_out, int32_t* inverse_out,
                int32_t* counts_out, int32_t* first_out, int32_t* num_out, void* workspace,
                size_t workspace_bytes, ft3d_stream_t stream);
/* coarse_out[i] = (floor(x/ratio)*ratio, ..., b); hash_out[i] = ft3d_hash(coarse_out[i]). */
int ft3d_coarsen_hash(const int32_t* coords, int64_t n, int32_t ratio, int32_t* coarse_out,
                      int64_t* hash_out, ft3d_stream_t stream);
/* out[g] = src[first[g]] for rows of `width` int32. */
int ft3d_gather_rows_i32(const int32_t* src, const int32_t* first, int64_t m, int32_t width,
                         int32_t* out, ft3d_stream_t stream);

/* ---- K3  spf.sphashquery (models/utils.py:21,50,80; torchsparse conv3d) -------------------- */
/* Open-addressing table owned by the caller: table_keys uint64 [cap], table_vals int32 [cap],
 * cap = ft3d_table_capacity(n) (power of two).  Key -> row index; duplicate keys keep the
 * smallest index.  The key value 0xFFFF...F is reserved. */
int64_t ft3d_table_capacity(int64_t n);
int ft3d_table_build(const int64_t* keys, int64_t n, uint64_t* table_keys, int32_t* table_vals,
                     int64_t cap, ft3d_stream_t stream);
int ft3d_table_query(const int64_t* queries, int64_t m, const uint64_t* table_keys,
                     const int32_t* table_vals, int64_t cap, int64_t* out, ft3d_stream_t stream);

/* ---- a6/a7/K9  kernel maps (torchsparse conv3d map build, triggered at models/spvcnn.py:99,
 *               105-124) ---------------------------------------------------------------------- */
/* nbr_out[j*kpad+k] = row in the table's coordinate set of (coords_q[j] + offsets[k]) or -1;
 * columns K..kpad-1 = -1.  The table must have been built from ft3d_hash of the input coords. */
int ft3d_kmap_build(const int32_t* coords_q, int64_t n_out, const int32_t* offsets, int32_t K,
                    const uint64_t* table_keys, const int32_t* table_vals, int64_t cap,
                    int32_t* nbr_out, int32_t kpad, ft3d_stream_t stream);
/* Reference-format map (torchsparse convert_neighbor_map_gpu): pairs_out int32 [L,2] =
 * (in,out), offset-major then out-ascending; offsets_out int32 [K+1] exclusive prefix of the
 * per-offset counts (offsets_out[K] = L).  pairs_out must hold K*n_out rows. */
size_t ft3d_kmap_pairs_workspace(int64_t n_out, int32_t kpad);
/* ppos_out (nullable) int32 [n_out,kpad]: the positions in pairs_out of row j's pairs, compacted to the front of the
 * row in ascending offset order and terminated by -1 (the row is the CSR segment of output j, pitch kpad). */
int ft3d_kmap_pairs(const int32_t* nbr, int64_t n_out, int32_t K, int32_t kpad,
                    int32_t* pairs_out, int32_t* offsets_out, int32_t* ppos_out, void* workspace,
                    size_t workspace_bytes, ft3d_stream_t stream);
/* ppos_out int32 [n_rows,kpad]: the same compacted pair-position rows seen from the other side of the map: row
 * r lists, in ascending offset order, the positions p with pairs[p][col] == r (col 0: input rows, for dgrad and
 * transposed conv; col 1 reproduces the table of ft3d_kmap_pairs). */
int ft3d_kmap_pair_positions(const int32_t* pairs, const int32_t* pair_offsets, int32_t K, int32_t kpad,
                             int32_t col, int64_t n_rows, int64_t max_pairs, int32_t* ppos_out,
                             ft3d_stream_t stream);
/* nbrT_out[i*kpad+k] = j where nbr[j*kpad+k] == i, else -1 (input-stationary view for dgrad and
 * transposed conv, models/spvcnn.py:38-50). */
int ft3d_kmap_transpose(const int32_t* nbr, int64_t n_out, int32_t K, int32_t kpad,
                        int32_t* nbrT_out, int64_t n_in, ft3d_stream_t stream);

/* ---- K5-K8  point<->voxel (models/utils.py:22-27,51-58,81-99) ------------------------------ */
int ft3d_count(const int32_t* idx, int64_t n, int32_t* cnt_out, int64_t m, ft3d_stream_t stream);
int ft3d_voxelize_fwd(const float* feat, const int32_t* idx, const int32_t* cnt, int64_t n,
                      int64_t m, int32_t c, float* out, ft3d_stream_t stream);
int ft3d_voxelize_bwd(const float* gout, const int32_t* idx, const int32_t* cnt, int64_t n,
                      int64_t m, int32_t c, float* gin, ft3d_stream_t stream);
int ft3d_devoxelize_fwd(const float* feat, const int32_t* idx, const float* w, int64_t n,
                        int64_t m, int32_t c, float* out, ft3d_stream_t stream);
int ft3d_devoxelize_bwd(const float* gout, const int32_t* idx, const float* w, int64_t n,
                        int64_t m, int32_t c, float* gfeat, ft3d_stream_t stream);
/* Deterministic form of the two scatter-adds above (FT3D_DETERMINISTIC=1 in the Python boundary): the caller sorts the
 * contributions by destination row (stable) and passes them as a CSR list; a thread group owns a destination row and
 * adds its terms in list order -- no float atomics, bit-identical from run to run.
 *   out[v,:] = sum_{e in [offsets[v], offsets[v+1])}  t(e),   t(e) = entry_w[e] * src[entry_row[e],:]   (cnt == NULL)
 *                                                              t(e) = src[entry_row[e],:] / cnt[v]      (cnt != NULL)
 * entry_row int32 [E], entry_w f32 [E] (nullable), offsets int32 [m+1], cnt int32 [m] (nullable), out f32 [m,c]. */
int ft3d_segsum_rows(const float* src, const int32_t* entry_row, const float* entry_w, const int32_t* offsets,
                     const int32_t* cnt, int64_t m, int32_t c, float* out, ft3d_stream_t stream);
/* spf.calc_ti_weights: pc f32 [n,4], idx int64 [8,n] -> w_out f32 [8,n]. */
int ft3d_ti_weights(const float* pc, const int64_t* idx, int64_t n, float scale, float* w_out,
                    ft3d_stream_t stream);
/* Fused voxel_to_point map build (models/utils.py:71-85): for each point the 8 corner voxels of
 * its stride-`stride` cell looked up in the table, plus normalised trilinear weights.
 * idx_out int32 [n,8], w_out f32 [n,8]. */
int ft3d_v2p_build(const float* pc, int64_t n, int32_t stride, const uint64_t* table_keys,
                   const int32_t* table_vals, int64_t cap, int32_t* idx_out, float* w_out,
                   ft3d_stream_t stream);
/* Fused point_to_voxel map build (models/utils.py:46-53): idx_out[i] = voxel row of
 * floor(pc[i]/stride)*stride or -1; cnt_out int32 [m] = points per voxel. */
int ft3d_p2v_build(const float* pc, int64_t n, int32_t stride, const uint64_t* table_keys,
                   const int32_t* table_vals, int64_t cap, int32_t* idx_out, int32_t* cnt_out,
                   int64_t m, ft3d_stream_t stream);

/* ---- a15/K16  2D->3D lift (models/image_models_billinear.py:117-124) ----------------------- */
/* out[p,c] = fmap[b(p), c, row(p), col(p)]; fmap addressed by element strides (sb,sc,sh,sw) so
 * both NCHW and channels-last maps work; rc int32 [n,2] = (row,col); bidx int32 [n]. */
int ft3d_lift_fwd(const float* fmap, int64_t sb, int64_t sc, int64_t sh, int64_t sw, int32_t B,
                  int32_t C, int32_t H, int32_t W, const int32_t* rc, const int32_t* bidx,
                  int64_t n, float* out, ft3d_stream_t stream);
/* gmap (same strides) += scatter of gout; the caller zero-fills gmap. */
int ft3d_lift_bwd(const float* gout, int64_t sb, int64_t sc, int64_t sh, int64_t sw, int32_t B,
                  int32_t C, int32_t H, int32_t W, const int32_t* rc, const int32_t* bidx,
                  int64_t n, float* gmap, ft3d_stream_t stream);

/* ---- a8-a10/K11-K13  sparse convolution (all 49 spnn.Conv3d of models/spvcnn.py) ----------- */
/* Output-stationary gather-GEMM:  out[j,:] = sum_k  in[nbr[j,k],:] @ B_k   (rows with nbr<0 skipped)
 *   forward    : in = features,   nbr = map,             B_k = W[k]            (red=Cin,  n=Cout)
 *   dgrad      : in = grad_out,   nbr = transposed map,  B_k = W[k]^T          (red=Cout, n=Cin)
 *   k3 stride-1 dgrad reuses the forward table with kflip=1 (nbrT[i,k] == nbr[i,K-1-k]).
 * fp32 CUDA-core variant (exact-precision mode): w f32 [K,Cin,Cout]; w_transposed=1 uses W[k]^T. */
int ft3d_conv_gather_f32(const float* in, const int32_t* nbr, int64_t n_out, int32_t K,
                         int32_t kpad, int32_t kflip, int32_t red, int32_t ncols, const float* w,
                         int32_t w_transposed, float* out, ft3d_stream_t stream);
/* Weight images of the bf16 tcgen05 kernels (ft3d_conv_os, ft3d_conv_pairs_tc): operands rounded to bf16, fp32
 * accumulation in TMEM.  `wpacked` is the image written by ft3d_conv_pack_weights for (K, cin, cout, w_transposed):
 * per offset and 64-wide reduction block one [ncols x 128 B] swizzled bf16 tile, streamed with cp.async.bulk.
 * Supported: red % 16 == 0 (16..512), ncols % 32 == 0 (32..256, or 384). */
size_t ft3d_conv_packed_bytes(int32_t K, int32_t red, int32_t ncols);
int ft3d_conv_pack_weights(const float* w, int32_t K, int32_t cin, int32_t cout,
                           int32_t w_transposed, void* wpacked, ft3d_stream_t stream);
/* Every weight image of a model in one launch (they are re-packed after each optimizer step).  desc: device array of
 * n_desc records { const float* w; void* img; int32 K, cin, cout, w_transposed; int64 chunk_begin; } (40 bytes,
 * ft3d_conv_pack_desc_bytes()), chunk_begin = running sum of ft3d_conv_packed_bytes(...)/16 of the records before. */
size_t ft3d_conv_pack_desc_bytes(void);
int ft3d_conv_pack_weights_multi(const void* desc, int32_t n_desc, int64_t total_chunks, ft3d_stream_t stream);
/* wgrad: gw[k] += sum over pairs p of offset k of  a[pairs[p,ca],:]^T  b[pairs[p,cb],:]
 *   forward conv : a = features [.,cin], ca = 0 ; b = grad_out [.,cout], cb = 1
 *   transposed   : ca = 1, cb = 0 (pair columns swapped, models/spvcnn.py:42-46)
 * pair_offsets int32 [K+1] on device; max_pairs = host-side upper bound on L used to size the grid.
 * gw f32 [K,cin,cout] is accumulated into (caller zero-fills). */
int ft3d_conv_wgrad_f32(const float* a, const float* b, const int32_t* pairs,
                        const int32_t* pair_offsets, int32_t K, int32_t ca, int32_t cin,
                        int32_t cout, int64_t max_pairs, float* gw, ft3d_stream_t stream);
/* Same result without atomics: one thread per element of gw walks the pairs of its offset in list order
 * (bit-identical from run to run; slower). */
int ft3d_conv_wgrad_f32_det(const float* a, const float* b, const int32_t* pairs,
                            const int32_t* pair_offsets, int32_t K, int32_t ca, int32_t cin,
                            int32_t cout, int64_t max_pairs, float* gw, ft3d_stream_t stream);

/* Pair-major tensor-core path (the default): gather -> GEMM -> sorted, atomic-free scatter.
 *   ft3d_to_bf16        : activations / gradients rounded once to bf16 (dst holds n bf16).
 *   ft3d_conv_pairs_tc  : partial_out[p,:] = in_bf16[pairs[p][gather_col],:] @ B_k(p)  for every pair p (fp32 rows in
 *                         pair order; tcgen05.mma, fp32 accumulate in TMEM).  pairs == NULL (with K == 1) is the
 *                         identity gather over max_pairs rows: a dense GEMM (k = 1 convolutions, torchsparse
 *                         conv3d kernel_size 1 == F.matmul) whose partial_out IS the result.
 *   ft3d_conv_reduce    : out[row,:] = sum_j partial[ppos[row,j],:] over the row's compacted positions, i.e. in
 *                         ascending offset order (deterministic, no atomics, every output row written once).
 *                         `partial` must hold at least one row even for a map without pairs (absent slots re-read
 *                         row 0 and discard it).
 *   ft3d_conv_reduce_bn : the same pass also folds the per-channel sum / sum of squares of the rows it writes and
 *                         leaves this layer's BatchNorm training statistics in stat (see ft3d_bn_stats).
 *   forward: gather_col 0, ppos of ft3d_kmap_pairs;   dgrad / transposed conv: gather_col 1, ppos of
 *   ft3d_kmap_pair_positions(col 0) and the w_transposed weight image. */
int ft3d_to_bf16(const float* src, int64_t n, void* dst, ft3d_stream_t stream);
int ft3d_conv_pairs_tc(const void* in_bf16, const int32_t* pairs, const int32_t* pair_offsets, int32_t K,
                       int32_t gather_col, int64_t max_pairs, int32_t red, int32_t ncols,
                       const void* wpacked, float* partial_out, ft3d_stream_t stream);
int ft3d_conv_reduce(const float* partial, const int32_t* ppos, int64_t n_rows, int32_t kpad, int32_t ncols,
                     float* out, ft3d_stream_t stream);
int ft3d_conv_reduce_bn(const float* partial, const int32_t* ppos, int64_t n_rows, int32_t kpad, int32_t ncols,
                        float* out, float eps, float momentum, float* stat, float* running_mean,
                        float* running_var, const int32_t* valid_rows, void* workspace, size_t workspace_bytes,
                        ft3d_stream_t stream);
/* wgrad on bf16 inputs: gw[k] += a_bf16[pairs[p][ca],:]^T b_bf16[pairs[p][1-ca],:]; pairs == NULL: identity (K == 1). */
int ft3d_conv_wgrad_pairs_tc(const void* a_bf16, const void* b_bf16, const int32_t* pairs,
                             const int32_t* pair_offsets, int32_t K, int32_t ca, int32_t cin, int32_t cout,
                             int64_t max_pairs, float* gw, ft3d_stream_t stream);

/* Output-stationary tensor-core path (csrc/conv_os.cu, csrc/os_plan.cu): the forward, dgrad and transposed
 * convolutions of every tcgen05-shaped layer.  Replaces the torchsparse gather -> mm -> scatter_add loop
 *  behind spnn.Conv3d (models/spvcnn.py:22-78) with ONE launch that also produces the layer's
 * BatchNorm statistics.
 *   ft3d_conv_os_plan : schedule of one side of a kernel map.  table int32 [n_rows,kpad] is the neighbour table
 *       seen from the rows being PRODUCED (nbr of ft3d_kmap_build for a forward conv, its transpose for dgrad /
 *       transposed conv).  Rows are sorted by occupancy mask (rarest offsets most significant) into T =
 *       ceil(n_rows/tile_rows) tiles (tile_rows = 128, 256 or 512 = 128 x the CTAs of the thread-block cluster that
 *       will share a tile's weight blocks); the offsets that occur in a tile are its passes.  A tile with more than `cap`
 *       passes is split into ceil(passes/cap) work units over disjoint pass ranges (cap = chunk_passes, or, if 0,
 *       ~3/4 of the mean passes per CTA clamped to [4,8]); units are listed longest first.  Outputs:
 *       units_out int32 [unit_cap,8] = {first pass, passes, tile, units of the tile, index among them, first scratch
 *       slot of the tile, 0, 0}; out_row_out int32 [T*tile_rows] (row of each tile slot, -1 = empty); pass_k_out int32
 *       [pass_cap]; pass_idx_out int32 [pass_cap,tile_rows] (gather row per slot, -1 = none); split_tiles_out int32 [T,4] =
 *       {tile, its units, its first scratch slot, 0} for the tiles that were split; num_out int32 [8] = {P passes,
 *       U units, S scratch slots (units of split tiles), cap, split tiles, 0, 0, 0}.  P <= min(pairs, T*K), U <= 4T
 *       (a tile is split evenly into at most 4 units); entries beyond pass_cap / unit_cap are dropped, so the caller
 *       checks the counts.  workspace:
 *       ft3d_conv_os_plan_workspace(n_rows, unit_cap) bytes, 256-byte aligned.
 *   ft3d_conv_os : out[out_row[s],:] = sum over the passes of the slot's tile of in_bf16[pass_idx[p][s],:] @ B_k(p)
 *       accumulated in TMEM in ascending offset order; every output row is written once (fp32), no atomics; the
 *       units of a split tile leave partial tiles in scratch slots and a second, tiny launch (one CTA per split
 *       tile) adds them in unit order, so the result is bit-identical from launch to launch.  in_bf16 [n_in,red]; wpacked as for ft3d_conv_pairs_tc
 *       (ft3d_conv_pack_weights); kflip != 0 uses B_{K-1-k} (dgrad of a symmetric stride-1 map re-uses the forward
 *       schedule).  tile_rows > 128 launches clusters of tile_rows/128 CTAs that walk a tile's passes in lockstep: rank
 *       0 multicasts each weight block B_k to all of them (one L2 read per cluster instead of one per 128 rows).
 *       valid_rows (nullable): out holds n_out rows of which the first *valid_rows are real; the others
 *       are zero-filled.  stat != NULL: BatchNorm training statistics of the rows written (semantics of
 *       ft3d_bn_stats: stat [2,ncols] = mean, rstd; running stats updated) computed in the epilogue, deterministic
 *       (per-CTA partial rows folded in double by a third, tiny launch).  workspace (needed with statistics or
 *       scratch_slots > 0): ft3d_conv_os_workspace(ncols, scratch_slots, tile_rows) bytes, 256-byte aligned, private to the
 *       stream.  trace (nullable): 8 x uint64 per CTA of device timestamps (tools/conv_os_probe.py).
 *   Gathers are TMA (cp.async.bulk.tensor tile::gather4); FT3D_OS_GATHER=ldgsts selects 16-byte cp.async. */
size_t ft3d_conv_os_plan_workspace(int64_t n_rows, int64_t unit_cap);
int ft3d_conv_os_plan(const int32_t* table, int64_t n_rows, int32_t K, int32_t kpad, int32_t tile_rows,
                      int64_t pass_cap, int64_t unit_cap, int32_t chunk_passes, int32_t* units_out,
                      int32_t* split_tiles_out,
                      int32_t* out_row_out, int32_t* pass_k_out, int32_t* pass_idx_out, int32_t* num_out,
                      void* workspace, size_t workspace_bytes, ft3d_stream_t stream);
size_t ft3d_conv_os_workspace(int32_t ncols, int64_t scratch_slots, int32_t tile_rows);
int ft3d_conv_os(const void* in_bf16, int64_t n_in, const int32_t* units, const int32_t* split_tiles,
                 const int32_t* num, const int32_t* out_row, const int32_t* pass_k, const int32_t* pass_idx,
                 int64_t unit_cap, int64_t tiles, int32_t tile_rows, int64_t scratch_slots, int32_t K, int32_t kflip,
                 int32_t red, int32_t ncols, const void* wpacked, float* out, int64_t n_out, const int32_t* valid_rows, float eps,
                 float momentum, float* stat, float* running_mean, float* running_var, void* workspace,
                 size_t workspace_bytes, void* trace, ft3d_stream_t stream);

/* Deterministic weight gradient (two-stage split-K, no atomics): every item of <= 512 pairs of one offset writes its
 * [128 x cout] block per Cin block into `workspace`, a second launch adds the blocks of each offset in item order and
 * writes gw (accumulate != 0: adds the sum to what gw holds -- a gradient arena).  Bit-identical from run to run.
 * workspace: ft3d_conv_wgrad_det_workspace(K, cin, cout, max_pairs, pairs != NULL) bytes, 256-byte aligned. */
size_t ft3d_conv_wgrad_det_workspace(int32_t K, int32_t cin, int32_t cout, int64_t max_pairs, int32_t has_pairs);
int ft3d_conv_wgrad_pairs_tc_det(const void* a_bf16, const void* b_bf16, const int32_t* pairs,
                                 const int32_t* pair_offsets, int32_t K, int32_t ca, int32_t cin, int32_t cout,
                                 int64_t max_pairs, float* gw, int32_t accumulate, void* workspace,
                                 size_t workspace_bytes, ft3d_stream_t stream);

/* ---- a11  spnn.BatchNorm / spnn.ReLU / residual add fused around the convolution
 *          (models/spvcnn.py:26-31,42-47,57-78; torchsparse BatchNorm == nn.BatchNorm1d on .F) ---------------- */
/* Column reductions are deterministic: per-CTA partial rows in `workspace` (ft3d_bn_workspace(C) bytes), folded in
 * double by a second tiny launch with a fixed-shape tree; no float atomics.  stat f32 [2,C] = (batch mean,
 * 1/sqrt(biased var + eps)); running_mean/var (nullable pair) are updated with `momentum` and the unbiased variance
 * exactly as nn.BatchNorm1d does in training mode.
 * valid_rows (nullable, device int32): the activation is padded to a static capacity of n rows of which only the
 * first *valid_rows are real -- statistics divide by that count and the apply kernels keep the padding rows exactly
 * zero, so a whole training step can be replayed as one CUDA graph for batches of different size (graph.py). */
size_t ft3d_bn_workspace(int32_t channels);
int ft3d_bn_stats(const float* y, int64_t n, int32_t channels, float eps, float momentum, float* stat,
                  float* running_mean, float* running_var, const int32_t* valid_rows, void* workspace,
                  size_t workspace_bytes, ft3d_stream_t stream);
/* out[c] (+)= sum over rows of x[:, c]  (deterministic two-stage sum in double; bias gradient of the point-branch
 * nn.Linear layers, models/spvcnn.py:164-180).  workspace: ft3d_bn_workspace(channels) bytes. */
int ft3d_col_sum(const float* x, int64_t n, int32_t channels, float* out, int32_t accumulate, void* workspace,
                 size_t workspace_bytes, ft3d_stream_t stream);
/* z = [relu]((y - mean) * rstd * gamma + beta [+ res]); writes z f32 [n,C] (nullable) and z16 bf16 [n,C]
 * (nullable) -- the copy the next convolution gathers.  Evaluation mode: pass stat built from the running stats. */
int ft3d_bn_apply(const float* y, int64_t n, int32_t channels, const float* stat, const float* gamma,
                  const float* beta, const float* res, int32_t relu, float* z, void* z16, int64_t ldz,
                  const int32_t* valid_rows, ft3d_stream_t stream);
/* torchsparse.cat([deconv(y), skip]) (models/spvcnn.py:212,216,224,228) without a concatenation pass: ft3d_bn_apply
 * writes z / z16 with row pitch ldz (0 = channels) straight into the left columns of the [n, C1+C2] buffers, and
 * ft3d_copy_cols puts the skip tensor (fp32 and its bf16 copy; src16 NULL: rounded here) at column col0 of the same
 * buffers.  The backward kernels read gz and the saved output with the same pitch (ldg). */
int ft3d_copy_cols(const float* src, const void* src16, int64_t n, int32_t c, float* dst, void* dst16, int64_t ld,
                   int32_t col0, ft3d_stream_t stream);
/* g' = gz * [z > 0] (mask from z16 if given, else z, else none).  red f32 [2,C] = (sum g'/n, sum g' xhat/n);
 * dgamma = sum g' xhat; dbeta = sum g' (accumulate != 0: added to the values already there, so the gradients can
 * be written straight into a flat gradient arena). */
int ft3d_bn_bwd_reduce(const float* gz, const float* y, const void* z16, const float* z, int64_t n,
                       int32_t channels, const float* stat, float* red, float* dgamma, float* dbeta,
                       int32_t accumulate, int64_t ldg, const int32_t* valid_rows, void* workspace,
                       size_t workspace_bytes, ft3d_stream_t stream);
/* gy = gamma rstd (g' - c1 - xhat c2)  (red == NULL: frozen statistics, gy = gamma rstd g'); outputs (each
 * nullable): gy f32, gy16 bf16 (the dgrad / wgrad operand), gres f32 = g' (gradient of the residual input). */
int ft3d_bn_bwd_apply(const float* gz, const float* y, const void* z16, const float* z, int64_t n,
                      int32_t channels, const float* stat, const float* gamma, const float* red, float* gy,
                      void* gy16, float* gres, int64_t ldg, const int32_t* valid_rows, ft3d_stream_t stream);

/* ---- segmentation loss and metric on the device ----------------------------------------------
 * ft3d_seg_loss: modules/SemanticTorchpackTrainer.py:70-108.  loss = (1-l) * CE + l * KL with
 *   CE = F.cross_entropy(logits, labels, weight=class_weight, ignore_index)   (weighted mean over live rows)
 *   KL = F.kl_div(log_softmax(logits), softmax(teacher), 'none').sum(1).mean()   (teacher = the other modality's
 *        detached logits; nullable with lambda_xm == 0)
 * and its gradient with respect to the logits, in one pass.  logits/teacher/grad_out f32 [n, C] (C <= 64), labels
 * int64 [n], loss_out f32 [4] = {loss, CE, KL, sum of weights}.  valid_rows (nullable): device int32 holding the
 * real row count when n is a padded capacity.  workspace: ft3d_seg_loss_workspace() bytes, 8-byte aligned.
 * ft3d_confusion_update: models/metric.py:37-58 (SegIoU.update_dict): mat[label, argmax(logits)] += 1 for
 * labels != ignore_index; mat int64 [C, C], accumulated in place. */
size_t ft3d_seg_loss_workspace(void);
int ft3d_seg_loss(const float* logits, const int64_t* labels, int64_t n, int32_t num_classes, int64_t ignore_index,
                  const float* class_weight, const float* teacher_logits, float lambda_xm, const int32_t* valid_rows,
                  float* loss_out, float* grad_out, void* workspace, size_t workspace_bytes, ft3d_stream_t stream);
int ft3d_confusion_update(const float* logits, const int64_t* labels, int64_t n, int32_t num_classes,
                          int64_t ignore_index, const int32_t* valid_rows, int64_t* mat, ft3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FT3D_H_ */
