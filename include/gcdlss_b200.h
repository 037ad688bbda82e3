/*
 * gcdlss_b200.h — C ABI of libgcdlss_sm100a.so
 *
 * The sm_100a (B200) implementation of the one data-parallel hot path of GCDLSS:
 * point->voxel quantisation, coordinate / kernel-map construction, sparse convolution
 * (forward, dgrad, wgrad), batch-norm(+ReLU,+residual) and voxel<->point gathers/reductions.
 *
 * What each group replaces in the reference (the reference has no FFI of its own: its hot
 * path calls MinkowskiEngine / mmcv native extensions from Python, so "the entry points the
 * reference's FFI for this path would bind" are the calls those Python sites make):
 *
 *   gcd_quantize_*, gcd_unique_*   ME.utils.sparse_quantize             utils/dataset_remission.py:868-873,
 *                                                                      modules/exp_merge_mean_teacher.py:2856-2861
 *                                  Voxelizer.voxelize / ravel_hash /    models/voxelizer.py:271-302, 312-360
 *                                  sparse_quantize (np.unique flavour)
 *   gcd_hash_build                 ME.SparseTensor(coordinates=...)     modules/exp.py:259, exp_merge_mean_teacher.py:2802
 *   gcd_coords_stride2,            ME CoordinateManager stride / kernel models/minkunet.py:62-128 (every MinkowskiConvolution /
 *   gcd_kmap_*                     map generation                       MinkowskiConvolutionTranspose call)
 *   gcd_conv_*                     MinkowskiConvolution(Transpose)      models/minkunet.py:140-214, resnet_block BasicBlock
 *                                  forward / backward
 *   gcd_bn_*                       ME.MinkowskiBatchNorm + MinkowskiReLU models/minkunet.py:65,130 (+ residual add of BasicBlock)
 *   gcd_rows_gather, gcd_csr_*,    voxel->point devoxelisation gather   models/decoder.py:416-424,
 *   gcd_segment_*                  and its segmented-sum backward;      modules/exp_merge_mean_teacher.py:2845-2846;
 *                                  point->voxel mean/max reduce         models/encoder.py:121-164 (mmcv DynamicScatter)
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no C++ or torch types cross the boundary.
 *   - Every pointer is a DEVICE pointer unless the name ends in _host.  The caller owns all
 *     memory, including workspaces and hash tables; the library never allocates, frees or keeps
 *     a pointer past the call.  All work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - Return value: 0 on success, a negative gcd_status otherwise; gcd_last_error_string()
 *     gives the text for the calling thread.  Conditions only detectable on the device (key
 *     range overflow, duplicate coordinates, hash table full) are reported through the
 *     caller-provided `status` word (int32, device) that the caller must zero beforehand and
 *     reads back when it next synchronises; bits are gcd_dev_status.
 *   - Row-major matrices with an explicit leading dimension where one is given.
 *   - Voxel coordinates are int32 (b, x, y, z).  Hash keys pack them into 64 bits:
 *     10 bits batch | 3 x 18 bits (coordinate + 2^17); anything outside sets GCD_DEV_KEY_RANGE.
 *   - Kernel maps are dense neighbour tables stored column-major: nbr[k * n_out + o] is the
 *     input row feeding output row o through kernel offset k, or -1.  Offset order: x fastest,
 *     odd K centred, even K one-sided (MinkowskiEngine convention).
 */
#ifndef GCDLSS_B200_H_
#define GCDLSS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  GCD_OK = 0,
  GCD_ERR_INVALID_ARG = -1,
  GCD_ERR_CUDA = -2,
  GCD_ERR_UNSUPPORTED = -3,
  GCD_ERR_WORKSPACE = -4
} gcd_status;

typedef enum {
  GCD_DEV_KEY_RANGE = 1,  /* a coordinate does not fit the 64-bit key */
  GCD_DEV_DUPLICATE = 2,  /* gcd_hash_build saw the same coordinate twice */
  GCD_DEV_TABLE_FULL = 4  /* probing wrapped around the whole table */
} gcd_dev_status;

typedef enum { GCD_ROUND_FLOOR = 0, GCD_ROUND_HALF_EVEN = 1 } gcd_round_mode;
typedef enum { GCD_F32 = 0, GCD_BF16 = 1 } gcd_dtype;
typedef enum { GCD_MATH_FP32_SIMT = 0, GCD_MATH_BF16_TCGEN05 = 1 } gcd_math_mode;

const char* gcd_last_error_string(void);
int32_t gcd_abi_version(void);
/* 1 when the library was built with the tcgen05 kernels (always, for sm_100a). */
int32_t gcd_has_tcgen05(void);

/* Process-wide tuning options.  They select between implementations that produce IDENTICAL results (bit-exact for the
 * integer kernels); nothing here changes what a call computes.  Values are plain atomics: reads and writes from any
 * thread are safe, a call in flight sees either the old or the new value.  Initial values come from the environment
 * variable of the same name (GCD_PAIRS_FUSED, ...) read once, at the first use of the library, then never again. */
typedef enum {
  GCD_OPT_PAIRS_FUSED = 0,   /* gcd_pairs_from_table: 2 (default) one pass over the table (decoupled look-back); 1: two passes
                                (count, scan of the tile counts, emit); 0: flag / scan / emit.  Identical lists. */
  GCD_OPT_GATHER_FLAT = 1,   /* 1 (default): gcd_rows_gather deals float4 elements flat; 0: one warp per row */
  GCD_OPT_TC_STAGES = 2,     /* > 0: cap on the shared-memory ring depth of the tcgen05 convolution (tuning aid) */
  GCD_OPT_TC_GROUP = 3,      /* 1 | 2 | 4 | 8: gather warps per ring slot; 0 = chosen per launch (tuning aid) */
  GCD_OPT_WG_CHUNK_MIN = 4,  /* > 0: minimum pairs per wgrad work item (tuning aid) */
  GCD_OPT_TC_WARPS = 5,      /* 8 | 16: gather warps per CTA of the tcgen05 forward / dgrad kernel */
  GCD_OPT_BN_FUSED = 6,      /* 1 (default): the batch norms inside gcd_block_* run as one two-phase cooperative launch per
                                direction (reduction, grid barrier, elementwise pass); 0: two launches */
  GCD_OPT_KMAP_COOP = 7,     /* 1: gcd_kmap_subm probes warp-cooperatively (four lanes per voxel, one 32-byte sector per probe);
                                0 (default): one thread per voxel.  Identical tables. */
  GCD_OPT_PDL = 8,           /* 1 (default): the kernels of the training step (convolutions, batch norms, adds / copies) are
                                launched with programmatic stream serialization: kernel i + 1 is scheduled and runs its set-up
                                while kernel i drains, and waits (griddepcontrol.wait) before touching global memory; 0: plain
                                launches.  Results are identical. */
  GCD_OPT_WGRAD_SIDE = 9,    /* gcd_run_ops_exec: 1 (default) weight gradients on the context's second stream, ordered after the
                                block's batch-norm backward (they may run beside the input-gradient kernel); 2: ordered after the
                                input-gradient kernel; 0: everything on the caller's stream */
  GCD_OPT_DYN_TILES = 10,    /* gcd_run_ops_exec: 1: dynamic tile schedule of the tcgen05 kernels (heaviest tiles first); 2: the same in
                                table order; 0 (default): static striding (measured faster on an otherwise idle GPU, r2 call 9) */
  GCD_OPT_DYN_AHEAD = 11,    /* > 0: tiles a CTA of the forward / dgrad kernel may claim ahead under the dynamic schedule (1..8, tuning aid) */
  GCD_OPT_BN_MASK_FROM_X = 12, /* 1 (default): the fused batch-norm backward of conv -> BN -> ReLU units inside gcd_block_backward re-derives
                                  the ReLU mask from the convolution output (x * scale + shift > 0, the forward pass's own arithmetic)
                                  and does not read the stored activation; 0: reads it */
  GCD_OPT_SCAN_LOOKBACK = 13, /* 1 (default): the device-wide exclusive scans behind gcd_unique_rows / gcd_coords_stride2 run as one
                                 decoupled-look-back launch; 0: three launches (tile scan, scan of the tile sums, add) */
  GCD_OPT_COUNT_ = 14
} gcd_option;
int32_t gcd_set_option(int32_t option, int32_t value);
int32_t gcd_get_option(int32_t option);

/* ------------------------------------------------------------------ quantisation -------- */
/* out[i, d] = (int32) round_mode( pts[i*ld + d] / q ), d < dims (dims = 3 or 4).  The division
 * is an IEEE division in the input precision (never a multiply by a reciprocal). */
int32_t gcd_quantize_f32(const float* pts, int64_t ld, int64_t n, int32_t dims, float q,
                         int32_t round_mode, int32_t* out, void* stream);
int32_t gcd_quantize_f64(const double* pts, int64_t ld, int64_t n, int32_t dims, double q,
                         int32_t round_mode, int32_t* out, void* stream);
/* Rigid / voxelisation transform of the dataset's augmentation step in float64, as numpy computes it from float32 points
 * and a float64 4x4 matrix (ref utils/dataset_remission.py:821-833: homo_coords @ rigid_transformation.T[:, :3]):
 *   out[i, j] = ((p[i,0]*m[j][0] + p[i,1]*m[j][1]) + p[i,2]*m[j][2]) + m[j][3]   (fused multiply-adds in that order)
 * m_host: HOST pointer to the 12 doubles of the first three rows of the matrix (row-major).  out [n, 3] float64 feeds
 * gcd_quantize_f64. */
int32_t gcd_affine_f64(const float* pts, int64_t ld, int64_t n, const double* m_host, double* out, void* stream);
/* Column-wise minimum of an int32 [n, dims] matrix (dims <= 4); mins must be pre-filled with
 * INT32_MAX.  Used for the `res_coors -= res_coors.min(0)` shift of models/voxelizer.py:276. */
int32_t gcd_colmin_i32(const int32_t* coords, int64_t n, int32_t dims, int32_t* mins, void* stream);
int32_t gcd_sub_cols_i32(int32_t* coords, int64_t n, int32_t dims, const int32_t* mins, void* stream);

/* Hash table capacity (slots, a power of two >= 2n) and workspace sizes. */
int64_t gcd_hash_capacity(int64_t n);
size_t gcd_unique_workspace_bytes(int64_t n);

/* Unique rows of an int32 [n, dims] matrix (dims 3: (x,y,z), batch taken as 0; dims 4: (b,x,y,z)).
 * order = 0: first-occurrence order (ME.utils.sparse_quantize): unique_idx ascending.
 * order = 1: ascending (b,x,y,z) key order (np.unique on ravel_hash, models/voxelizer.py:334-360);
 *            unique_idx[j] = first point of the j-th smallest voxel.
 * Outputs: unique_idx [<= n] int64, inverse [n] int64, m_out (device int32) = number of voxels,
 * and the table (keys/vals, `cap` slots) mapping voxel key -> voxel index on return. */
int32_t gcd_unique_rows(const int32_t* coords, int64_t n, int32_t dims, int32_t order,
                        uint64_t* table_keys, int32_t* table_vals, int64_t cap,
                        int64_t* unique_idx, int64_t* inverse, int32_t* m_out,
                        void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* ------------------------------------------------------------------ coordinate maps ----- */
/* Insert n unique (b,x,y,z) rows; table maps key -> row index. */
int32_t gcd_hash_build(const int32_t* coords, int64_t n, uint64_t* table_keys, int32_t* table_vals,
                       int64_t cap, int32_t* status, void* stream);

size_t gcd_stride2_workspace_bytes(int64_t n);
/* Coarse map of a stride-2 convolution at tensor stride ts (output stride 2 ts):
 * coarse[parent[f]] = floor(coords[f] / 2ts) * 2ts, coarse voxels numbered in first-occurrence
 * order of their children; code[f] = dx + 2 dy + 4 dz, d = (c - coarse)/ts.
 * The coarse table (cap_coarse slots) maps coarse key -> coarse row on return. */
int32_t gcd_coords_stride2(const int32_t* coords, int64_t n, int32_t ts,
                           uint64_t* coarse_keys, int32_t* coarse_vals, int64_t cap_coarse,
                           int32_t* coarse_coords, int32_t* parent, int32_t* code, int32_t* m_out,
                           void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* Stride-1 kernel map for kernel_size 3 or 5 at tensor stride ts: nbr [K^3][n] column-major. */
int32_t gcd_kmap_subm(const int32_t* coords, int64_t n, const uint64_t* table_keys,
                      const int32_t* table_vals, int64_t cap, int32_t kernel_size, int32_t ts,
                      int32_t* nbr, void* stream);
/* Stride-2 K=2 maps from parent/code: down: nbr [8][n_coarse] (child rows);
 * up (transposed conv): nbr [8][n_fine] with the single entry nbr[code[f]][f] = parent[f]. */
int32_t gcd_kmap_down2(const int32_t* parent, const int32_t* code, int64_t n_fine, int64_t n_coarse,
                       int32_t* nbr, void* stream);
int32_t gcd_kmap_up2(const int32_t* parent, const int32_t* code, int64_t n_fine, int32_t* nbr, void* stream);

/* Run table (opt-in, GCDLSS_KMAP=runs): the same stride-1 kernel maps from a hash whose 32-byte slots hold runs of
 * four x-adjacent cells of one (b, y, z) row, so the K neighbours of a row cost one or two sector loads instead of K
 * key loads + K value loads (csrc/runtable.cuh).  Replaces, for this purpose, the table ME builds per coordinate map
 * (ME.SparseTensor(coordinates=...), modules/exp.py:259) and the per-offset search behind every stride-1
 * MinkowskiConvolution (models/minkunet.py:62-128).
 * slots: caller-owned, 32-byte aligned, cap * gcd_runtable_slot_bytes() bytes, cap = gcd_hash_capacity(n).
 * gcd_runtable_build: inserts n unique rows whose x is a multiple of ts (status: key range / duplicate / full).
 * gcd_kmap_subm_runs: nbr [K^3][n] column-major, bit-identical to gcd_kmap_subm. */
size_t gcd_runtable_slot_bytes(void);
int32_t gcd_runtable_build(const int32_t* coords, int64_t n, int32_t ts, void* slots, int64_t cap,
                           int32_t* status, void* stream);
int32_t gcd_kmap_subm_runs(const int32_t* coords, int64_t n, const void* slots, int64_t cap,
                           int32_t kernel_size, int32_t ts, int32_t* nbr, void* stream);

/* Tile sort of a 3x3x3 (kv = 27) or 2x2x2 (kv = 8) table for the tcgen05 convolution (opt-in, GCDLSS_TILE_SORT=1):
 * columns sorted (stably) by the mask of present neighbours (3x3x3: rarest offsets in the top bits), so that the
 * 128-column tiles of the kernel see 8-12 offsets with a hit instead of 21-25 (2x2x2: 1-4 instead of 7-8)
 * (csrc/tilesort.cuh).  nbr_sorted [kv][n] = nbr[:, out_rows], out_rows [n] the permutation; pass both to
 * gcd_conv_forward (nbr = nbr_sorted, out_rows).  Pair lists are still built from nbr.
 * tile_masks (optional, [ceil(n / 128)] uint32): bit k of entry t is set iff some column of tile t of nbr_sorted has a
 * neighbour at offset k; gcd_conv_args.tile_masks lets the kernel stage only those slices of the table. */
size_t gcd_tile_sort_workspace_bytes(int64_t n);
int32_t gcd_kmap_tile_sort(const int32_t* nbr, int64_t n, int32_t kv, int32_t* nbr_sorted, int32_t* out_rows,
                           uint32_t* tile_masks, void* workspace, size_t workspace_bytes, void* stream);

size_t gcd_pairs_workspace_bytes(int64_t n_out, int32_t kv);
/* Per-offset pair lists of a table: pairs of offset k are [pair_off[k], pair_off[k+1]), sorted by
 * output row.  pair_in / pair_out hold up to n_out*kv entries; pair_off is int32 [kv+1] (device). */
int32_t gcd_pairs_from_table(const int32_t* nbr, int64_t n_out, int32_t kv, int32_t* pair_in,
                             int32_t* pair_out, int32_t* pair_off, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ convolution --------- */
/* Weight operand description shared by forward and dgrad:
 *   B_k[c, j] = w[wsel(k) * w_stride_k + c * w_stride_c + j * w_stride_n],  wsel(k) = mirror ? kv-1-k : k
 * forward:  c over Cin, j over Cout of kernel [kv, Cin, Cout]:  strides (Cin*Cout, Cout, 1), mirror 0
 * dgrad of a stride-1 conv: same table, mirror 1, strides (Cin*Cout, 1, Cout) (i.e. W^T)
 * dgrad of down / up convs: the opposite table (up / down), mirror 0, transposed strides. */
typedef struct {
  const void* in;        /* [n_in, c_in] features (dtype in_dtype), leading dimension ld_in */
  int64_t ld_in;
  int64_t n_in;
  const int32_t* nbr;    /* [kv][n_out] table, or NULL = identity (kv must be 1, n_in == n_out) */
  int32_t kv;
  int64_t n_out;
  int32_t c_in;
  int32_t c_out;
  const float* w;        /* fp32 master weights (SIMT path) */
  const void* w_packed;  /* bf16 operand images made by gcd_conv_pack_weights (tcgen05 path) */
  int64_t w_stride_k, w_stride_c, w_stride_n;
  int32_t mirror;
  const float* bias;     /* [c_out] or NULL */
  void* out;             /* [n_out, c_out] (dtype out_dtype), leading dimension ld_out */
  int64_t ld_out;
  int32_t in_dtype;      /* gcd_dtype */
  int32_t out_dtype;     /* gcd_dtype */
  double* stats;         /* optional [2*c_out] fp64: per-channel sum and sum of squares of the result,
                            accumulated atomically (caller zeroes), or NULL.  SIMT path: from the fp32
                            accumulator in the epilogue; tcgen05 path: a gcd_bn_stats pass over the
                            stored tensor issued by the same call */
  int32_t math_mode;     /* gcd_math_mode */
  const int32_t* out_rows; /* optional (tcgen05 path only): nbr is a tile-sorted table (gcd_kmap_tile_sort) and the
                              result of table column i is written to row out_rows[i] of out; NULL = row i */
  const uint32_t* tile_masks; /* optional (tcgen05 path only): per 128-column tile of nbr, the mask of offsets with a hit
                                 (gcd_kmap_tile_sort); NULL = the kernel derives it from the whole table slice */
  int32_t* sched;          /* optional (tcgen05 path only): two device int32, ZERO on entry and zero again when the kernel has
                              finished: the CTAs claim row tiles from this counter (heaviest first) instead of striding over
                              them, so a CTA that starts late because another kernel holds its SM steals no one's time.  One
                              pair per stream in flight (launches of one stream may share it); NULL = static striding */
} gcd_conv_args;

int32_t gcd_conv_forward(const gcd_conv_args* args, void* stream);

/* bf16 operand images for the tcgen05 path.  transpose = 0: forward operand of kernel
 * [kv, c_in, c_out]; transpose = 1: dgrad operand (W[k]^T, offsets optionally mirrored). */
size_t gcd_conv_packed_weight_bytes(int32_t kv, int32_t c_in, int32_t c_out);
int32_t gcd_conv_pack_weights(const float* w, int32_t kv, int32_t c_in, int32_t c_out,
                              int32_t transpose, int32_t mirror, void* packed, void* stream);

/* dW[k] += in[pair_in]^T · gout[pair_out] over the pairs of each offset; dw is fp32 [kv, c_in, c_out]
 * and is ACCUMULATED into (caller zeroes).  pair_in == NULL means identity (1x1 conv, kv == 1,
 * n_pairs rows). */
typedef struct {
  const void* in;  int64_t ld_in;       /* [n_in, c_in] */
  const void* gout; int64_t ld_gout;    /* [n_out, c_out] */
  const int32_t* pair_in; const int32_t* pair_out; const int32_t* pair_off; /* device */
  int64_t n_pairs;                      /* identity: number of rows; pair lists: an upper bound of the
                                           total pair count (capacity of pair_in), the exact per-offset
                                           counts are read from pair_off on the device */
  int32_t kv, c_in, c_out;
  float* dw;
  float* dbias;                         /* optional [c_out]: += column sums of gout (caller zeroes) */
  int64_t n_out;                        /* rows of gout (for dbias) */
  int32_t in_dtype, gout_dtype, math_mode;
  int32_t* sched;                       /* optional (tcgen05 path): as gcd_conv_args.sched, for the kernel's work items */
} gcd_wgrad_args;

int32_t gcd_conv_wgrad(const gcd_wgrad_args* args, void* stream);

/* Re-pack many kernels in one launch (after an optimiser step).  descs: DEVICE array of n_descs
 * gcd_pack_desc; descriptor i is handled by thread blocks [block_start_i, block_start_{i+1}) of 256
 * threads, one thread per element of its image; total_blocks = sum of ceil(image elements / 256). */
typedef struct {
  const float* w;        /* fp32 kernel [kv, c_in, c_out] */
  void* dst;             /* bf16 image, gcd_conv_packed_weight_bytes(kv, c_in, c_out) bytes */
  int32_t kv, c_in, c_out, transpose, mirror;
  int32_t block_start;
} gcd_pack_desc;
int32_t gcd_conv_pack_weights_batched(const void* descs, int32_t n_descs, int32_t total_blocks, void* stream);

/* 1 when (c_in, c_out, kv) can run on the tcgen05 path (both multiples of 16, c_out <= 512, kv <= 27). */
int32_t gcd_conv_tc_supported(int32_t c_in, int32_t c_out, int32_t kv);

/* Explicit im2col for thin inputs (the 5x5x5, Cin = 1 stem): out[o, k*c_in + c] = in[nbr[k][o], c],
 * zero where the table holds -1 and in the padding columns up to ld_out.  The stem then runs as
 * a dense [n, ld_out] x [ld_out, c_out] product through gcd_conv_forward with an identity map. */
int32_t gcd_im2col(const void* in, int64_t ld_in, int32_t c_in, const int32_t* nbr, int32_t kv, int64_t n_out,
                   void* out, int64_t ld_out, int32_t in_dtype, int32_t out_dtype, void* stream);

/* ------------------------------------------------------------------ batch norm ---------- */
/* Per-channel sum / sum of squares of x [n, c] into stats [2c] fp64 (accumulated; caller zeroes). */
int32_t gcd_bn_stats(const void* x, int64_t ld, int64_t n, int32_t c, int32_t dtype, double* stats, void* stream);
/* Training-mode finalise: mean/invstd [c] fp32 from stats, running stats update
 * (momentum, unbiased variance as nn.BatchNorm1d), scale = gamma*invstd, shift = beta - mean*scale. */
int32_t gcd_bn_finalize(const double* stats, int64_t n, int32_t c, const float* gamma, const float* beta,
                        float eps, float momentum, float* running_mean, float* running_var,
                        float* mean, float* invstd, float* scale, float* shift, void* stream);
/* Training-mode finalise + apply in ONE launch: y = act((x - mean) * invstd * gamma + beta (+ residual)) with mean / invstd
 * derived from stats; also writes mean / invstd [c] for the backward pass and updates the running statistics. */
int32_t gcd_bn_apply_train(const void* x, int64_t ld_x, int64_t n, int32_t c, const double* stats, const float* gamma,
                           const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                           float* mean, float* invstd, const void* residual, int64_t ld_res, int32_t relu, void* y,
                           int64_t ld_y, int32_t dtype, void* stream);
/* Eval-mode: scale/shift from running stats. */
int32_t gcd_bn_fold_eval(int32_t c, const float* gamma, const float* beta, const float* running_mean,
                         const float* running_var, float eps, float* scale, float* shift, void* stream);
/* y = act(x * scale + shift (+ residual)); relu = 0/1; residual may be NULL; y may alias x. */
int32_t gcd_bn_apply(const void* x, int64_t ld_x, int64_t n, int32_t c, const float* scale, const float* shift,
                     const void* residual, int64_t ld_res, int32_t relu, void* y, int64_t ld_y,
                     int32_t dtype, void* stream);
/* Backward of y = act(bn(x) + residual) in training mode.
 *  pass 1 (reduce): g = dy * (y > 0 if relu); sums[0:c] += sum g, sums[c:2c] += sum g * xhat  (fp64)
 *  pass 2 (apply):  dx = gamma*invstd * (g - sum_g/n - xhat * sum_gxhat/n); dres = g (optional);
 *                   dgamma = sum_gxhat, dbeta = sum_g (accumulated into dgamma/dbeta, fp32). */
int32_t gcd_bn_backward_reduce(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
                               int64_t n, int32_t c, const float* mean, const float* invstd, int32_t relu,
                               int32_t dtype, double* sums, void* stream);
int32_t gcd_bn_backward_apply(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
                              int64_t n, int32_t c, const float* mean, const float* invstd, const float* gamma,
                              const double* sums, int32_t relu, int32_t training, void* dx, int64_t ld_dx,
                              void* dres, int64_t ld_dres, float* dgamma, float* dbeta, int32_t dtype, void* stream);
/* y = max(x, 0) and its backward, for stand-alone MinkowskiReLU. */
int32_t gcd_relu(const void* x, void* y, int64_t numel, int32_t dtype, void* stream);
int32_t gcd_relu_backward(const void* dy, const void* y, void* dx, int64_t numel, int32_t dtype, void* stream);

/* ------------------------------------------------------------------ voxel <-> point ----- */
/* out[p, :] = in[idx[p], :]  (devoxelisation gather; idx int64). */
int32_t gcd_rows_gather(const float* in, int64_t ld_in, const int64_t* idx, int64_t n_out, int32_t c,
                        float* out, int64_t ld_out, void* stream);
size_t gcd_csr_workspace_bytes(int64_t n_points, int64_t n_segments);
/* Group points by segment (counting sort, stable): seg_off [n_segments+1] int32, order [n_points] int32
 * lists the points of segment s at order[seg_off[s]:seg_off[s+1]] in ascending point index. */
int32_t gcd_csr_build(const int64_t* idx, int64_t n_points, int64_t n_segments, int32_t* seg_off,
                      int32_t* order, void* workspace, size_t workspace_bytes, void* stream);
/* out[s, :] = reduce over the points of segment s of in[p, :]; mode 0 sum, 1 mean, 2 max
 * (empty segments give 0).  The backward of gcd_rows_gather is mode 0. */
int32_t gcd_segment_reduce(const float* in, int64_t ld_in, const int32_t* seg_off, const int32_t* order,
                           int64_t n_segments, int32_t c, int32_t mode, float* out, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------ loss side ----------- */
/* Teacher/student consistency terms of the Stage-2 step on fp32 voxel logits [n, c], one pass (SURVEY 8(f) rank 3;
 * replaces F.softmax x2 + F.mse_loss + torch.max, ref modules/exp_merge_mean_teacher.py:2832-2850):
 *   sq_err[i] = sum_c (softmax(logits_s[i])_c - softmax(logits_t[i])_c)^2      (mse_loss = sum_i sq_err[i] / (n c))
 *   max_prob[i], label[i] = max / first argmax of softmax(logits_t[i]); label = -1 where max_prob < threshold (> 0)
 *   grad (optional, [n, c]) = d sq_err[i] / d logits_s[i, :]. */
int32_t gcd_consistency_rows(const float* logits_s, int64_t ld_s, const float* logits_t, int64_t ld_t, int64_t n, int32_t c,
                             float threshold, float* sq_err, float* max_prob, int64_t* label, float* grad, int64_t ld_g,
                             void* stream);

/* ------------------------------------------------------------------ fused blocks -------- */
/* conv -> batch norm (-> ReLU) triples and whole residual blocks sequenced from C: one call per block and
 * direction instead of one per launch (the step is launch-bound from Python).  Replaces, as one unit,
 * MinkowskiEngine's BasicBlock.forward / its autograd backward (MinkowskiEngine/modules/resnet_block.py,
 * imported at ref models/minkunet.py:30) and the conv-bn-relu triples of ref models/minkunet.py:140-147.
 * Training mode only (batch statistics); all pointers device, caller allocated.  Activations of a block are
 * dense row-major [n, c] of one dtype. */
typedef struct {
  const int32_t* nbr;        /* forward table [kv][n_out], NULL = identity (kv == 1) */
  int64_t n_in, n_out;
  const int32_t* back_nbr;   /* table of the transposed map [kv][n_in] (dgrad) */
  int32_t back_mirror;       /* 1: dgrad uses W[kv-1-k]^T (self-transposed stride-1 maps) */
  const int32_t* pair_in; const int32_t* pair_out; const int32_t* pair_off;   /* wgrad pair lists (NULL = identity) */
  int64_t n_pairs;
  int32_t kv, c_in, c_out;
  const float* w;            /* fp32 [kv, c_in, c_out] */
  const void* w_packed_fwd;  /* tcgen05 operand images; NULL selects the fp32 SIMT kernels */
  const void* w_packed_bwd;
  const float* gamma; const float* beta; float* running_mean; float* running_var;
  float eps, momentum;
  double* stats;             /* [2 c_out] zero on entry of gcd_block_forward  */
  double* sums;              /* [2 c_out] zero on entry of gcd_block_backward */
  float* mean; float* invstd;                /* [c_out] each: written by forward, read by backward */
  float* dw; float* dgamma; float* dbeta;    /* zero on entry of backward, accumulated into */
  const int32_t* out_rows;       /* optional: nbr is tile-sorted (see gcd_conv_args.out_rows); tcgen05 units only */
  const int32_t* back_out_rows;  /* optional: the same for back_nbr */
  const uint32_t* tile_masks;      /* optional: per-tile offset masks of nbr (see gcd_conv_args.tile_masks) */
  const uint32_t* back_tile_masks; /* optional: the same for back_nbr */
} gcd_convbn;

typedef struct {
  gcd_convbn u1, u2, ud;     /* first unit; second unit (has_u2); shortcut 1x1 conv + bn (has_ud) */
  int32_t has_u2, has_ud;
  int32_t relu1;             /* single unit (has_u2 == 0): apply ReLU after the batch norm */
  int32_t dtype;             /* gcd_dtype of every activation / gradient buffer */
  const void* x; int64_t ld_x;
  /* forward buffers: y = conv result, a1 = relu(bn(y1)), rd = bn(yd), out = relu(bn(y2) + shortcut) */
  void* y1; void* a1; void* y2; void* yd; void* rd; void* out;
  /* backward buffers */
  const void* gout;          /* gradient of out (of a1 for a single unit) */
  void* dy2; void* dres; void* da1; void* dy1; void* dyd; void* dx; void* dxd;
  int32_t need_dx;
  int32_t launches;          /* out: kernels launched by the call */
  int64_t ld_gout;           /* leading dimension of gout; 0 = dense (the width of the block's output) */
} gcd_block_args;

int32_t gcd_block_forward(gcd_block_args* args, void* stream);
int32_t gcd_block_backward(gcd_block_args* args, void* stream);

/* A whole network pass as ONE call: the host fills an array of operations (blocks and the column copies / adds that
 * stand for ME.cat and its backward, ref models/minkunet.py:178-208) and the library issues them in order on `stream`.
 * With one call per block the MinkUNet34 step costs the host ~70 calls and as many autograd nodes; with this it is two.
 *   GCD_OP_BLOCK_FORWARD / _BACKWARD : gcd_block_forward / gcd_block_backward on `block`
 *   GCD_OP_COPY_COLS : dst[i, 0:c] = src[i, 0:c]  for i < n   (row pitches ld_dst / ld_src, elements of `dtype`)
 *   GCD_OP_ADD_COLS  : dst[i, 0:c] += src[i, 0:c]
 *   GCD_OP_RECORD_EVENT : dst is a cudaEvent_t; it is recorded behind everything the call has issued so far on the caller's
 *                      stream and on the context's side stream (a data-parallel caller starts the all-reduce of the parameter
 *                      gradients a network stage has finished while the rest of the backward pass runs)
 * c, both pitches and both pointers must be multiples of 16 bytes.  *launches (host, optional) receives the kernel count. */
typedef enum { GCD_OP_BLOCK_FORWARD = 0, GCD_OP_BLOCK_BACKWARD = 1, GCD_OP_COPY_COLS = 2, GCD_OP_ADD_COLS = 3, GCD_OP_RECORD_EVENT = 4 } gcd_op_kind;
typedef struct {
  int32_t op;                /* gcd_op_kind */
  int32_t dtype;             /* gcd_dtype (copy / add) */
  gcd_block_args* block;     /* block operations */
  void* dst; int64_t ld_dst;
  const void* src; int64_t ld_src;
  int64_t n;
  int32_t c;
  int32_t reserved;
} gcd_op;
int32_t gcd_run_ops(gcd_op* ops, int32_t n_ops, void* stream, int32_t* launches);

/* The same with an execution context (caller-owned; use one per stream, from one host thread at a time):
 *   - the tcgen05 kernels claim their tiles dynamically (gcd_conv_args.sched; the context owns the counters);
 *   - in backward blocks the weight-gradient kernels, which feed nothing but the optimiser, are issued on the context's second
 *     stream behind an event and joined at the end of the call: they fill the SMs the gradient chain leaves idle (the deep
 *     levels of a U-Net have fewer row tiles than SMs) and the tails of its persistent kernels (GCD_OPT_WGRAD_SIDE).
 * Results are those of gcd_run_ops (weight gradients are accumulated with atomics in both). */
int32_t gcd_exec_create(void** exec);
int32_t gcd_exec_destroy(void* exec);
int32_t gcd_run_ops_exec(void* exec, gcd_op* ops, int32_t n_ops, void* stream, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* GCDLSS_B200_H_ */
