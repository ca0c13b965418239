/*
 * insmos_b200.h -- C ABI of libinsmos_b200.so
 *
 * B200-native (sm_100a) kernels for the sparse-voxel forward path of InsMOS
 * (SURVEY.md section 8).  This is the drop-in boundary: plain pointers and sizes,
 * no torch / ATen / pybind types.  Every pointer marked [dev] is a device
 * pointer owned by the caller (the Python host side allocates through the
 * PyTorch caching allocator); `stream` is a cudaStream_t passed as void*.
 * Every entry point returns 0 on success or a negative INSMOS_ERR_* code; none
 * of them exits the process or throws (the reference's native code calls
 * exit(-1) on error: models/bbox_post_process/src/iou3d_nms.cpp:14-38).
 * No entry point synchronises the stream; data-dependent sizes are written to
 * small [dev] counter arrays that the caller reads back when it needs them.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   - MinkowskiEngine (external, un-vendored): ME.utils.sparse_collate,
 *     ME.TensorField(...).sparse(), SparseTensor.slice, MinkowskiConvolution,
 *     MinkowskiConvolutionTranspose, MinkowskiBatchNorm, MinkowskiReLU, ME.cat
 *     -- call sites models/backbones_3d/motionnet.py:21-50,
 *        models/MinkowskiEngine/minkunet.py:52-181, resnet.py:87-126
 *   - spconv 2.3.6 (external): PointToVoxel.generate_voxel_with_id,
 *     SparseConvTensor(.dense), SubMConv3d, SparseConv3d, SparseInverseConv3d,
 *     gather_features_by_pc_voxel_id
 *     -- call sites models/backbones_3d/voxel_generate.py:17-31,
 *        models/backbones_3d/spconv_unet.py:120-208,284-410
 *   - iou3d_nms_cuda.nms_gpu  (models/bbox_post_process/src/iou3d_nms.cpp:90-136,
 *     iou3d_nms_kernel.cu:267-311)
 *   - Array_Index.find_features_by_bbox_with_yaw (models/utils/src/Array_Index.cpp:14-79)
 *   - CenterHead.generate_predicted_boxes (models/backbones_2d/center_head.py:251-276)
 *     + sigmoid/max of post_processing (models/post_process.py:146-192)
 */
#ifndef INSMOS_B200_H
#define INSMOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INSMOS_OK                 0
#define INSMOS_ERR_INVALID_ARG   -1
#define INSMOS_ERR_CUDA          -2
#define INSMOS_ERR_UNSUPPORTED   -3

/* bits of the device-side error word (counters[INSMOS_CNT_ERR]) */
#define INSMOS_DEVERR_COORD_RANGE  1   /* coordinate does not fit the 64-bit key packing */
#define INSMOS_DEVERR_ROW_RANGE    2   /* row index does not fit 25 bits of a rule-book entry */

/* layout of the int32 counter arrays written by the coordinate ops */
#define INSMOS_CNT_ROWS   0   /* number of unique rows / voxels produced            */
#define INSMOS_CNT_AUX    1   /* op specific (voxelize4d: number of t==0 points)   */
#define INSMOS_CNT_ERR    2   /* INSMOS_DEVERR_* bits                              */
#define INSMOS_CNT_TOTAL  3   /* op specific (voxelize3d: distinct voxels before the cap) */
#define INSMOS_NUM_COUNTERS 4

/* One slot of the coordinate hash table: 16 bytes.
 *   key   packed (batch,c0,c1,c2,c3); all-ones = empty
 *   first smallest input index that produced the key (defines first-occurrence order)
 *   row   output row id (-1 while unassigned / dropped by a cap)                     */
typedef struct {
    uint64_t key;
    int32_t  first;
    int32_t  row;
} insmos_slot_t;

/* Rule-book geometry: how the input coordinate probed for (output row, kernel offset k)
 * is derived.  Coordinates are rows of int32 [ncol] = (batch, c0..c{ndim-1}).
 * mode 0 (affine): in[d] = (out[d]*a[d] + b[d] + k_d*e[d]) / q[d]   (must divide exactly)
 * mode 1 (ME transposed, kernel 2 stride 2): in[d] = floor(out[d]/up_q[d])*up_q[d]; the pair
 *        exists only for the k whose digits equal (out[d]-in[d])/up_ts[d].
 * k digits: k = sum_d k_d * prod_{d' before d} ksize[d'], dimension 0 fastest when
 * first_fastest != 0 (MinkowskiEngine), last dimension fastest otherwise (spconv zyx). */
typedef struct {
    int32_t mode;
    int32_t ncol;
    int32_t ndim;
    int32_t first_fastest;
    int32_t K;
    int32_t ksize[4];
    int32_t a[4];
    int32_t b[4];
    int32_t e[4];
    int32_t q[4];
    int32_t up_q[4];
    int32_t up_ts[4];
} insmos_mapspec_t;

/* Epilogue fused into convolution / linear kernels:
 *   v = acc; v = v*scale[c] + shift[c] (if scale) ; v += bias[c] (if bias) ;
 *   v += residual[row,c] (if residual) ; v = max(v,0) (if relu)                      */
typedef struct {
    const float* scale;     /* [dev] [Cout] or NULL */
    const float* shift;     /* [dev] [Cout] or NULL (required when scale given) */
    const float* bias;      /* [dev] [Cout] or NULL */
    const float* residual;  /* [dev] [N_out, Cout] or NULL */
    int32_t relu;
    /* [dev] one int32 or NULL.  Dead-row elimination: output rows below *first_row are not needed by
     * the caller and MAY be left unwritten (kernels that do not implement the hint compute every row).  The value lives on the
     * device (insmos_time_row_starts), so no host read sits between the coordinate ops and the convolutions. */
    const int32_t* first_row;
} insmos_epilogue_t;

const char* insmos_version(void);
const char* insmos_last_error(void);          /* text of the last CUDA error seen by this thread */
/* number of CUDA kernels this library has launched in the process so far (incremented at every launch site);
 * bench.py reports the difference over its timed region as `gpu_launches`.  No reference counterpart. */
uint64_t insmos_launch_count(void);
int64_t insmos_hash_capacity(int64_t n);      /* power of two >= 2n (min 1024) */
int64_t insmos_scan_scratch_bytes(int64_t n); /* scratch needed by ops that scan n elements */

/* ---- coordinate ops ---------------------------------------------------------------------- */

int insmos_table_clear(insmos_slot_t* table, int64_t cap, void* stream);

/* a1+a2 (motionnet.py:22-36): fused quantise (fp32 true division, floor) + unique.
 * points [n, point_stride] f32 rows (x,y,z,intensity,t); quant = (vx,vy,vz,dt) host array.
 * Outputs: out_coords [<=n,5] i32 (0,x,y,z,t) in first-occurrence order, inverse [n] i32
 * (point -> voxel row, -1 on range error), cur_index [<=n] i32 = ascending indices of the points
 * whose t/dt == 0 (motionnet.py:42), counters[ROWS]=#voxels, counters[AUX]=#current points. */
int insmos_voxelize4d(const float* points, int64_t n, int32_t point_stride, const float* quant,
                      insmos_slot_t* table, int64_t cap, int32_t* slot_of_point,
                      int32_t* out_coords, int32_t* inverse, int32_t* cur_index,
                      int32_t* counters, void* scratch, void* stream);

/* unique rows of int32 coords [n,ncol] in first-occurrence order (ME TensorField.sparse on
 * already-floored coordinates; ME coordinate-manager stride when q given: dims are floored to
 * multiples of q[d] first).  q = host array [ncol-1] or NULL. */
int insmos_unique_coords(const int32_t* coords, int64_t n, int32_t ncol, const int32_t* q,
                         insmos_slot_t* table, int64_t cap, int32_t* slot_of_point,
                         int32_t* out_coords, int32_t* inverse,
                         int32_t* counters, void* scratch, void* stream);

/* spconv SparseConv3d output coordinate generation (oracle order: input rows ascending,
 * kernel offsets ascending, first occurrence).  in_coords [n,4] (b,z,y,x). */
int64_t insmos_spconv_out_scratch_bytes(int64_t n, int32_t K);   /* size of `scratch` below */
int insmos_spconv_out_coords(const int32_t* in_coords, int64_t n,
                             const int32_t* ksize, const int32_t* stride, const int32_t* pad,
                             const int32_t* out_shape,
                             insmos_slot_t* table, int64_t cap,
                             int32_t* out_coords, int32_t* counters, void* scratch, void* stream);

/* a6+a7 (voxel_generate.py:17-31, mean_vfe.py:47-52): capped hard voxelisation with ids +
 * fused per-voxel mean.  points [n,C] f32 (first three columns x,y,z); range = (xmin,ymin,zmin,
 * xmax,ymax,zmax), vsize = (vx,vy,vz), grid = (gx,gy,gz) host arrays.
 * Outputs: coords [<=max_voxels,4] i32 (0,z,y,x), num_points [<=max_voxels] i32,
 * voxels [<=max_voxels,max_points,C] f32 zero padded (may be NULL), mean [<=max_voxels,C] f32,
 * pc_voxel_id [n] i32 (-1 dropped). work = [max_voxels*(1+max_points)] i32 scratch. */
int insmos_voxelize3d(const float* points, int64_t n, int32_t C,
                      const float* range, const float* vsize, const int32_t* grid,
                      int32_t max_voxels, int32_t max_points,
                      insmos_slot_t* table, int64_t cap, int32_t* slot_of_point,
                      int32_t* coords, int32_t* num_points, float* voxels, float* mean,
                      int32_t* pc_voxel_id, int32_t* work,
                      int32_t* counters, void* scratch, void* stream);

/* ---- rule books --------------------------------------------------------------------------- */

/* Output-stationary, tiled, k-bucketed rule book (a4 + spconv indice pairs):
 * output rows are cut into tiles of TM rows; tile t owns entries[t*TM*K .. ) ; seg[t*(K+1)+k]
 * (uint16) = start of offset k's bucket inside the tile; entry = (out_row_in_tile << 25) | in_row.
 * pair_count[0] (uint64, [dev]) accumulates the number of pairs. */
int64_t insmos_rulebook_entries_capacity(int64_t n_out, int32_t K, int32_t TM);
int insmos_rulebook_build(const int32_t* out_coords, int64_t n_out,
                          const insmos_slot_t* in_table, int64_t in_cap,
                          const insmos_mapspec_t* spec, int32_t TM,
                          uint16_t* seg, uint32_t* entries, unsigned long long* pair_count,
                          int32_t* counters, void* stream);

/* X-block table of a coordinate set: one 32-byte slot per (floor(x/xstep) >> 2, other coordinates) holding the rows of
 * the block's 4 voxels.  The cube kernel maps (ME 5x5x5x1, 3x3x3x3; minkunet.py:55-60,71-128) probe runs of consecutive
 * x, which cost 1-2 slot lookups here instead of 3-5 voxel-table probes.  table: insmos_xblock_capacity(n) slots of
 * 32 bytes, [dev], caller-owned; coordinates must be multiples of xstep in x (a set at tensor stride xstep) and inside
 * the packable range already checked by insmos_voxelize4d / insmos_unique_coords. */
int64_t insmos_xblock_capacity(int64_t n);
int insmos_xblock_build(const int32_t* coords, int64_t n, int32_t ncol, int32_t xstep, void* table, int64_t cap, void* stream);
/* insmos_rulebook_build through the x-block table (mode-0 spec with q = 1, a[0] = 1, e[0] = xstep, first dimension
 * fastest, 3 <= ksize[0] <= 8; INSMOS_ERR_UNSUPPORTED otherwise).  Output identical to insmos_rulebook_build. */
int insmos_rulebook_build_xb(const int32_t* out_coords, int64_t n_out,
                             const insmos_slot_t* in_table, int64_t in_cap,
                             const void* xtable, int64_t xcap, int32_t xstep,
                             const insmos_mapspec_t* spec, int32_t TM,
                             uint16_t* seg, uint32_t* entries, unsigned long long* pair_count, void* stream);

/* LEAF GRID of a coordinate set: the same voxels stored as a sparse grid of 4x4x4 (x1 in time) leaves -- a hash table of
 * leaf keys plus, per slot, the dense 64-entry array of the voxels' rows.  step[d] = lattice step of the set in dimension d
 * (tensor stride of a MinkowskiEngine level, 1 for spconv indices).  grid = insmos_leafgrid_bytes(cap) bytes of scratch,
 * cap = insmos_leafgrid_capacity(n).  insmos_rulebook_build_lg builds the same map as insmos_rulebook_build (bit-identical
 * seg / entries) for an affine spec (mode 0, q = 1, e[d] == step[d]): per output row at most 24 leaf probes, then every
 * kernel offset is one indexed load.  Inputs with too many leaves for the bounded probing (isolated voxels) raise an
 * overflow word inside the grid and the builder falls back to voxel-table probes on the device. */
int64_t insmos_leafgrid_capacity(int64_t n);
int64_t insmos_leafgrid_bytes(int64_t cap);
int insmos_leafgrid_build(const int32_t* coords, int64_t n, int32_t ncol, const int32_t* step,
                          void* grid, int64_t cap, void* stream);
int insmos_rulebook_build_lg(const int32_t* out_coords, int64_t n_out,
                             const insmos_slot_t* in_table, int64_t in_cap,
                             const void* grid, int64_t grid_cap, const int32_t* step,
                             const insmos_mapspec_t* spec, int32_t TM,
                             uint16_t* seg, uint32_t* entries, unsigned long long* pair_count, void* stream);

/* Transposed (MinkowskiConvolutionTranspose, kernel == stride) map built without hash probes: the only pair of fine row
 * i is (offset of i inside its coarse cell, parent[i]) where parent is the inverse map insmos_unique_coords(q) returned
 * when the coarse set was made (ME derives the same map by swapping the strided one: minkunet.py:96-125).  `spec` is the
 * mode-1 spec of insmos_rulebook_build; identical output layout and ordering. */
int insmos_rulebook_build_up(const int32_t* fine_coords, int64_t n_fine, const int32_t* parent,
                             const insmos_mapspec_t* spec, int32_t TM,
                             uint16_t* seg, uint32_t* entries, unsigned long long* pair_count, void* stream);

/* ---- feature ops -------------------------------------------------------------------------- */

/* sparse convolution forward over a tiled rule book (a3, a8, a12).
 * in [n_in,Cin] f32, weight [K,Cin,Cout] f32, out [n_out,Cout] f32.
 * This entry point is the SIMT fp32 FFMA path (any Cin/Cout); `algo` is reserved (pass 0 or 1). */
int insmos_sparse_conv_fwd(const float* in, int64_t n_in, int32_t Cin,
                           const float* weight, int32_t K, int32_t Cout,
                           const uint16_t* seg, const uint32_t* entries, int32_t TM,
                           float* out, int64_t n_out,
                           const insmos_epilogue_t* ep, int32_t algo, void* stream);

/* Default path of the same convolution: exact fp32 FFMA, one thread per rule-book pair with register-blocked
 * outputs (conv_ffma.cu).  Requires Cout % 4 == 0; returns INSMOS_ERR_UNSUPPORTED otherwise. */
int insmos_sparse_conv_fwd_ffma(const float* in, int64_t n_in, int32_t Cin,
                                const float* weight, int32_t K, int32_t Cout,
                                const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                float* out, int64_t n_out,
                                const insmos_epilogue_t* ep, void* stream);

/* Exact-fp32 block-cooperative path of the same convolution for the NARROW layers (conv_fma.cu): a block owns a
 * super-tile of rule-book tiles with its accumulators in shared memory, the kernel offsets are phases with W[k] staged
 * once per block (cp.async, double buffered), lane = (pair, output-channel group), FFMA from registers x broadcast
 * shared-memory weights.  Replaces MinkowskiConvolution / SubMConv3d forward for Cin % 8 == 0, Cout in {8,16,32}
 * (minkunet.py:139-181, spconv_unet.py:284-406); insmos_sparse_conv_fma_supported tells; other shapes return
 * INSMOS_ERR_UNSUPPORTED.  weight is the plain [K,Cin,Cout] fp32 tensor. */
int insmos_sparse_conv_fma_supported(int32_t K, int32_t Cin, int32_t Cout);
int insmos_sparse_conv_fwd_fma(const float* in, int64_t n_in, int32_t Cin,
                               const float* weight, int32_t K, int32_t Cout,
                               const uint16_t* seg, const uint32_t* entries, int32_t TM,
                               float* out, int64_t n_out,
                               const insmos_epilogue_t* ep, void* stream);

/* Tensor-core path of the same convolution (3xTF32 on mma.sync m16n8k8, fp32 accumulate, fp32-accurate).
 * The weights are first rearranged ONCE per layer into tensor-core fragment order, pre-split into TF32 hi/lo
 * (wfrag: insmos_conv_wfrag_elems(K,Cin,Cout) 32-bit words, 16-byte aligned); the convolution then takes wfrag. */
int64_t insmos_conv_wfrag_elems(int32_t K, int32_t Cin, int32_t Cout);
int insmos_conv_prep_weights(const float* weight, int32_t K, int32_t Cin, int32_t Cout, void* wfrag, void* stream);
int insmos_sparse_conv_fwd_tc(const float* in, int64_t n_in, int32_t Cin,
                              const void* wfrag, int32_t K, int32_t Cout,
                              const uint16_t* seg, const uint32_t* entries, int32_t TM,
                              float* out, int64_t n_out,
                              const insmos_epilogue_t* ep, void* stream);

/* 5th-generation tensor-core path of the same convolution (tcgen05.mma kind::tf32 with TMEM accumulators, TMA bulk
 * copies of the weight tiles; spconv_umma.cu), for the wide layers: spconv SubMConv3d / SparseConv3d /
 * SparseInverseConv3d with Cout in {16,32,...,256} (reference call sites spconv_unet.py:120-208).  Output-stationary
 * implicit GEMM over 128-row super-tiles, 3xTF32 (fp32-accurate).  The weights are first rearranged ONCE per layer into
 * pre-swizzled TF32 hi/lo operand images (wimg: insmos_conv_wimg_elems(K,Cin,Cout) floats; 0 = shape unsupported). */
int64_t insmos_conv_wimg_elems(int32_t K, int32_t Cin, int32_t Cout);
int insmos_conv_prep_weights_umma(const float* weight, int32_t K, int32_t Cin, int32_t Cout, float* wimg, void* stream);
int insmos_sparse_conv_fwd_umma(const float* in, int64_t n_in, int32_t Cin,
                                const float* wimg, int32_t K, int32_t Cout,
                                const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                float* out, int64_t n_out,
                                const insmos_epilogue_t* ep,
                                void* workspace, int64_t workspace_bytes, void* stream);
/* Optional scratch for the call above ([dev], caller-owned, contents irrelevant, may be reused by later calls on the
 * same stream): layers with few 128-row super-tiles split the kernel offsets over several CTAs and reduce the partial
 * tiles through it in a fixed order.  0 = not needed; passing NULL only disables the split. */
int64_t insmos_sparse_conv_umma_workspace_bytes(int64_t n_out, int32_t Cout);

/* Dense 3x3 / pad 1 convolution of a channels-last image [H*W, Cin] (BEV backbone, base_bev_backbone.py:33-82, BatchNorm
 * folded into weight/bias) through the same tcgen05 kernel as insmos_sparse_conv_fwd_umma: rows = pixels, arithmetic
 * neighbour table.  wimg = insmos_conv_prep_weights_umma of the [9, Cin, Cout] weight; Cout % 16 == 0, Cout <= 128. */
int insmos_conv2d_nhwc_umma(const float* in, int32_t H, int32_t W, int32_t Cin,
                            const float* wimg, int32_t Cout, const float* bias, int32_t relu, float* out,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* out[n,Cout] = in[n,Cin] . weight[Cin,Cout] + epilogue (ME kernel_size==1 conv, nn.Linear) */
int insmos_linear_fwd(const float* in, int64_t n, int32_t Cin, const float* weight, int32_t Cout,
                      float* out, const insmos_epilogue_t* ep, void* stream);

/* elementwise epilogue alone on x [n,C] (MinkowskiBatchNorm eval, ReLU, residual add) */
int insmos_affine_act(const float* x, int64_t n, int32_t C, float* out,
                      const insmos_epilogue_t* ep, void* stream);

/* out[n,C1+C2] = cat(a[n,C1], b[n,C2]) (ME.cat, torch.cat on features) */
int insmos_concat2(const float* a, int32_t C1, const float* b, int32_t C2, int64_t n,
                   float* out, void* stream);

/* out[n,C] = a[n,C] + view(b[n,2C], n,C,2).sum(2)   (spconv_unet.py:213-238) ; a may be NULL */
int insmos_pairsum_add(const float* a, const float* b, int64_t n, int32_t C, float* out, void* stream);

/* out[i,:] = idx[i] >= 0 ? src[idx[i],:] : 0   (SparseTensor.slice, gather_features_by_pc_voxel_id) */
int insmos_gather_rows(const float* src, int32_t C, const int32_t* idx, int64_t n, float* out, void* stream);

/* per-voxel unweighted average of point features (ME TensorField.sparse default quantisation) */
int insmos_segment_mean(const float* feat, int32_t C, const int32_t* inverse, int64_t n,
                        float* out, int32_t* cnt, int64_t n_rows, void* stream);

/* a5 (motionnet.py:42-48): out[j,0:4] = points[cur_index[j],0:4]; out[j,4:4+Cm] = vox_feat[inverse[cur_index[j]],0:Cm] */
int insmos_build_current_points(const float* points, int32_t point_stride, const int32_t* cur_index,
                                int64_t n_cur, const int32_t* inverse, const float* vox_feat, int32_t Cfeat,
                                int32_t Cm, float* out, void* stream);

/* SparseConvTensor.dense() for batch 1: out[C,D,H,W] (pre-zeroed by the callee) ; coords [n,4] (b,z,y,x) */
int insmos_dense_scatter(const float* feat, const int32_t* coords, int64_t n, int32_t C,
                         int32_t D, int32_t H, int32_t W, float* out, void* stream);

/* Dense BEV convolutions (a9; base_bev_backbone.py:84-115; the reference runs them through cuDNN) on the 5th-generation
 * tensor cores: tcgen05.mma kind::tf32 (3xTF32 split, fp32-accurate), accumulators in TMEM, weight tiles by TMA bulk copy,
 * mbarrier pipeline (bev_tcgen05.cu).  NHWC activations [H*W, C]; weight [taps, Cin, Cout] f32 with BatchNorm(eval)
 * pre-folded, rearranged once per layer into pre-swizzled TF32 hi/lo tile images (wimg: insmos_bev_wimg_elems(...) floats);
 * bias [Cout] or NULL.  mode 0: 3x3 stride 1 zero-pad 1 (taps=9, tap = ky*3+kx); mode 1: 1x1 (taps=1); mode 2: 2x2
 * stride-2 transposed conv (taps=4, tap = dy*2+dx, out is [2H*2W, Cout]).  Requires Cin % 32 == 0 and Cout % 128 == 0. */
int64_t insmos_bev_wimg_elems(int32_t taps, int32_t Cin, int32_t Cout);
int insmos_bev_prep_weights_tcgen05(const float* weight, int32_t taps, int32_t Cin, int32_t Cout, float* wimg, void* stream);
int insmos_conv2d_nhwc_tcgen05(const float* in, int32_t H, int32_t W, int32_t Cin,
                               const float* wimg, int32_t mode, int32_t Cout,
                               const float* bias, int32_t relu, float* out, void* stream);

/* SparseConvTensor.dense() + HeightCompression view, channels-last: out[(y*W+x), c*D+z] (height_compression.py:26-30) */
int insmos_dense_scatter_nhwc(const float* feat, const int32_t* coords, int64_t n, int32_t C,
                              int32_t D, int32_t H, int32_t W, float* out, void* stream);

/* ---- detection head ----------------------------------------------------------------------- */

/* CenterHead decode + sigmoid/max (center_head.py:251-276, post_process.py:146-192).
 * cls / box are the head outputs for batch 1 in any layout: element (channel c, pixel i) of cls is
 * cls[c*cls_cs + i*cls_ps] (NCHW: cs=H*W, ps=1; NHWC [H*W,stride]: cs=1, ps=stride), same for box (8 channels).
 * boxes [H*W,7], scores [H*W], labels [H*W] (1-based). */
int insmos_center_decode(const float* cls, int64_t cls_cs, int64_t cls_ps, const float* box, int64_t box_cs, int64_t box_ps,
                         int32_t ncls, int32_t H, int32_t W,
                         float out_size_factor, float vx, float vy, float x_min, float y_min,
                         float* boxes, float* scores, int32_t* labels, void* stream);

/* rotated BEV NMS on boxes [n,7] already sorted by descending score (nms_gpu).
 * mask = [n*ceil(n/64)] uint64 scratch. keep [<=max_keep] i32 ascending, num_keep [1] i32. */
int insmos_nms_rotated(const float* boxes, int32_t n, float thresh, int32_t max_keep,
                       unsigned long long* mask, int32_t* keep, int32_t* num_keep, void* stream);
/* Same result through a PAIR LIST (the default of the host side): the centre-distance test of the dense kernel only appends
 * the surviving (i, j) pairs to `pairs` (uint2 [pair_cap], insmos_nms_pair_capacity(n) entries suffice for BEV detections),
 * a second kernel evaluates one pair per thread and ORs the mask bit; if the list overflows the dense kernel runs instead
 * (decided on the device).  count = uint32 [1] scratch.  Replaces iou3d_nms_cuda.nms_gpu (iou3d_nms.cpp:90-136). */
int64_t insmos_nms_pair_capacity(int32_t n);
int insmos_nms_rotated_pairs(const float* boxes, int32_t n, float thresh, int32_t max_keep,
                             unsigned long long* mask, void* pairs, int64_t pair_cap, uint32_t* count,
                             int32_t* keep, int32_t* num_keep, void* stream);

/* boxes7 [nb,7] metric (x,y,z,dx,dy,dz,yaw) + labels [nb] -> boxes8 [nb,8] in voxel units of the
 * stride-`stride` level (spconv_unet.py:321-330), fp32 op order preserved. */
int insmos_boxes_to_voxel_units(const float* boxes7, const int32_t* labels, int32_t nb,
                                const float* range_min, const float* vsize, float stride,
                                float* boxes8, void* stream);

/* Array_Index.find_features_by_bbox_with_yaw on device, including its first-hit pruning quirk.
 * coords [n,4] (b,z,y,x) i32; boxes8 [nb,8] (cx,cy,cz,dx,dy,dz,yaw,label) scaled by `mult`
 * (exact power of two, spconv_unet.py:358,373,388). out[j*out_stride + label-1] = 1.0f for hits
 * (caller zero-initialises). first_hit [nb] i32 scratch. */
int insmos_box_membership(const int32_t* coords, int64_t n, const float* boxes8, int32_t nb, float mult,
                          float* out, int32_t out_stride, int32_t* first_hit, void* stream);

/* iou3d_nms_utils.boxes_iou3d_gpu (models/bbox_post_process/iou3d_nms_utils.py:27-61, iou3d_nms_kernel.cu:236-249,377-387):
 * iou [na, nb] of boxes [.,7] (x,y,z,dx,dy,dz,heading): rotated BEV overlap x height overlap over the union volume.  Used by
 * the recall record of the 'eval' mode (models/post_process.py:66-109). */
int insmos_boxes_iou3d(const float* boxes_a, int32_t na, const float* boxes_b, int32_t nb, float* iou, void* stream);

/* ---- steps either side of the forward path ---------------------------- */

/* N1 input staging (scripts/predict_mos.py:114-159 DemoDataset.__getitem__, :161-166 transform_point_cloud, :174-179
 * timestamp_tensor): raw scans concatenated oldest -> newest as [total,4] f32 (x,y,z,intensity), scan s occupying rows
 * [scan_offsets[s], scan_offsets[s+1]) (int64 [n_scans+1], [dev]); transforms = n_scans row-major 4x4 float64
 * inv(to_pose) @ from_pose ([dev], may be NULL when apply_transform == 0); timestamps f32 [n_scans] ([dev]).
 * out [total,5] f32 (x,y,z,intensity,t): xyz = float32(T @ (x,y,z,1)) evaluated in float64 like the reference. */
int insmos_stage_scans(const float* raw_xyzi, const int64_t* scan_offsets, int32_t n_scans,
                       const double* transforms, const float* timestamps, int32_t apply_transform,
                       float* out_xyzit, int64_t total_points, void* stream);

/* N2 output labelling (scripts/predict_mos.py:440-454, :279-283): logits [n,n_class] f32 -> classes in ignore_mask
 * (bit c = class c) to -inf, softmax, confidence [n,n_class-1] (columns 1..), argmax (first maximum), labels[n] =
 * label_map[argmax] (int32 [n_class], [dev]; NULL = identity).  confidence may be NULL. */
int insmos_mos_labels(const float* logits, int64_t n, int32_t n_class, uint32_t ignore_mask,
                      const int32_t* label_map, int32_t* labels, float* confidence, void* stream);

/* ---- instance refinement after the forward path (scripts/refine.py:169-302) --------------------------
 * insmos_point_instance_ids replaces Array_Index.find_point_in_instance_bbox_with_yaw (models/utils/src/Array_Index.cpp:83-149):
 * points [n,stride] f32 (x,y,z first), boxes8 [nb,8] (cx,cy,cz,dx,dy,dz,yaw,label), out_ground added to the box centre z;
 * ids [n,ncls] i32 receives box index + 1 in column label-1 (0 = no box; zeroed by the call), with the reference's first-hit
 * pruning window; where two boxes of one class contain a point the later box wins (the reference's serial order).
 * first_hit = i32 [nb] scratch. */
int insmos_point_instance_ids(const float* points, int64_t n, int32_t stride, const float* boxes8, int32_t nb,
                              float out_ground, int32_t* ids, int32_t ncls, int32_t* first_hit, void* stream);
/* per-instance reductions of refine.py:208-221: stats [nb,3] i32 = points of instance b+1 (column col of ids), of them
 * labelled moving_label, of them with conf[j*conf_stride] >= conf_thresh (conf may be NULL). */
int insmos_instance_stats(const int32_t* ids, int32_t ncls, int32_t col, int64_t n, const int32_t* labels,
                          int32_t moving_label, const float* conf, int32_t conf_stride, float conf_thresh,
                          int32_t nb, int32_t* stats, void* stream);
/* label overwrites of refine.py:240-257,287-292: labels[j] = new_label[id] where id = ids[j,col] > 0 and new_label[id] >= 0;
 * new_label is i32 [nb+1] (entry 0 unused). */
int insmos_relabel_instances(const int32_t* ids, int32_t ncls, int32_t col, int64_t n, const int32_t* new_label,
                             int32_t nb, int32_t* labels, void* stream);

/* ---- dead-row elimination for the 4D MotionNet decoder ----------------------------------------------
 * Only the rows of the newest scan (time index 0) leave MotionNet (motionnet.py:42-45); a 3x3x3x3 convolution reaches one
 * time index back, so the last decoder layers need only the rows with time index >= -j.
 * starts[j] (j = 0..15) = smallest row index whose time coordinate (column tcol of coords [n,ncol]) is >= -j, or n when there
 * is none: every row needed at threshold -j has an index >= starts[j] whatever the input order (time-ordered input, as
 * predict_mos.py:146-158 produces it, makes the needed rows exactly a suffix).  starts must hold 33 int32 (16 results + scratch). */
int insmos_time_row_starts(const int32_t* coords, int64_t n, int32_t ncol, int32_t tcol, int32_t* starts, void* stream);

/* insmos_rulebook_build_lg restricted to the output tiles that contain rows >= *first_row ([dev], may be NULL = all);
 * seg / entries of the skipped tiles are left unwritten and pair_count counts the built tiles only. */
int insmos_rulebook_build_lg_from(const int32_t* out_coords, int64_t n_out,
                                  const insmos_slot_t* in_table, int64_t in_cap,
                                  const void* grid, int64_t grid_cap, const int32_t* step,
                                  const insmos_mapspec_t* spec, int32_t TM,
                                  uint16_t* seg, uint32_t* entries, unsigned long long* pair_count,
                                  const int32_t* first_row, void* stream);

/* ---- training step (BASELINE config 5) -----------------------------------------------
 * The reference trains through autograd over MinkowskiEngine / spconv (models/models.py:61-98,330-347,
 * scripts/train.py:74-79).  The DATA gradient of a sparse convolution is insmos_sparse_conv_fwd_* over the transposed rule
 * book with transposed weights; the entry points below are the pieces that have no forward counterpart. */

/* number of tile slices (partial matrices) insmos_sparse_conv_wgrad needs for this map: partial holds S*K*Cin*Cout floats */
int32_t insmos_sparse_conv_wgrad_slices(int64_t n_out, int32_t TM, int32_t K, int32_t Cin);

/* WEIGHT gradient of MinkowskiConvolution / SubMConv3d / SparseConv3d / SparseInverseConv3d (replaces the autograd backward of
 * ME's ConvolutionFunction / spconv's SparseConvFunction at the call sites minkunet.py:139-181, spconv_unet.py:284-406):
 * dweight[k][ci][co] = sum over the pairs (i,o) of bucket k of in[i][ci] * dout[o][co].  Deterministic (fixed summation
 * order: per-slice partial matrices added in slice order).  [dev] in [n_in,Cin], dout [n_out,Cout], seg/entries = the
 * forward rule book, partial [S,K,Cin,Cout] scratch, dweight [K,Cin,Cout]. */
int insmos_sparse_conv_wgrad(const float* in, int64_t n_in, int32_t Cin, const float* dout, int64_t n_out, int32_t Cout,
                             const uint16_t* seg, const uint32_t* entries, int32_t TM, int32_t K,
                             float* partial, int32_t S, float* dweight, void* stream);

/* column reductions over [n,C] fp32 matrices into fp64 sums (train-mode MinkowskiBatchNorm / nn.BatchNorm1d,
 * minkunet.py:52-131, spconv_unet.py:118):  mode 0: out0 = sum a;  mode 1: out0 = sum (a - mean)^2;
 * mode 2: g = (gate == NULL || gate > 0) ? a : 0, out0 = sum g, out1 = sum g * (b - mean) * invstd;
 * mode 3: out0 = sum a, out1 = sum a^2 (one pass, exact products in fp64).   C <= 1024. */
int insmos_column_moments(const float* a, const float* b, const float* gate, const float* mean, const float* invstd,
                          int64_t n, int32_t C, int32_t mode, double* out0, double* out1, void* stream);

/* per-channel constants of train-mode BatchNorm from the fp64 sums of insmos_column_moments(mode 3): mean, invstd, the fused
 * affine (scale = gamma*invstd, shift = beta - mean*scale) for insmos_affine_act, and nn.BatchNorm1d's running-statistics
 * update (running = (1-momentum)*running + momentum*batch, variance unbiased); running_* may be NULL. */
int insmos_bn_train_finalize(const double* sum, const double* sumsq, int64_t n, int32_t C, const float* gamma, const float* beta,
                             float eps, float momentum, float* mean, float* invstd, float* scale, float* shift,
                             float* running_mean, float* running_var, void* stream);

/* BatchNorm backward, elementwise part: dx = gamma*invstd*(g - s0/n - xhat*s1/n), g = dy gated by gate > 0 (fused ReLU), s0/s1 =
 * the fp64 sums of insmos_column_moments(mode 2); also writes dbeta = s0, dgamma = s1 (either may be NULL). */
int insmos_bn_bwd_apply(const float* dy, const float* x, const float* gate, const float* mean, const float* invstd,
                        const float* gamma, const double* s0, const double* s1, int64_t n, int32_t C, float* dx,
                        float* dgamma, float* dbeta, void* stream);

/* out[idx[i],0:C] += src[i,0:C] (idx < 0 skipped; src row stride ldsrc): backward of insmos_gather_rows /
 * insmos_build_current_points / gather_features_by_pc_voxel_id.  out must be initialised by the caller. */
int insmos_scatter_add_rows(const float* src, int32_t C, int32_t ldsrc, const int32_t* idx, int64_t n, float* out, void* stream);

/* CenterHead.get_targets_single (models/backbones_2d/center_head.py:171-249) for one sample, one block per box instead of
 * the reference's Python loop: gt_boxes [n_box,8] (x,y,z,dx,dy,dz,yaw,class 1..ncls) -> heatmap [ncls,H,W] (Gaussian
 * peaks, running maximum), anno_boxes [max_objs,8], inds int64 [max_objs], masks uint8 [max_objs] (all zeroed first).
 * range_is_fp64: 0 when the configured point-cloud range is integral (torch then keeps the centre arithmetic in fp32), 1 when
 * it holds floats (fp32 box - fp64 range promotes to fp64). */
int insmos_center_targets(const float* gt_boxes, int32_t n_box, int32_t max_objs, int32_t H, int32_t W, int32_t ncls,
                          double x_min, double y_min, int32_t range_is_fp64, float vx, float vy, int32_t out_size_factor,
                          float min_overlap, int32_t min_radius, float* heatmap, float* anno_boxes, int64_t* inds,
                          uint8_t* masks, void* stream);

/* torch.optim.Adam step (models/models.py:185-190) over one flat fp32 parameter buffer; grad_scale multiplies the gradient
 * first (1/world_size after the data-parallel all-reduce); step counts from 1. */
int insmos_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INSMOS_B200_H */
