/*
 * droid_b200.h -- C ABI of the B200-native (sm_100a) dense-BA update hot path of DROID-SLAM.
 *
 * This is the drop-in boundary: every entry point takes plain device pointers, extents, a dtype code and a
 * CUDA stream, and returns an int status (0 = ok).  No torch types cross it.  The Python extension
 * `droid_backends` (droid_slam_b200/csrc/binding/droid_backends.cpp) is a thin pybind11 layer that forwards the
 * reference's nine callables (reference src/droid.cpp:246-259) to these functions.
 *
 * All pointers are DEVICE pointers on the current CUDA device unless stated otherwise.  All tensors are dense
 * row-major ("contiguous") exactly as the reference requires (src/droid.cpp:89-90).
 * Index tensors are int64 (reference `LongType`, src/droid_kernels.cu:19-24).
 */
#ifndef DROID_B200_H
#define DROID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dba_stream_t; /* a cudaStream_t */

/* dtype codes for the correlation ops (reference dispatch: AT_DISPATCH_FLOATING_TYPES_AND_HALF,
 * src/correlation_kernels.cu:146, src/altcorr_kernel.cu:150; bf16 is an extension) */
enum { DBA_F32 = 0, DBA_F16 = 1, DBA_F64 = 2, DBA_BF16 = 3 };

/* status codes */
enum {
  DBA_OK = 0,
  DBA_ERR_INVALID = 1,   /* bad argument (null pointer, negative extent, unsupported dtype/radius) */
  DBA_ERR_CUDA = 2,      /* a CUDA runtime call or launch failed; see dba_last_error() */
  DBA_ERR_WORKSPACE = 3  /* workspace too small */
};

const char* dba_last_error(void);   /* thread-local, human readable */
/* device-wide L2 fetch granularity hint (32/64/128 B; cudaLimitMaxL2FetchGranularity); -1 when it cannot be read */
int dba_set_l2_fetch_granularity(int bytes);
int dba_get_l2_fetch_granularity(void);
int dba_version(void);              /* 100 * major + minor */

/* ---- correlation volume lookup ------------------------------------------------------------------
 * replaces corr_index_cuda_forward / corr_index_cuda_backward (reference src/correlation_kernels.cu:127-185,
 * bound at src/droid.cpp:175-196).
 * volume [n,h1,w1,h2,w2], coords [n,2,h1,w1] f32, corr [n,2r+1,2r+1,h1,w1] (x-offset major), dtype of volume.
 * forward fully overwrites `corr`; backward fully overwrites `volume_grad`. */
int dba_corr_index_forward(const void* volume, const float* coords, void* corr,
                           int n, int h1, int w1, int h2, int w2, int radius, int dtype, dba_stream_t stream);
int dba_corr_index_backward(const float* coords, const void* corr_grad, void* volume_grad,
                            int n, int h1, int w1, int h2, int w2, int radius, int dtype, dba_stream_t stream);

/* ---- correlation volume + pooled pyramid (tensor cores) -----------------------------------------------
 * replaces CorrBlock.__init__ / CorrBlock.corr (reference droid_slam/modules/corr.py:24-38,63-71: torch.matmul of the
 * /4-scaled feature maps + 3x avg_pool2d).  fmap1 [n_frames1,C,ht,wd], fmap2 [n_frames2,C,ht,wd] (f16, C = 128),
 * ii,jj [E] int64 frame indices into fmap1 / fmap2;  out_l [E,ht,wd,ht/2^l,wd/2^l] f16 for l = 0..3, fully overwritten.
 * One tcgen05/TMEM/TMA kernel writes all four levels in a single pass over the accumulator.
 * Implemented for wd = 64, ht % 8 == 0 (DBA_ERR_INVALID otherwise). */
/* 1 when dba_corr_volume_pyramid has a kernel for this shape / dtype (f16, 128 channels, wd = 64, ht % 8 == 0), else 0 */
int dba_corr_volume_supported(int channels, int ht, int wd, int dtype);
int dba_corr_volume_pyramid(const void* fmap1, const void* fmap2, const int64_t* ii, const int64_t* jj,
                            void* out0, void* out1, void* out2, void* out3,
                            int n_edges, int n_frames1, int n_frames2, int channels, int ht, int wd, int dtype, dba_stream_t stream);

/* private-layout variant: levels 0 and 1 of every plane stored as 4x8-element tiles ([h2/4][w2/8][4][8] f16 = one 64-byte DRAM atom per
 * tile) for dba_corr_lookup_pyramid(tiled_mask = 3); levels 2, 3 and all tensor shapes are unchanged.  NOT readable by
 * dba_corr_index_forward / the reference's CorrBlock.__call__. */
int dba_corr_volume_pyramid_tiled(const void* fmap1, const void* fmap2, const int64_t* ii, const int64_t* jj,
                                  void* out0, void* out1, void* out2, void* out3,
                                  int n_edges, int n_frames1, int n_frames2, int channels, int ht, int wd, int dtype, dba_stream_t stream);
/* CorrBlock.__call__ (reference droid_slam/modules/corr.py:40-50) in one launch: out [n,196,h1,w1] = concatenation over the four levels
 * of corr_index_forward(volume_l, coords / 2^l, 3) -- bit-identical values; coords [n,2,h1,w1] f32 at level-0 scale are read once.
 * f16 volumes, h1 % 8 == 0, w1 % 64 == 0.  tiled_mask: 0 = reference layout, 3 = levels 0 and 1 in the tiled layout above. */
int dba_corr_lookup_pyramid(const void* v0, const void* v1, const void* v2, const void* v3, const float* coords, void* out,
                            int n, int h1, int w1, int tiled_mask, int dtype, dba_stream_t stream);

/* ---- on-the-fly correlation ---------------------------------------------------------------------
 * replaces altcorr_cuda_forward / altcorr_cuda_backward (reference src/altcorr_kernel.cu:132-225, bound at
 * src/droid.cpp:198-226).
 * fmap1 [B,N1,C,H,W], fmap2 [B,N2,C,H2,W2], coords [B,M,2,H,W] f32, ii,jj [M] int64 (frame indices into
 * fmap1 / fmap2).  forward writes `out` as a CONTIGUOUS [B,M,2r+1(y-off),2r+1(x-off),H,W] tensor; the binding
 * returns its permute(0,1,3,2,4,5) view like the reference (:171).
 * backward consumes corr_grad [B,M,2r+1(x-off),2r+1(y-off),H,W] f32 -- the gradient of the returned (permuted)
 * tensor, as the reference host function does (:175-207) -- and fully overwrites fmap1_grad / fmap2_grad. */
int dba_altcorr_forward(const void* fmap1, const void* fmap2, const float* coords,
                        const int64_t* ii, const int64_t* jj, void* out,
                        int B, int N1, int N2, int C, int H, int W, int H2, int W2, int M,
                        int radius, int dtype, dba_stream_t stream);
int dba_altcorr_backward(const void* fmap1, const void* fmap2, const float* coords, const float* corr_grad,
                         const int64_t* ii, const int64_t* jj, void* fmap1_grad, void* fmap2_grad,
                         int B, int N1, int N2, int C, int H, int W, int H2, int W2, int M,
                         int radius, int dtype, dba_stream_t stream);

/* ---- streaming geometry -------------------------------------------------------------------------
 * poses [n_poses,7] (tx,ty,tz,qx,qy,qz,qw) f32, disps [n_disps,ht,wd] f32, intrinsics [4] f32 (fx,fy,cx,cy).
 * replace projmap_cuda / frame_distance_cuda / depth_filter_cuda / iproj_cuda
 * (reference src/droid_kernels.cu:1447-1550, bound at src/droid.cpp:125-171,228-242). */
int dba_projmap(const float* poses, const float* disps, const float* intrinsics,
                const int64_t* ii, const int64_t* jj, float* coords /*[E,ht,wd,3]*/, float* valid /*[E,ht,wd,1]*/,
                int n_edges, int ht, int wd, dba_stream_t stream);
/* fused reprojection feeding the update operator: replaces pops.projective_transform(jacobian=False)
 * (reference droid_slam/geom/projective_ops.py:165-198 via DepthVideo.reproject, depth_video.py:171-179).
 * intrinsics_per_frame [n_frames,4]; coords [E,ht,wd,2], valid [E,ht,wd,1]; MIN_DEPTH 0.2, stereo baseline for ii == jj. */
int dba_reproject(const float* poses, const float* disps, const float* intrinsics_per_frame,
                  const int64_t* ii, const int64_t* jj, float* coords, float* valid, int n_edges, int ht, int wd, dba_stream_t stream);
int dba_frame_distance(const float* poses, const float* disps, const float* intrinsics,
                       const int64_t* ii, const int64_t* jj, float* dist /*[K]*/,
                       int n_pairs, int ht, int wd, float beta, dba_stream_t stream);
int dba_depth_filter(const float* poses, const float* disps, const float* intrinsics,
                     const int64_t* ix, const float* thresh, float* counter /*[num,ht,wd]*/,
                     int num, int n_disps, int ht, int wd, dba_stream_t stream);
int dba_iproj(const float* poses, const float* disps, const float* intrinsics, float* points /*[n,ht,wd,3]*/,
              int n, int ht, int wd, dba_stream_t stream);

/* convex upsampling of inverse depth maps: replaces cvx_upsample / upsample_disp (reference droid_slam/droid_net.py:21-42) as called by
 * DepthVideo.upsample (depth_video.py:155-159).  disps [n,ht,wd] f32, mask [n,576,ht,wd] (f16 or f32; 9 taps x 8 x 8 sub-pixels, the
 * update operator's `upmask`), out [n,8*ht,8*wd] f32 = softmax-over-taps weighted sum of the 3x3 neighbourhood (zero padded). */
int dba_cvx_upsample(const float* disps, const void* mask, float* out, int n, int ht, int wd, int mask_dtype, dba_stream_t stream);

/* ---- factor-graph edge selection (row F1) --------------------------------------------------------
 * replaces the body of FactorGraph.add_proximity_factors between `video.distance(...)` and `add_factors(...)` (reference
 * droid_slam/factor_graph.py:357-411: masking, suppression around the edges the graph already has, temporal-neighbour edges, then the
 * greedy selection by increasing distance with non-maximum suppression -- a Python / NumPy triple loop on a CPU copy of d).
 * d [(t-t0)*(t-t1)] f32: frame distances over the grid i in [t0,t), j in [t1,t), row-major (what dba_frame_distance returns for the
 * meshgrid of factor_graph.py:351-356), read only.  ii_known / jj_known [n_known] int64: cat(ii, ii_bad, ii_inac) / cat(jj, ...).
 * es [cap][2] int64 receives the (i, j) rows in the reference's emission order; n_out_status [2] int32 on the device: [0] rows written,
 * [1] status (bit 0: cap too small, bit 1: the reference's unchecked index d[(i-t0)*(t-t1) + (j-t1)] left the array, where it raises
 * IndexError).  cap >= 2*n + (1+2*(rad+1))*(t-t0) always suffices.  Equal finite distances are visited in index order (torch.argsort
 * gives no order for ties).  No host synchronisation. */
size_t dba_proximity_workspace_bytes(int t0, int t1, int t);
int dba_proximity_edges(const float* d, int t0, int t1, int t, const int64_t* ii_known, const int64_t* jj_known, int n_known, int rad, int nms,
                        float thresh, int max_factors, int stereo, int64_t* es, int cap, int* n_out_status, void* workspace, size_t workspace_bytes,
                        dba_stream_t stream);

/* ---- dense bundle adjustment --------------------------------------------------------------------
 * replaces ba_cuda (reference src/droid_kernels.cu:1323-1443, bound at src/droid.cpp:93-122).
 * In place on poses [n_frames,7] and disps [n_frames,ht,wd]; disps_sens like disps; targets, weights
 * [E,2,ht,wd]; eta [M,ht,wd] with M = |unique(ii U [t0,t1))| rows in ascending frame order (or 1 row,
 * broadcast); ii,jj [E] int64.  Outputs of the LAST Gauss-Newton iteration: dx_out [t1-t0,6], dz_out [M,ht*wd]
 * (dz_out untouched when motion_only).  Everything runs on `stream` with no host synchronisation.
 *
 * The call is split at the one point where a multi-GPU run exchanges data:
 *   dba_ba_prepare   graph bookkeeping (unique / CSR by source frame), once per call
 *   dba_ba_build     per-edge blocks + depth elimination -> reduced pose system  Hsys [6P,6P] f64 (LOWER triangle valid),
 *                    bsys [6P] f64 (edge-sharded runs all-reduce exactly this buffer, 8*(36P^2+6P) bytes)
 *   dba_ba_solve     damping + Cholesky (fp64) -> dx, back-substitution -> dz, retraction of poses and disps
 * dba_ba runs prepare + iterations x (build, solve).
 * workspace: dba_ba_workspace_bytes(); layout is private, the reduced system sits at dba_ba_system_offset(). */
size_t dba_ba_workspace_bytes(int n_frames, int n_edges, int ht, int wd, int t0, int t1);
size_t dba_ba_system_offset(int n_frames, int n_edges, int ht, int wd, int t0, int t1);  /* bytes from workspace start */
size_t dba_ba_system_bytes(int t0, int t1);                                              /* 8*(36P^2+6P) */

typedef struct {
  float* poses; float* disps; const float* intrinsics; const float* disps_sens;
  const float* targets; const float* weights; const float* eta; int eta_rows;
  const int64_t* ii; const int64_t* jj;
  int n_frames, n_edges, ht, wd, t0, t1;
  float lm, ep; int motion_only;
  float* dx_out; float* dz_out;
  void* workspace; size_t workspace_bytes;
  dba_stream_t stream;
  /* edge-sharded multi-GPU runs (one rank = the out-edges of a contiguous range of source frames):
   * only depth frames in [own_lo, own_hi) get their inverse depth updated by dba_ba_solve (the other ranks own the
   * rest); single-GPU callers pass own_lo = 0, own_hi = n_frames.  eta_by_frame != 0: eta has n_frames rows indexed by
   * FRAME id instead of M rows in depth-frame order (a rank's local depth-frame set differs from the global one). */
  int own_lo, own_hi, eta_by_frame;
  /* optional fused peer-to-peer reduction of the pose system over NVLink peer memory (p2p_world > 1), replacing the separate
   * all-reduce: the rank accumulates its partial system in p2p_system[p2p_rank] (slot p2p_epoch & 1 of two, each 36P^2+6P
   * doubles; 8 uint64 flags follow the two slots), dba_ba_p2p_signal() publishes it with release stores into every peer's
   * flags[p2p_rank], and the Cholesky kernel of dba_ba_solve waits for the W flags and sums the W peer copies in rank order
   * (bit-identical on every rank) straight out of peer memory.  p2p_system[] are peer-mapped device pointers (e.g. from
   * torch.distributed._symmetric_memory); p2p_epoch must increase by one per Gauss-Newton iteration on every rank.
   * p2p_epoch_dev (optional, rank-local device memory, zero-initialised once): when set, the published / awaited epoch VALUE is
   * this device counter (dba_ba_p2p_signal increments it on the stream), so the whole iteration can be captured in a CUDA graph
   * and replayed; p2p_epoch then only selects the slot, and a captured graph must hold an even number of iterations. */
  int p2p_world, p2p_rank;
  unsigned long long p2p_epoch;
  void* p2p_system[8];
  unsigned long long* p2p_epoch_dev;
} dba_ba_args;

int dba_ba_prepare(const dba_ba_args* a);
int dba_ba_build(const dba_ba_args* a);
int dba_ba_solve(const dba_ba_args* a);
int dba_ba(const dba_ba_args* a, int iterations);
int dba_ba_p2p_signal(const dba_ba_args* a);   /* after dba_ba_build, before dba_ba_solve, when p2p_world > 1 */
/* synchronises `stream` and reads back M = number of depth frames found by the last dba_ba_prepare on this
 * workspace and the sticky device status word (0 = ok, bit0 = index out of range, bit1 = eta rows != M,
 * bit2 = Cholesky hit a non-positive pivot in some iteration -> that iteration's dx = 0 like the reference,
 * bit3 = a source frame has more than 254 out-edges: its Schur complement would be truncated, the result is not usable). */
int dba_ba_read_info(const dba_ba_args* a, int* n_depth_frames, int* device_status);

/* ---- update operator (ConvGRU + heads + GraphAgg) on the tensor cores -----------------------------------
 * replaces UpdateModule.forward (reference droid_slam/droid_net.py:111-143), ConvGRU.forward (droid_slam/modules/gru.py:19-32) and
 * GraphAgg.forward (droid_net.py:59-75) -- in the reference a chain of 19 cuDNN convolutions + ~25 elementwise launches.
 * Every convolution runs as an implicit GEMM (tcgen05 / TMEM / TMA, f16 operands, fp32 accumulation) on channels-last
 * activations; gates, activations, the global-context sum and output layouts are fused into the epilogues.
 *
 * Packed weights (device memory, made once per checkpoint by the host side, droid_slam_b200/update.py:pack_update_weights):
 *   w_* : f16 [taps][N][Kpad]  (tap = dy*k + dx, K = input channels in the reference's concatenation order, zero padded to a
 *         multiple of 64), b_* : f32 [N]
 *   w_corr0 [1][128][256]  corr_encoder.0 (196 in)        w_corr2 [9][128][128] corr_encoder.2
 *   w_flow0 [1][128][256]  flow_encoder.0, K = (dy*7+dx)*4 + c (7x7 taps folded into K)     w_flow2 [9][64][128] flow_encoder.2
 *   w_gate  [1][128][128]  gru.w          w_glo f32 [384][128] = gru.convz_glo | convr_glo | convq_glo, b_glo [384]
 *   w_zr    [9][256][448]  gru.convz | gru.convr         w_q [9][128][448] gru.convq
 *   w_stem  [9][384][128]  delta.0 | weight.0 | agg.conv1
 *   w_heads [1][64][256]   per-tap rows: row 4t+o (t = 3 dy + dx) = tap t of delta.2 (o = 0,1; K 0..127) / weight.2 (o = 2,3; K 128..255),
 *                          rows 36..63 zero -- the 3x3 / 2-channel heads run as one 1x1 convolution + a 9-tap gather;  b_heads [4]
 *   w_agg2  [9][128][128]  agg.conv2      w_eta [1][32][128] row t = tap t of agg.eta.0, b_eta [1]      w_upmask [1][576][128] agg.upmask.0
 *   b_zero  f32 [64] zeros (bias of the per-tap partial-sum convolutions) */
typedef struct {
  const void *w_corr0, *w_corr2, *w_flow0, *w_flow2, *w_gate, *w_zr, *w_q, *w_stem, *w_heads, *w_agg2, *w_eta, *w_upmask;
  const float *b_corr0, *b_corr2, *b_flow0, *b_flow2, *b_gate, *b_zr, *b_q, *b_stem, *b_heads, *b_agg2, *b_eta, *b_upmask;
  const float *w_glo, *b_glo, *b_zero;
} dba_update_weights;

typedef struct {
  int n_edges, ht, wd;
  const void* net;  int net_dtype;  int net_layout;   /* hidden state [E,128,ht,wd] (layout 0, DBA_F16 / DBA_F32) or channels-last f16 [E,ht,wd,128] (layout 1) */
  const void* inp;  int inp_dtype;                     /* context features [E,128,ht,wd] */
  const void* corr; int corr_dtype;                    /* correlation features [E,196,ht,wd] */
  const float* flow;                                   /* motion features [E,4,ht,wd] f32, or NULL (= zeros, MotionFilter's call) */
  const int64_t* seg; int n_src;                       /* seg[e] = rank of edge e's source frame among the distinct sources (torch.unique
                                                          inverse); n_src = number of distinct sources; n_src = 0: no aggregation outputs */
  const dba_update_weights* weights;                   /* HOST struct of DEVICE pointers */
  void* net_out;                                       /* new hidden state, channels-last f16 [E,ht,wd,128] */
  float* delta;  float* weight;                        /* [E,ht,wd,2] f32: flow revision, confidence (sigmoid) */
  float* eta;    void* upmask;                         /* [n_src,ht,wd] f32 (0.01 * softplus), [n_src,576,ht,wd] f16 */
  void* workspace; size_t workspace_bytes;             /* dba_update_workspace_bytes(), 256-byte aligned */
  dba_stream_t stream;
} dba_update_args;

size_t dba_update_workspace_bytes(int n_edges, int n_src, int ht, int wd);
int dba_update_forward(const dba_update_args* a);

/* the building block of dba_update_forward, exported: 1x1 / 3x3 'same' convolution of channels-last f16 activations on the tensor
 * cores.  src0 (+ optional src1, concatenated along channels after src0) [n_images,ht,wd,stride] using channels [0,c); wpk f16
 * [ksize*ksize][n_out][Kpad] with Kpad = 64*ceil(c0/64) + 64*ceil(c1/64), K contiguous; bias f32 [n_out]; out f16
 * [n_images,ht,wd,out_stride] channels [0,n_out) written.  n_out in {32,64,...,256,384}. */
int dba_conv_nhwc(const void* src0, int c0, int stride0, const void* src1, int c1, int stride1, const void* wpk, const float* bias,
                  void* out, int out_stride, int n_images, int ht, int wd, int ksize, int n_out, int relu, dba_stream_t stream);

/* ---- standalone damped SPD solve (the solver inside dba_ba_solve) ---------------------------------------
 * (H + diag(ep + lm*diag(H))) x = b with H [n,n] fp64 (full symmetric), b [n] fp64 -> x [n] fp32, on the device in
 * fp64; replaces SparseBlock::solve (reference src/droid_kernels.cu:1201-1222).  *fail_flag_device is set to 1 and x to 0
 * when a pivot is not positive (reference: solver.info() != Eigen::Success). */
size_t dba_solve_workspace_bytes(int n);
int dba_solve_spd(const double* H, const double* b, int n, float lm, float ep, float* x, int* fail_flag_device,
                  void* workspace, size_t workspace_bytes, dba_stream_t stream);
/* host only: where the resident-tile Cholesky (n <= 448) places its tiles -- map_i / map_j [128]: tile (i, j) of warp slot 8*cta + warp
 * (i == number of tile rows: a right-hand-side piece; 0xFF: none).  Returns the cluster size, 0 when n is served by the barrier kernel. */
int dba_solve_tile_placement(int n, unsigned char* map_i, unsigned char* map_j);

#ifdef __cplusplus
}
#endif
#endif /* DROID_B200_H */
