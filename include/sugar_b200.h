/*
 * sugar_b200.h -- C ABI of libsugar_b200.so: a Blackwell (sm_100a) differentiable
 * Gaussian-splat rasterizer + SuGaR density/SDF field evaluator.
 *
 * Drop-in boundary.  These entry points are what a binding of the reference's rasterizer
 * path would call instead of `CudaRasterizer::Rasterizer::{forward,backward,markVisible}`
 * (gaussian_splatting/submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:20-85),
 * and instead of the PyTorch ops behind `SuGaR.get_field_values` / `SuGaR.compute_density`
 * (sugar_scene/sugar_model.py:1247-1316, 1345-1368).  Plain pointers and sizes only: every
 * pointer is a DEVICE pointer unless stated otherwise; the library allocates nothing that
 * outlives a call except one pinned 4 KiB page + one event per device (the num_rendered
 * read-back).  All work is enqueued on the caller's `stream` (a cudaStream_t passed as void*).
 *
 * Every function returns 0 on success or a negative SGR_E* code; sgr_last_error() gives the
 * thread-local message.  CUDA launch errors surface synchronously only when `debug` is set
 * (reference: CHECK_CUDA, cuda_rasterizer/auxiliary.h:166-173).
 */
#ifndef SUGAR_B200_H_
#define SUGAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SGR_API __attribute__((visibility("default")))
#else
#define SGR_API
#endif

#define SGR_OK 0
#define SGR_EINVAL (-1)   /* bad argument (shape, null pointer, ambiguous optionals) */
#define SGR_ECUDA (-2)    /* CUDA runtime error */
#define SGR_ENOMEM (-3)   /* an allocator callback returned NULL */

/* Per-view constants: field-for-field `GaussianRasterizationSettings`
 * (diff_gaussian_rasterization/__init__.py:157-169). */
typedef struct SgrView {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    const float *bg;         /* f32[3] */
    float scale_modifier;
    const float *viewmatrix; /* f32[16], world->view transposed (read column-major) */
    const float *projmatrix; /* f32[16], full projection transposed */
    int32_t sh_degree;       /* active degree D, 0..3 */
    const float *campos;     /* f32[3] */
    int32_t prefiltered;
    int32_t debug;
} SgrView;

/* Per-Gaussian inputs of `_C.rasterize_gaussians` (rasterize_points.h:19-38).  Optional
 * arrays are NULL when absent, exactly as the reference tests `ptr == nullptr`
 * (forward.cu:205,241): exactly one of {shs, colors_precomp} and exactly one of
 * {scales+rotations, cov3D_precomp} must be non-NULL. */
typedef struct SgrGaussians {
    int32_t P;                   /* number of Gaussians */
    int32_t M;                   /* SH coefficients stored per Gaussian (0 if shs == NULL) */
    const float *means3D;        /* f32[P,3] */
    const float *opacities;      /* f32[P]   */
    const float *shs;            /* f32[P,M,3] or NULL */
    const float *colors_precomp; /* f32[P,3]   or NULL */
    const float *scales;         /* f32[P,3]   or NULL */
    const float *rotations;      /* f32[P,4] (w,x,y,z; pre-normalised by the caller) or NULL */
    const float *cov3D_precomp;  /* f32[P,6]   or NULL */
    /* Raw-parameter mode (0 = the reference's contract above).  SGR_ACT_RAW: the arrays are a SuGaR model's raw
     * parameters and its activations (sugar_scene/sugar_model.py:400-479) are applied in the kernels, their chain
     * rule in the backward: opacities = all_densities (sigmoid), scales = _scales (exp), rotations = _quaternions
     * (normalize), and the SH coefficients stay in the model's two arrays: shs = _sh_coordinates_dc f32[P,1,3],
     * sh_rest = _sh_coordinates_rest f32[P,M-1,3].  The backward's dL_dopacity / dL_dscales / dL_drotations are
     * then gradients of the RAW parameters, dL_dsh is f32[P,1,3] and dL_dsh_rest (SgrBackwardPlan) f32[P,M-1,3]. */
    int32_t activations;
    const float *sh_rest;        /* f32[P,M-1,3], raw mode only */
} SgrGaussians;
#define SGR_ACT_RAW 1

/* Scratch allocator: replaces `std::function<char*(size_t)>` (rasterizer.h:31-34,
 * rasterize_points.cu:27-33).  Must return device memory aligned to >= 256 bytes that stays
 * valid until the matching backward has run.  The three buffers are opaque round-trip state. */
typedef void *(*SgrAlloc)(void *ctx, size_t bytes);

/* Forward: replaces Rasterizer::forward (rasterizer_impl.cu:198-336).
 *   out_color  f32[3,H,W]  (fully written)
 *   radii      i32[P]      (fully written; >0 == visible)
 *   capacity_hint: 0 -> wait for the instance count before sizing the binning buffer (what the
 *     reference does, rasterizer_impl.cu:281); >0 -> allocate that many instances up front and
 *     enqueue everything without a host wait, re-running binning+blend only on overflow.
 *   *num_rendered receives R (number of (Gaussian, tile) instances). */
SGR_API int sgr_rasterize_forward(const SgrView *view, const SgrGaussians *g,
                          SgrAlloc geom_alloc, void *geom_ctx,
                          SgrAlloc binning_alloc, void *binning_ctx,
                          SgrAlloc image_alloc, void *image_ctx,
                          float *out_color, int32_t *radii,
                          int64_t capacity_hint, int64_t *num_rendered, void *stream);

/* Backward: replaces Rasterizer::backward (rasterizer_impl.cu:340-434) together with the
 * zero-initialised outputs of RasterizeGaussiansBackwardCUDA (rasterize_points.cu:151-159):
 * every output row is written by the kernels (zeros for culled Gaussians), so the caller may
 * pass uninitialised memory.  `grad_scratch` must hold sgr_backward_scratch_bytes(P).
 *   dL_dmeans2D f32[P,3], dL_dcolors f32[P,3], dL_dopacity f32[P,1], dL_dmeans3D f32[P,3],
 *   dL_dcov3D f32[P,6], dL_dsh f32[P,M,3], dL_dscales f32[P,3], dL_drotations f32[P,4].
 * dL_dcolors is the gradient of colors_precomp when colours were precomputed; with SH colours it is the
 * clamp-masked dL/dRGB per Gaussian (the reference's value is unmasked there, but no caller can observe
 * it: colors_precomp is absent).  dL_dsh may be NULL: when M == 0, or with SH present to select "factor
 * mode" -- dL_dsh is not produced, and sgr_view_grad_finalize / sgr_sh_grad_from_factors rebuild (the sum over
 * views of) dL_dsh from dL_dcolors.  All other outputs are unchanged.  Output pointers that are 16-byte aligned
 * leave through TMA bulk stores. */
SGR_API int sgr_rasterize_backward(const SgrView *view, const SgrGaussians *g, const int32_t *radii,
                           const void *geom_buffer, const void *binning_buffer, const void *image_buffer,
                           int64_t num_rendered, const float *dL_dout_color,
                           float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity,
                           float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh,
                           float *dL_dscales, float *dL_drotations,
                           void *grad_scratch, void *stream);

/* The same backward with a plan: stage hooks, a chunked per-Gaussian pass and record-form outputs.  This is
 * what the view-parallel (one view per GPU) step drives, sugar_b200/parallel.py; the reference has no
 * counterpart (it renders one view on one GPU, sugar_trainers/coarse_sdf.py:98,507).
 *
 * hook(hook_ctx, stage) is called on the HOST, between kernel enqueues on `stream`:
 *   SGR_STAGE_BLEND_DONE        the blend pass is enqueued: everything on `stream` so far makes dL_dcolors final.
 *                               With SH colours dL_dcolors holds the clamp-masked dL/dRGB per Gaussian -- this
 *                               view's rank-1 SH factor: a caller that exchanges factors instead of dL_dsh
 *                               (dL_dsh == NULL, "factor mode") starts its all-gather here, underneath the
 *                               per-Gaussian pass that is enqueued next.
 *   SGR_STAGE_CHUNK_DONE + c    chunk c of the per-Gaussian pass is enqueued: its rows of every per-Gaussian
 *                               output are final on `stream`; the caller can start reducing them while the
 *                               next chunk is computed.
 * num_chunks (<= 1: one) splits the per-Gaussian pass into Gaussian ranges; sgr_backward_chunk_range() gives
 * the range [p0, p1) of a chunk.
 * reduce_records (optional, f32[P,11]): when non-NULL the pass writes (dL_dmeans3D 3 | dL_dopacity 1 |
 * dL_dscales 3 | dL_drotations 4) of Gaussian i as ONE 44-byte record there INSTEAD of the four arrays
 * (whose pointers are then ignored), so that a chunk's all-reduce is one contiguous range;
 * sgr_view_grad_finalize() splits the reduced records back into the arrays. */
#define SGR_STAGE_BLEND_DONE 1
#define SGR_STAGE_CHUNK_DONE 16
typedef void (*SgrStageHook)(void *ctx, int32_t stage);
typedef struct SgrBackwardPlan {
    SgrStageHook hook; /* may be NULL */
    void *hook_ctx;
    int32_t num_chunks;
    float *reduce_records; /* may be NULL */
    float *dL_dsh_rest;    /* raw-parameter mode (SgrGaussians.activations): f32[P,M-1,3], fully written */
    /* Peer mode (sgr_peer_* below).  When peer_flag_tab is non-NULL the backward signals the peer ranks itself, on
     * `stream`, with sequence number peer_seq: slot peer_slot_blend once the blend pass is enqueued (dL_dcolors, which
     * the caller placed in its peer-visible factor block, is final), slot peer_slot_chunk0 + c after chunk c of the
     * per-Gaussian pass (reduce_records, placed in the peer-visible record array, holds the chunk's rows). */
    void *const *peer_flag_tab; /* DEVICE array of peer_nranks pointers: every rank's flag words, see sgr_peer_signal */
    int32_t peer_nranks, peer_rank, peer_slot_blend, peer_slot_chunk0;
    uint32_t peer_seq;
    /* With SH colours and dL_dsh non-NULL: peer_view_blocks is a DEVICE array of peer_nranks pointers to the ranks'
     * factor blocks (f32[3P] dL/dRGB followed by the rank's camera position; entry peer_rank is dL_dcolors itself).
     * The backward then waits (on `stream`, through peer_flags = this rank's own flag words) until every rank has
     * signalled peer_slot_blend, and the per-Gaussian pass writes dL_dsh = peer_dsh_scale * SUM over all ranks' views
     * of their SH gradients, loading the other ranks' factors straight from their memory -- the all-gather of the
     * factors is fused into the kernel and no separate SH epilogue runs. */
    const float *const *peer_view_blocks;
    const void *peer_flags;
    float peer_dsh_scale;
    double peer_timeout_s;
    int32_t chunk_taper; /* 0: equal chunks; else every chunk half the size of the one before it */
    /* With peer_view_blocks: DEVICE array of peer_nranks pointers, entry j = rank j's staging array (f32[11P + 4]) for
     * records COMING FROM this rank (entry peer_rank = reduce_records itself).  CTA b of a chunk's launch then stores its
     * 64 records into the array of the rank that owns them (owner = b * nranks / CTAs of the chunk; what
     * sgr_peer_reduce_records assumes) with a TMA bulk store over NVLink -- the reduce-scatter half of the records'
     * all-reduce is fused into the pass.  NULL: all records go to reduce_records. */
    float *const *peer_record_stages;
    /* Optional second stream for the backward's own signals: each signal kernel then runs there, behind an event
     * recorded on `stream`, and the per-Gaussian chunks follow one another on `stream` without the signals' launches
     * in between.  The caller orders whatever must see the signals' completion behind this stream. */
    void *peer_signal_stream;
    /* Optional: the backward also drives the records' exchange, on this side stream, chunk by chunk underneath the
     * per-Gaussian pass of the later chunks: once this rank's CHUNK c signal is out (an event behind the signal kernel;
     * a spinning kernel must never sit in front of a signal another rank waits for) it enqueues
     * sgr_peer_reduce_records_synced (wait CHUNK c of every rank, sum the owned slice over peer_rec_tab into every rank's
     * peer_sum_tab, signal peer_slot_reduced0 + c) and, behind the NEXT chunk's reduce, sgr_peer_wait(REDUCED c) + the
     * split of peer_sums (this rank's sum array) into dL_dmeans3D / dL_dopacity / dL_dscales / dL_drotations scaled by
     * peer_dsh_scale.  The caller orders its stream behind peer_side_stream (and peer_signal_stream) afterwards.
     * peer_emulate_ranks (tests): bit r set = also reduce rank r's slices here. */
    void *peer_side_stream;
    const void *const *peer_rec_tab;
    void *const *peer_sum_tab;
    const float *peer_sums;
    int32_t peer_slot_reduced0;
    uint64_t peer_emulate_ranks;
} SgrBackwardPlan;
SGR_API int sgr_rasterize_backward_staged(const SgrView *view, const SgrGaussians *g, const int32_t *radii,
                                          const void *geom_buffer, const void *binning_buffer, const void *image_buffer,
                                          int64_t num_rendered, const float *dL_dout_color, float *dL_dmeans2D,
                                          float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D, float *dL_dcov3D,
                                          float *dL_dsh, float *dL_dscales, float *dL_drotations, void *grad_scratch,
                                          void *stream, const SgrBackwardPlan *plan);
SGR_API int sgr_backward_chunk_range(int32_t P, int32_t num_chunks, int32_t chunk, int32_t *p0, int32_t *p1);
SGR_API int sgr_backward_chunk_range_tapered(int32_t P, int32_t num_chunks, int32_t chunk, int32_t taper, int32_t *p0,
                                             int32_t *p1); /* SgrBackwardPlan.chunk_taper */

/* Epilogue of the view-parallel step for the Gaussians [p0, p1), after the exchange (either half optional):
 *  - dL_dsh[P,M,3] rows = sum over views v of basis_k(normalize(mean - campos[v])) * dRGB[v][P,3]  (the SH part of
 *    backward.cu:20-139 is an outer product per view): the gathered factors of view v start at dRGB + v * view_stride
 *    floats (f32[P,3]; view_stride >= 3P) and its camera position at campos + v * campos_stride floats -- the
 *    exchange appends every view's camera position to its factor block so that one all-gather carries both;
 *  - reduced_records f32[P,11] (all-reduced) are split into dL_dmeans3D / dL_dopacity / dL_dscales / dL_drotations.
 * Every output row of the range is fully written; `scale` (e.g. 1/num_views) multiplies everything. */
SGR_API int sgr_view_grad_finalize(int32_t P, int32_t p0, int32_t p1, int32_t M, int32_t sh_degree, int32_t num_views,
                                   const float *means3D, const float *campos, const float *dRGB, int64_t view_stride,
                                   int32_t campos_stride, float *dL_dsh, const float *reduced_records, float scale,
                                   float *dL_dmeans3D, float *dL_dopacity, float *dL_dscales, float *dL_drotations,
                                   void *stream);
/* The same epilogue reading every view's factor block through a DEVICE-resident table of num_views pointers
 * (view_blocks[v] -> f32[3P] factors followed by that view's camera position, 3 floats).  The entries may point into
 * other GPUs' memory mapped with sgr_peer_import: the all-gather of the factors is then fused into this kernel's loads
 * over NVLink and no gathered copy exists. */
SGR_API int sgr_view_grad_finalize_peers(int32_t P, int32_t p0, int32_t p1, int32_t M, int32_t sh_degree,
                                         int32_t num_views, const float *means3D, const float *const *view_blocks,
                                         float *dL_dsh, const float *reduced_records, float scale, float *dL_dmeans3D,
                                         float *dL_dopacity, float *dL_dscales, float *dL_drotations, void *stream);

/* ---- View-parallel exchange over peer memory (NVLink / NVSwitch): csrc/sgr_peer.cu.  No counterpart in the reference.
 * sgr_peer_alloc     one zero-filled device allocation on the current device whose CUDA IPC handle (sgr_peer_export,
 *                    64 bytes, to be sent to the other ranks by any means) maps it at offset 0;
 * sgr_peer_import    maps a peer's allocation into this process (lazy peer access); sgr_peer_close unmaps it.
 * Flag words: the first sgr_peer_flag_bytes() of an allocation used with signal / wait are u32
 * [SGR_PEER_MAX_SLOTS][SGR_PEER_MAX_RANKS]; word [slot][j] is written by rank j only.
 * sgr_peer_signal    (on `stream`) writes `seq` into word [slot][my_rank] of EVERY rank's flags (flag_tab: device array
 *                    of nranks pointers to the ranks' flag words), after a system-scope fence;
 * sgr_peer_wait      (on `stream`) spins until words [slot0 .. slot0+nslots)[0 .. nranks) of the LOCAL flags have all
 *                    reached seq (wrap-safe); traps after timeout_s seconds (<= 0: 20 s) instead of hanging;
 * sgr_peer_reduce_records  second half of the records' all-reduce for the chunk of Gaussians [p0, p1) (p0 a multiple of
 *                    64): this rank sums the slice of the chunk it OWNS -- the 64-record blocks b with
 *                    b * nranks / blocks == my_rank, the ones the per-Gaussian pass stored here
 *                    (SgrBackwardPlan.peer_record_stages) -- over the nranks staging arrays rec_tab points to (device
 *                    array of pointers to f32[11P + 4]; normally all LOCAL), in rank order, and stores the sums into
 *                    every rank's sum array (sum_tab, likewise; remote entries are posted NVLink writes). */
#define SGR_PEER_MAX_RANKS 64
#define SGR_PEER_MAX_SLOTS 64
SGR_API int sgr_peer_alloc(size_t bytes, void **ptr);
SGR_API int sgr_peer_free(void *ptr);
SGR_API int sgr_peer_export(const void *ptr, void *handle64);
SGR_API int sgr_peer_import(const void *handle64, void **ptr);
SGR_API int sgr_peer_close(void *ptr);
SGR_API size_t sgr_peer_flag_bytes(void);
SGR_API int sgr_peer_signal(void *const *flag_tab, int32_t nranks, int32_t slot, int32_t my_rank, uint32_t seq,
                            void *stream);
SGR_API int sgr_peer_wait(const void *flags, int32_t nranks, int32_t slot0, int32_t nslots, uint32_t seq,
                          double timeout_s, void *stream);
SGR_API int sgr_peer_reduce_records(const void *const *rec_tab, void *const *sum_tab, int32_t nranks, int32_t my_rank,
                                    int32_t p0, int32_t p1, void *stream);
/* The same with the flag handshake folded into the kernel: it first waits until every rank's word of wait_slot in the
 * LOCAL flags has reached seq (flags NULL: no wait), and its last CTA to finish writes seq into word
 * [signal_slot][my_rank] of every rank's flags (flag_tab NULL: no signal; `counter`: a zero-initialised u32 in local
 * device memory that the kernel re-arms, one per concurrently running launch). */
SGR_API int sgr_peer_reduce_records_synced(const void *const *rec_tab, void *const *sum_tab, int32_t nranks,
                                           int32_t my_rank, int32_t p0, int32_t p1, const void *flags, int32_t wait_slot,
                                           void *const *flag_tab, int32_t signal_slot, uint32_t seq, void *counter,
                                           double timeout_s, void *stream);

/* The SH half alone, over all Gaussians (kept for callers that only exchange factors). */
SGR_API int sgr_sh_grad_from_factors(int32_t P, int32_t M, int32_t sh_degree, int32_t num_views, const float *means3D,
                                     const float *campos, const float *dRGB, float *dL_dsh, void *stream);

/* markVisible (rasterizer_impl.cu:141-153): present[i] = view-space z > 0.2. */
SGR_API int sgr_mark_visible(int32_t P, const float *means3D, const float *viewmatrix,
                     const float *projmatrix, uint8_t *present, void *stream);

/* Sizes of the opaque buffers (what `required<T>(P)` is in rasterizer_impl.h:66-72). */
SGR_API size_t sgr_geometry_bytes(int32_t P);
SGR_API size_t sgr_binning_bytes(int64_t capacity);
SGR_API size_t sgr_image_bytes(int32_t width, int32_t height);
SGR_API size_t sgr_backward_scratch_bytes(int32_t P);

/* Decode the opaque buffers into the reference's named arrays, for parity tests
 * (GeometryState / BinningState / ImageState, rasterizer_impl.h:30-63).  Any output may be NULL.
 *   depths f32[P], means2D f32[P,2], conic_opacity f32[P,4], rgb f32[P,3], clamped u8[P,3],
 *   tiles_touched u32[P]; keys u64[R] = (tile<<32 | depth bits) in sorted order,
 *   point_list u32[R]; ranges u32[T,2]; final_T f32[H,W]; n_contrib u32[H,W];
 *   footprint u8[R]: per sorted instance, which of its tile's eight 8x4-pixel blocks (bit = band*2 + half)
 *   the splat can reach with alpha >= 1/255 -- this library's own culling state, not a reference array
 *   (0xff when the list carries no masks, i.e. P > 2^24). */
SGR_API int sgr_inspect_state(int32_t P, int32_t width, int32_t height, int64_t num_rendered,
                      const void *geom_buffer, const void *binning_buffer, const void *image_buffer,
                      float *depths, float *means2D, float *conic_opacity, float *rgb, uint8_t *clamped,
                      uint32_t *tiles_touched, uint64_t *keys, uint32_t *point_list, uint32_t *ranges,
                      float *final_T, uint32_t *n_contrib, uint8_t *footprint, void *stream);

/* ------------------------------------------------------------------------------------------
 * SuGaR surface-regularisation field: fused K-neighbour gather + anisotropic density / SDF.
 * Replaces the PyTorch ops of SuGaR.get_field_values (sugar_model.py:1247-1316, beta_mode
 * 'average' sugar_model.py:1192-1195) and SuGaR.compute_density (:1345-1368).
 *   x          f32[N,3]   sample points
 *   nbr_idx    i64[N,K]   neighbour Gaussian indices per sample (knn_idx[gaussian_idx])
 *   points f32[P,3], scaling f32[P,3] (activated), quaternions f32[P,4] (normalised, w first),
 *   strengths f32[P] (sigmoid(densities))
 * Outputs (any may be NULL): density f32[N] (before the straight-through clamp),
 *   nbr_opacity f32[N,K], beta f32[N], sdf f32[N].
 * ------------------------------------------------------------------------------------------ */
typedef struct SgrFieldParams {
    int32_t N, K, P;
    float density_factor;
    float density_threshold;
    float opacity_min_clamp; /* 1e-16 in the reference */
    int32_t samples_per_idx_row; /* 0/1: nbr_idx is i64[N,K]; g > 1: i64[ceil(N/g),K], row n/g serves sample n
                                  * (ray samples that share their pixel's neighbours, sugar_model.py:1980) */
} SgrFieldParams;

/* `scratch` must hold sgr_field_scratch_bytes(P) (packed per-Gaussian records; device memory). */
SGR_API size_t sgr_field_scratch_bytes(int32_t P);

SGR_API int sgr_field_forward(const SgrFieldParams *p, const float *x, const int64_t *nbr_idx,
                      const float *points, const float *scaling, const float *quaternions,
                      const float *strengths,
                      float *density, float *nbr_opacity, float *beta, float *sdf,
                      void *scratch, void *stream);

/* Backward of the above.  Upstream grads (any may be NULL = zero): g_density f32[N],
 * g_nbr_opacity f32[N,K], g_beta f32[N], g_sdf f32[N].  Downstream grads are fully WRITTEN
 * (any may be NULL): g_x f32[N,3], g_points f32[P,3], g_scaling f32[P,3],
 * g_quaternions f32[P,4] (w.r.t. the quaternion passed in, through two_s = 2/|q|^2 exactly as
 * pytorch3d's quaternion_to_matrix differentiates), g_strengths f32[P]. */
SGR_API int sgr_field_backward(const SgrFieldParams *p, const float *x, const int64_t *nbr_idx,
                       const float *points, const float *scaling, const float *quaternions,
                       const float *strengths,
                       const float *g_density, const float *g_nbr_opacity, const float *g_beta,
                       const float *g_sdf,
                       float *g_x, float *g_points, float *g_scaling, float *g_quaternions,
                       float *g_strengths, void *scratch, void *stream);

/* ------------------------------------------------------------------------------------------
 * "Better normal" regularisation of the trainers (sugar_trainers/coarse_sdf.py:688-716, with
 * SuGaR.get_normals(estimate_from_points=False) -> get_smallest_axis, sugar_model.py:930-968), in the
 * trainers' only mode sdf_better_normal_gradient_through_normal_only=True (weights and signs detached):
 *   n_g = column argmin_j scaling[g][j] of R(quaternions[g]);  m_k = sign(<n_k, n_own>) n_k
 *   w_k = nbr_opacity[n,k] |<x_n - mu_k, m_k>| / max(min_j scaling[k][j], 1e-6)^2, / max(sum_k w_k, 1e-6)
 *   loss[n] = | n_own - sum_k w_k m_k |^2
 * Inputs: x f32[N,3], own_idx i64[N] (sdf_gaussian_idx), nbr_idx i64[N,K] (knn_idx[own_idx]),
 * points/scaling f32[P,3], quaternions f32[P,4], nbr_opacity f32[N,K] (fields['closest_gaussian_opacities']).
 * Backward: g_loss f32[N] -> g_quaternions f32[P,4], fully written (the only differentiable input).
 * `scratch` must hold sgr_normal_scratch_bytes(P).
 * ------------------------------------------------------------------------------------------ */
SGR_API size_t sgr_normal_scratch_bytes(int32_t P);
SGR_API int sgr_normal_loss_forward(int32_t N, int32_t K, int32_t P, const float *x, const int64_t *own_idx,
                                    const int64_t *nbr_idx, const float *points, const float *scaling,
                                    const float *quaternions, const float *nbr_opacity, float *loss, void *scratch,
                                    void *stream);
SGR_API int sgr_normal_loss_backward(int32_t N, int32_t K, int32_t P, const float *x, const int64_t *own_idx,
                                     const int64_t *nbr_idx, const float *points, const float *scaling,
                                     const float *quaternions, const float *nbr_opacity, const float *g_loss,
                                     float *g_quaternions, void *scratch, void *stream);

/* ------------------------------------------------------------------------------------------
 * Gaussians bound to a triangle mesh (refinement stage): the property code of a bound SuGaR model,
 * sugar_scene/sugar_model.py:384-398 (points), :415-441 (scaling), :443-479 (quaternions), as one kernel
 * and its backward.  F faces, n_per Gaussians per face (P = F * n_per), V vertices.
 *   verts f32[V,3], faces i64[F,3], bary f32[n_per,3] (surface_triangle_bary_coords, device memory),
 *   scales_raw f32[P,2] (_scales: log of the two in-plane scales), complex_raw f32[P,2] (_quaternions: the
 *   learned in-plane rotation as a complex number, normalised here), thickness = surface_mesh_thickness.
 * Outputs, fully written: points f32[P,3], scaling f32[P,3] = (thickness, exp, exp), quaternions f32[P,4]
 * (normalised, w first; pytorch3d 0.7.4 matrix_to_quaternion of [normal | rotated edge | cross]).
 * Backward: g_verts f32[V,3] (zeroed here, then accumulated), g_scales_raw, g_complex_raw f32[P,2] fully written.
 * ------------------------------------------------------------------------------------------ */
SGR_API int sgr_meshbind_forward(int32_t F, int32_t n_per, int32_t V, const float *verts, const int64_t *faces,
                                 const float *bary, const float *scales_raw, const float *complex_raw, float thickness,
                                 float *points, float *scaling, float *quaternions, void *stream);
SGR_API int sgr_meshbind_backward(int32_t F, int32_t n_per, int32_t V, const float *verts, const int64_t *faces,
                                  const float *bary, const float *scales_raw, const float *complex_raw,
                                  const float *g_points, const float *g_scaling, const float *g_quaternions,
                                  float *g_verts, float *g_scales_raw, float *g_complex_raw, void *stream);

/* ------------------------------------------------------------------------------------------
 * Exact K-nearest-neighbour search (uniform grid).  Replaces pytorch3d.ops.knn_points as SuGaR
 * calls it: reset_neighbors (sugar_model.py:1013-1030, queries == points, K = 16) and
 * get_gaussians_closest_to_samples (sugar_model.py:1335-1343).
 *   points f32[P,3] reference cloud, queries f32[Q,3]; 0 < K <= min(P, 64)
 *   idx i64[Q,K], dist2 f32[Q,K]: neighbours ordered by increasing squared distance.
 * `workspace` must hold sgr_knn_workspace_bytes(P).
 * ------------------------------------------------------------------------------------------ */
SGR_API size_t sgr_knn_workspace_bytes(int32_t P);
SGR_API int sgr_knn(int32_t P, const float *points, int32_t Q, const float *queries, int32_t K, int64_t *idx,
                    float *dist2, void *workspace, void *stream);

/* Measurement hooks (bench.py): number of kernels this library has launched in this process;
 * optional CUDA-event bracketing of every launch on its own stream.  sgr_profile_read fills
 * total_ms[kind] / counts[kind] for kind < sgr_num_kernel_kinds() and resets the log. */
SGR_API unsigned long long sgr_launch_count(void);
SGR_API int sgr_num_kernel_kinds(void);
SGR_API const char *sgr_kernel_name(int kind);
SGR_API int sgr_profile_enable(int on);
SGR_API int sgr_profile_read(float *total_ms, int *counts);
SGR_API int sgr_profile_timeline(int max_records, int *kinds, float *t_begin_ms, float *t_end_ms);

/* sizeof of an ABI struct as the library was compiled (0 SgrView, 1 SgrGaussians, 2 SgrBackwardPlan, 3 SgrFieldParams;
 * else 0): lets a foreign-language binding check its own struct layout. */
SGR_API size_t sgr_struct_bytes(int32_t which);
SGR_API const char *sgr_last_error(void);
SGR_API const char *sgr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SUGAR_B200_H_ */
