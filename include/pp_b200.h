/*
 * pp_b200.h -- C ABI of libpp_b200.so: the B200 (sm_100a) replacement for the
 * data-parallel hot path of yifita/pytorch_points.
 *
 * Each entry point replaces one function of the reference's pybind11 extension
 * modules `pytorch_points._ext.losses` / `pytorch_points._ext.sampling` (cited
 * per function, paths relative to /root/reference/pytorch_points/).  No torch
 * types cross this boundary: plain device pointers, sizes, a device ordinal and
 * a CUDA stream handle (`void*` == cudaStream_t; NULL = legacy default stream).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major memory on
 *     `device`; the library never allocates, frees or retains caller memory;
 *   - all launches are asynchronous on `stream`;
 *   - return value: PP_OK (0) on success, a negative PP_E* code for argument
 *     errors, or a positive cudaError_t; pp_last_error_string() describes the
 *     last failure on the calling thread.  Nothing ever calls exit() (the
 *     reference does: _ext/sampling_cuda.cu:40-44) or prints-and-continues
 *     (_ext/nmdistance_cuda.cu:132-137);
 *   - indices are int32 like the reference's IntTensor outputs.
 */
#ifndef PP_B200_H
#define PP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_OK 0
#define PP_EINVAL (-22)   /* bad size / null pointer / unsupported shape */
#define PP_ENOSPC (-28)   /* workspace too small */

/* Library / ABI version (bumped when a signature changes). */
int pp_version(void);
/* Human-readable description of the last error on this thread ("" if none). */
const char *pp_last_error_string(void);

/* ------------------------------------------------------------------ losses */

/* Scratch bytes pp_chamfer_fwd needs for (B,N,M): the larger of the exact one-pass kernel's packed
 * keys (8 per point of either cloud) and the tensor-core sweep's layout (operand tiles of every point
 * in query and in reference role, norms, keys, runner-up words, block masks, rescan lists: about 160
 * per point, clouds padded to multiples of 128). */
size_t pp_chamfer_fwd_workspace_bytes(int B, int N, int M);

/*
 * Chamfer / nndistance forward.  Replaces losses.nmdistance_forward
 * (_ext/nmdistance.cpp:13-15,31 -> _ext/nmdistance_cuda.cu:118-140, kernel :8-49).
 *   xyz1 (B,N,c) xyz2 (B,M,c) fp32 -> dist1 (B,N) dist2 (B,M) squared distances,
 *   idx1 (B,N) idx2 (B,M) int32 argmin, lowest index on ties.
 *   sums: optional (may be NULL) 2 floats receiving [sum(dist1), sum(dist2)]
 *   (fused reduction for the loss mean / NCCL all-reduce input).
 *   workspace: >= pp_chamfer_fwd_workspace_bytes(B,N,M) bytes, 8-byte aligned.
 *   flags: 0, or PP_CHAMFER_WS_CLEAN when the first 8*B*(N+M) bytes are all
 *   0xff -- true for a buffer filled once with 0xff and since then only ever used by
 *   successful pp_chamfer_fwd calls on the same stream (each call restores that state).
 */
#define PP_CHAMFER_WS_CLEAN 1
/* Which forward the last pp_chamfer_fwd / pp_chamfer_fwd_bwd_uniform call on this thread ran (diagnostics;
 * bench.py labels its roofline with it): 0 = exact FFMA one-pass kernel (or the generic c != 3 kernel),
 * 1 = tensor-core sweep (tcgen05.mma kind::tf32 into TMEM) + exact resolution.  Same results either way. */
int pp_chamfer_last_path(void);
int pp_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int N, int M, int c,
                   float *dist1, float *dist2, int32_t *idx1, int32_t *idx2, float *sums,
                   void *workspace, size_t workspace_bytes, int flags, int device, void *stream);

/*
 * Labeled Chamfer forward.  Replaces losses.labeled_nmdistance_forward
 * (_ext/nmdistance.cpp:17-20,32 -> _ext/nmdistance_cuda.cu:56-115,142-166).
 *   label1 (B,N) label2 (B,M) fp32 (the reference casts labels to the point dtype,
 *   network/model_loss.py:452-453); unmatched points get idx -1, dist 0.
 *   workspace / workspace_bytes / flags: as for pp_chamfer_fwd (same size, same 0xff protocol;
 *   the two entry points may share one buffer).  With a workspace and c == 3 the one-pass
 *   kernel runs with a label mask; it returns the reference's results whenever every
 *   same-label distance is below the reference's own 1e10 sentinel (nmdistance_cuda.cu:70).
 *   workspace == NULL (or c != 3) selects the literal chunk-by-chunk restatement of the
 *   reference kernel, quirks included.
 */
int pp_chamfer_labeled_fwd(const float *xyz1, const float *xyz2, const float *label1,
                           const float *label2, int B, int N, int M, int c, float *dist1,
                           float *dist2, int32_t *idx1, int32_t *idx2, void *workspace,
                           size_t workspace_bytes, int flags, int device, void *stream);

/*
 * Chamfer backward.  Replaces losses.nmdistance_backward
 * (_ext/nmdistance.cpp:23-27,33 -> _ext/nmdistance_cuda.cu:169-221).
 *   gradxyz1 (B,N,c) / gradxyz2 (B,M,c) are fully overwritten (no pre-zeroing needed);
 *   entries with idx < 0 contribute nothing (:175).
 */
int pp_chamfer_bwd(const float *xyz1, const float *xyz2, const float *graddist1,
                   const float *graddist2, const int32_t *idx1, const int32_t *idx2, int B, int N,
                   int M, int c, float *gradxyz1, float *gradxyz2, int device, void *stream);

/*
 * Chamfer backward for sum/mean-type losses (extension, no reference counterpart): identical
 * to pp_chamfer_bwd with graddist1[b,i] = gw[0] and graddist2[b,j] = gw[1] for all points,
 * gw being a 2-float DEVICE vector (the upstream gradient of [sum(dist1), sum(dist2)]).
 * Saves materialising the two constant graddist arrays.
 */
int pp_chamfer_bwd_uniform(const float *xyz1, const float *xyz2, const float *gw,
                           const int32_t *idx1, const int32_t *idx2, int B, int N, int M, int c,
                           float *gradxyz1, float *gradxyz2, int device, void *stream);

/*
 * Chamfer forward AND backward for sum/mean-type losses in two launches (extension, no reference
 * counterpart): pp_chamfer_fwd followed by pp_chamfer_bwd_uniform, with the backward folded into
 * the kernel that resolves the nearest-neighbour indices.  c == 3.  Outputs as for pp_chamfer_fwd
 * (dist/idx/sums) plus gradxyz1 (B,N,3) / gradxyz2 (B,M,3), fully overwritten; gw = 2-float DEVICE
 * vector d(loss)/d[sum(dist1), sum(dist2)], known before the forward for a mean-type loss.
 * Gradients equal pp_chamfer_bwd_uniform's up to fp32 summation order (both use unordered
 * RED.ADD.F32 like the reference's atomicAdd, _ext/nmdistance_cuda.cu:176-181).
 * workspace / workspace_bytes / flags: as for pp_chamfer_fwd (the entry points may share it).
 */
int pp_chamfer_fwd_bwd_uniform(const float *xyz1, const float *xyz2, const float *gw, int B, int N,
                               int M, float *dist1, float *dist2, int32_t *idx1, int32_t *idx2,
                               float *sums, float *gradxyz1, float *gradxyz2, void *workspace,
                               size_t workspace_bytes, int flags, int device, void *stream);

/* ---------------------------------------------------------------- sampling */

/*
 * Farthest point sampling.  Replaces sampling.furthest_sampling
 * (_ext/sampling.cpp:68-80,207 -> _ext/sampling_cuda.cu:162-325).
 *   xyz (B,N,3); temp (B,N) caller-filled with 1e10 (network/geo_operations.py:33), on
 *   return it holds the final running minima; idx (B,m) int32, idx[:,0] = seed.
 *   Tie-break is the reference's (k mod bs, k) key, bs = min(2^floor(log2 N), 512).
 */
int pp_fps(const float *xyz, int B, int N, int m, int seed, float *temp, int32_t *idx,
           int device, void *stream);

/*
 * Farthest point sampling with the gather fused in ("next" row N2): as pp_fps, and additionally
 * new_xyz (B,m,3) = xyz[b, idx[b,j], :].  Replaces the furthest_sampling -> transpose ->
 * gather_points -> transpose sequence of furthest_point_sample(xyz, m, NCHW=False)
 * (network/geo_operations.py:44-64), which is how PointnetSAModule obtains its centres
 * (network/pointnet2_modules.py:34).
 */
int pp_fps_gather(const float *xyz, int B, int N, int m, int seed, float *temp, int32_t *idx,
                  float *new_xyz, int device, void *stream);

/*
 * Launch shape the last pp_fps / pp_fps_gather call on this thread chose (diagnostics; bench.py's
 * FPS roofline uses it): thread-block cluster width (CTAs, hence SMs, per cloud; 0 = the streaming
 * fallback kernel) and points held per thread.
 */
int pp_fps_last_plan(int *cluster_width, int *points_per_thread);

/*
 * gather_points forward / backward.  Replace sampling.gather_forward / gather_backward
 * (_ext/sampling.cpp:19-41,208-209 -> _ext/sampling_cuda.cu:9-84).
 *   points (B,C,N), idx (B,npoint) -> out (B,C,npoint);
 *   backward ACCUMULATES into grad_points (B,C,N) (caller zero-fills, as
 *   network/operations.py:76-77 does).
 */
int pp_gather_fwd(const float *points, const int32_t *idx, int B, int C, int N, int npoint,
                  float *out, int device, void *stream);
int pp_gather_bwd(const float *grad_out, const int32_t *idx, int B, int C, int N, int npoint,
                  float *grad_points, int device, void *stream);

/*
 * ball_query.  Replaces sampling.ball_query
 * (_ext/sampling.cpp:85-104,210 -> _ext/sampling_cuda.cu:340-398).
 *   new_xyz (B,M,3) centres, xyz (B,N,3) -> idx (B,M,nsample) int32: first nsample
 *   points (ascending index) with d^2 < radius^2, padded with the first hit, all 0 when
 *   the ball is empty.  Every output slot is written (no pre-zeroing needed).
 */
int pp_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                  int nsample, int32_t *idx, int device, void *stream);

/*
 * group_points forward / backward ("next" row N1).  Replace sampling.group_points /
 * group_points_grad (_ext/sampling.cpp:113-161,211-212 -> _ext/sampling_cuda.cu:447-514).
 *   points (B,C,N), idx (B,npoint,nsample) -> out (B,C,npoint,nsample);
 *   backward ACCUMULATES into grad_points (B,C,N).
 */
int pp_group_fwd(const float *points, const int32_t *idx, int B, int C, int N, int npoint,
                 int nsample, float *out, int device, void *stream);
int pp_group_bwd(const float *grad_out, const int32_t *idx, int B, int C, int N, int npoint,
                 int nsample, float *grad_points, int device, void *stream);

/*
 * QueryAndGroup in one kernel ("next" rows N1/N2).  Replaces the ball_query -> transpose ->
 * group_points -> subtract centre -> group_points -> cat sequence of
 * QueryAndGroup.forward (network/operations.py:166-213).
 *   new_xyz (B,M,3), xyz (B,N,3), features (B,C,N) or NULL with C = 0 ->
 *   idx (B,M,nsample) int32 exactly as pp_ball_query,
 *   out (B, 3*use_xyz + C, M, nsample): channels 0..2 = xyz[idx] - new_xyz (one rounded
 *   subtraction), then the C feature channels features[:, :, idx].
 * Backward: grad_out (B, 3*use_xyz + C, M, nsample); ACCUMULATES into grad_features (B,C,N)
 * and grad_xyz (B,N,3) (caller zero-fills; either may be NULL), and WRITES
 * grad_new_xyz (B,M,3) = -sum over samples of the xyz channels (may be NULL).
 */
int pp_query_group_fwd(const float *new_xyz, const float *xyz, const float *features, int B, int N,
                       int M, int C, float radius, int nsample, int use_xyz, int32_t *idx,
                       float *out, int device, void *stream);
int pp_query_group_bwd(const float *grad_out, const int32_t *idx, int B, int N, int M, int C,
                       int nsample, int use_xyz, float *grad_features, float *grad_xyz,
                       float *grad_new_xyz, int device, void *stream);
/*
 * The same forward with the features staged point-major: features_pm (B,N,C), e.g. from
 * pp_channels_to_points (in (B,C,N) -> out (B,N,C)).  Same idx and out, bit for bit; a ball member's
 * channels are then read as full lines instead of one 32-byte sector per value.  A multi-scale level
 * stages once and runs every scale on the copy.  (The backward pass is pp_query_group_bwd either way.)
 */
int pp_query_group_fwd_pm(const float *new_xyz, const float *xyz, const float *features_pm, int B, int N,
                          int M, int C, float radius, int nsample, int use_xyz, int32_t *idx,
                          float *out, int device, void *stream);
int pp_channels_to_points(const float *in, int B, int C, int N, float *out, int device, void *stream);

/* --------------------------------------------------------------------- knn */

/* Scratch bytes pp_knn needs for the path it would choose with the options in force now (pp_set_option):
 * query it after setting them.  pp_knn returns PP_ENOSPC for a workspace that is too small. */
size_t pp_knn_workspace_bytes(int B, int M, int N, int c, int k);

/*
 * group_knn core: k nearest neighbours of every query among `points`.
 * The reference snapshot names the op (README.md:12) but ships no implementation; its
 * callers use pytorch3d.ops.knn_points (network/layers.py:52, network/geo_operations.py:112,139,
 * network/model_loss.py:120,147,378).  Contract (defined by this repo):
 *   query (B,M,c), points (B,N,c), 1 <= k <= min(N, PP_KNN_MAX_K) ->
 *   dist (B,M,k) squared L2 in the Chamfer op order, idx (B,M,k) int32,
 *   sorted ascending by (distance, index).
 */
#define PP_KNN_MAX_K 64
int pp_knn(const float *query, const float *points, int B, int M, int N, int c, int k,
           float *dist, int32_t *idx, void *workspace, size_t workspace_bytes, int device,
           void *stream);

/*
 * three_nn ("next" row N3).  Replaces sampling.three_nn
 * (_ext/sampling.cpp:163-176,213 -> _ext/interpolate_gpu.cu:9-75).
 *   unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3) squared, idx (B,N,3) int32.
 */
int pp_three_nn(const float *unknown, const float *known, int B, int N, int M, float *dist2,
                int32_t *idx, int device, void *stream);

/*
 * three_interpolate forward / backward ("next" row N3).  Replace sampling.three_interpolate_wrapper /
 * three_interpolate_grad_wrapper (_ext/sampling.cpp:176-203,214-215 -> _ext/interpolate_gpu.cu:77-160).
 *   points (B,C,M), idx (B,N,3), weight (B,N,3) -> out (B,C,N) = sum_k weight[k] * points[idx[k]];
 *   backward ACCUMULATES into grad_points (B,C,M) (caller zero-fills, network/pointnet2_utils.py:81).
 */
int pp_three_interpolate_fwd(const float *points, const int32_t *idx, const float *weight, int B, int C,
                             int M, int N, float *out, int device, void *stream);
int pp_three_interpolate_bwd(const float *grad_out, const int32_t *idx, const float *weight, int B, int C,
                             int N, int M, float *grad_points, int device, void *stream);

/* ------------------------------------------------------- multi-GPU exchange */

/*
 * The one collective of the batch-sharded Chamfer step: the global
 * [sum(dist1), sum(dist2)] over all ranks, exchanged through NVLink peer memory instead of a
 * collective-library call.  No reference counterpart (the reference is single-GPU,
 * SURVEY.md D7); replaces `torch.distributed.all_reduce(sums)` in the sharded step.
 *
 *   create : allocates this rank's mailbox and returns its CUDA IPC handle
 *            (pp_loss_exchange_handle_bytes() bytes) for the host side to distribute;
 *   open   : maps a PEER's mailbox (from its handle) into this process;
 *   bind   : tells the mailbox where every rank's mailbox is mapped (peer_mailboxes[rank] =
 *            this rank's own pointer); world <= 32;
 *   send   : enqueue after pp_chamfer_fwd -- stores `sums` (2 floats, device) into every
 *            peer's mailbox;
 *   wait   : enqueue where the total is needed (after the backward, so the link latency is
 *            hidden) -- polls the own mailbox for this step's contributions and writes their
 *            sum, added in rank order, to sums_out (2 floats, device).  Bounded (option lx_timeout_ms, default 10 min):
 *            after that long without a peer the result is NaN and *status (device int, may be NULL) is 1.
 * Every rank must issue the same sequence of send/wait pairs.  Both launches are graph-capturable.
 */
size_t pp_loss_exchange_handle_bytes(void);
int pp_loss_exchange_create(void **mailbox, unsigned char *handle, int device);
int pp_loss_exchange_open(const unsigned char *handle, void **peer_mailbox, int device);
int pp_loss_exchange_bind(void *mailbox, void *const *peer_mailboxes, int world, int device);
int pp_loss_exchange_close(void *mailbox, void *const *peer_mailboxes, int rank, int world, int device);
int pp_loss_exchange_send(const float *sums, void *mailbox, int rank, int world, int device, void *stream);
int pp_loss_exchange_wait(void *mailbox, int world, float *sums_out, int32_t *status, int device,
                          void *stream);

/* ------------------------------------------------------------- diagnostics */

/*
 * Micro-benchmarks used to measure the roofline denominators that
 * MEASURED_PEAKS.json lacks (FP32 pipe, shared memory, L2).  `which` selects the
 * probe, `iters` its length; returns elapsed milliseconds in *ms and the work
 * (flop or bytes) in *work.  Probes: 0 FFMA, 1 FFMA2, 2 / 3 the Chamfer op mix (scalar / packed),
 * 4 shared memory, 5 L2, 6 REDUX, 7 tensor-memory reads (tcgen05.ld 32x32b.x32 from the sweep kernel's
 * CTA shape, nothing behind them), 8 the same loads followed by the granule minimum tree,
 * 9 / 10 / 11 FMNMX3 / FMNMX / FMNMX3 with rotating sources (work = warp instructions).
 * Not part of the reference surface.
 */
int pp_microbench(int which, int iters, float *ms, double *work, int device);

/*
 * With pp_set_option("timing", 1) every launch of a dominant kernel is bracketed by CUDA
 * events on the launching stream.  pp_timing_collect(name) waits for the recorded events and
 * returns their summed duration and count, then forgets them.  Names: "chamfer_fwd",
 * "chamfer_finalize", "chamfer_bwd", "fps", "ball_query", "query_group", "query_group_bwd", "knn" (the
 * sweep kernel), "gather_fwd".
 */
int pp_timing_collect(const char *name, double *total_ms, int *count);

/*
 * out[0] = a[0] * b[0] + a[1] * b[1] on the device (one thread): the scalar loss from the two fused sums and
 * their weights, so that the autograd node's forward needs no library reduction for two numbers.
 */
int pp_dot2(const float *a, const float *b, float *out, int device, void *stream);

/*
 * Asynchronous copy between host and device memory on `stream` (cudaMemcpyAsync; direction inferred from the
 * pointers).  The host side of the input pipeline (pipeline.HostPrefetcher / HostScalarReader) calls this
 * instead of tensor.copy_ under a stream context: a few microseconds of host time per call instead of 10-20.
 * Host buffers should be pinned, otherwise the call is synchronous.
 */
int pp_memcpy_async(void *dst, const void *src, size_t bytes, int device, void *stream);

/*
 * With pp_set_option("knn_stats", 1) the ordered-sweep KNN path counts the (query block, point
 * tile) pairs it actually evaluated (this synchronises the stream).  Returns the counts of the
 * last such call: visited and total.  The rest were skipped by exact bounding-box pruning.
 */
int pp_knn_stats(double *tiles_visited, double *tiles_total);

/*
 * Process-wide integer options (diagnostics and A/B switches; results never depend on them):
 *   "timing" (0)                 per-kernel CUDA events, see pp_timing_collect
 *   "pdl" (1)                    programmatic dependent launch of the short follow-up kernels
 *   "chamfer_variant" (0)        0 = automatic: the tensor-core sweep (51) from "chamfer_tc_min_pairs"
 *                                (default 2048 * 2048) pairs per cloud on, else the exact FFMA kernel
 *                                (1 above 4096 points, else 32 / 35); 51 = tensor-core sweep: distances
 *                                from tcgen05.mma (3xTF32-split operands, TMEM accumulators), exact
 *                                resolution of the surviving candidates, identical results; 1 / 2 = 256- /
 *                                128-point reference blocks, 128-thread CTAs; 5 = 128-point
 *                                blocks, 64-thread CTAs; 21 / 22 / 25 = 1 / 2 / 5 with the
 *                                query tile staged through shared memory; 31 / 32 / 35 = those
 *                                with the election-free column publish; 13 / 14 = 1 / 2
 *                                without the per-warp sweep rotation
 *   "chamfer_sweep_ctas_per_sm" (8)   CTA target of the tensor-core sweep (reference chunks per query tile)
 *   "chamfer_noelect" (1)        automatic choice below 4097 points uses 32 / 35 (1) or 22 / 25 (0)
 *   "chamfer_blocks_per_sm" (24) target CTA count per SM for the query split heuristic
 *   "chamfer_generic" (0)        force the generic (any point dimension) kernel
 *   "chamfer_ws_check" (0)       debug: verify (synchronously) that a workspace passed with
 *                                PP_CHAMFER_WS_CLEAN really is all-ones; PP_EINVAL if not
 *   "fps_cluster" (0)            0 = automatic, else the cluster width 1 / 2 / 4 / 8
 *   "fps_stream" (0)             force the streaming fallback kernel
 *   "knn_morton" (-1)            -1 = automatic, 0 / 1 = never / always use the ordered sweep
 *   "knn_tc" (-1)                -1 = automatic (from N = 2048: 16 < k <= 32, and k <= 16 on jobs of up to 65536
 *                                queries with clouds of up to 16384 points), 0 / 1 = never / always use the
 *                                tensor-core path (tcgen05 flagging pass + exact resolution; c == 3, k <= 32,
 *                                N <= 262144); timers "knn_sort", "knn_prep", "knn_seed", "knn", "knn_select"
 *   "knn_prune" (1), "knn_estimate" (1), "knn_fused_prep" (1)
 *                                pieces of the ordered sweep: box pruning, threshold seed,
 *                                one-launch preparation
 *   "knn_smem_lists" (0), "knn_generic" (0)   force the k <= 64 / any-dimension kernels
 *   "knn_stats" (0)              see pp_knn_stats
 *   "lx_timeout_ms" (600000)     bound of pp_loss_exchange_wait's poll in wall-clock milliseconds
 *                                (0 = unbounded); on expiry the sums are NaN and *status is set
 */
int pp_set_option(const char *name, int value);

#ifdef __cplusplus
}
#endif
#endif /* PP_B200_H */
