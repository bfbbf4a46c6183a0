/* efmtl.h -- C ABI of libefmtl.so: the B200 (sm_100a) embedding-metric hot path.
 *
 * Drop-in boundary for the triplet-loss head, the EFM/MFM max-out activation and the cosine
 * verification / gallery scoring of joannhsiao/Improving_Face_recognition_Performance_using_Triplet_Loss.
 * Each entry point cites the reference interface it replaces (paths relative to the reference root;
 * "Feature.hpp" = feature_extraction/c_version/Feature.hpp).
 *
 * Conventions (all entry points)
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise; the library
 *     never allocates, frees or synchronises (one exception: the IPC exchange regions of the multi-GPU
 *     step, efm_xchg_alloc / efm_xchg_free); work is enqueued on `stream` (a cudaStream_t passed
 *     as void*, NULL = legacy default stream);
 *   - return 0 on success, <0 on error (EFM_E_*); efm_last_error() returns a thread-local message
 *     (mirrors the reference's return -1 / false convention, Feature.hpp:215-217,314-317);
 *   - row-major, innermost dimension contiguous; dtype codes below; labels are int32;
 *   - re-entrant; the only shared state is a mutex-protected cache of the TMA encode entry point.
 */
#ifndef EFMTL_H_
#define EFMTL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFM_ABI_VERSION 2   /* 2: multi-GPU fused step (efm_bh_dist_fwd_bwd, efm_xchg_*, efm_ipc_*), efm_bh_profile_events */

/* dtype codes */
#define EFM_F32  0
#define EFM_BF16 1
#define EFM_F16  2

/* error codes */
#define EFM_OK            0
#define EFM_E_ARG        -1   /* bad shape / null pointer / unsupported dtype */
#define EFM_E_ALIGN      -2   /* pointer or stride alignment the kernel needs is missing */
#define EFM_E_WORKSPACE  -3   /* workspace too small (see *_workspace_bytes) */
#define EFM_E_CUDA       -4   /* a CUDA runtime/driver call failed (message has the detail) */
#define EFM_E_UNSUPPORTED -5  /* valid request this build does not implement (e.g. D too large) */

/* activation modes (efm_act_*) */
#define EFM_ACT_EFM3 3   /* concat(max3, min3): C -> 2C/3  (lightcnn.py:32-37)           */
#define EFM_ACT_MFM2 2   /* max(s0, s1):        C -> C/2   (efm_symbol.py:63-64,76-77)   */

/* batch-hard flags */
#define EFM_BH_NORMALIZE       1  /* per-row L2 normalise before mining; gradient chains through    */
#define EFM_BH_DETACH_GATHERED 2  /* rows outside [row_offset,row_offset+B_local) get no gradient
                                     (mirrors the detached negatives, pre-trained_efm_v3.py:206-207) */
#define EFM_BH_MINE_BF16       4  /* fp32 input: run the mining Gram on bf16-rounded rows through the
                                     tcgen05 path (loss/gradient stay fp32).  Implied for bf16 input. */

/* gallery flags */
#define EFM_GAL_POSITIVE_ONLY  1  /* keep only sim > 0: `maxsim` starts at 0, Feature.hpp:347,371 */

const char* efm_last_error(void);
int efm_abi_version(void);
/* compute capability of the current device as major*10+minor (100 on B200), <0 on error */
int efm_device_arch(void);

/* ---- a6: EFM-3 / MFM-2 max-out activation --------------------------------------------------
 * Replaces SliceChannel(num_outputs=3|2, axis=1) + maximum/minimum + Concat
 * (lightcnn.py:22-27,32-37,53-58,61-66,123-128; efm_symbol.py:63-77; mutli_gpu_v3.py:20-32).
 * x: (N, C, HW) NCHW flattened; y: (N, 2C/3, HW) for EFM3, (N, C/2, HW) for MFM2.
 * NaN and tie rules follow MXNet's maximum/minimum (NaN propagates; ties -> lowest slice). */
int efm_act_fwd(int mode, const void* x, void* y, long N, long C, long HW, int dtype, void* stream);
int efm_act_bwd(int mode, const void* x, const void* dy, void* dx, long N, long C, long HW, int dtype,
                void* stream);

/* ---- a3: normalisation ------------------------------------------------------------------------
 * Per-row  fc[i] / mx.nd.norm(fc[i])  (final_efm.py:240-243, extract_feacture_v2.py:70,97) and its
 * backward; whole-block  anc / mx.nd.norm(anc)  (train_efm.py:241).  No epsilon (zero row -> NaN). */
int efm_l2norm_fwd(const void* x, void* y, float* row_norm /*B, nullable*/, long B, long D, int dtype,
                   void* stream);
int efm_l2norm_bwd(const void* x, const void* dy, void* dx, long B, long D, int dtype, void* stream);
/* sum of squares of all n elements -> *out (device float, must be zeroed by the caller) */
int efm_sumsq(const void* x, long n, int dtype, float* out, void* stream);
/* y = x / sqrt(*sumsq)   (mx.nd.norm(x) is sqrt of the sum over ALL elements) */
int efm_frob_fwd(const void* x, const float* sumsq, void* y, long n, int dtype, void* stream);
/* dx = (dy - y * dot) / sqrt(*sumsq), y = x/|x|_F, *dot = sum(y*dy) (device float, see efm_dot) */
int efm_frob_bwd(const void* x, const void* dy, const float* sumsq, const float* dot_x_dy, void* dx,
                 long n, int dtype, void* stream);
/* sum(x*y) over n elements -> *out (device float, zeroed by the caller) */
int efm_dot(const void* x, const void* y, long n, int dtype, float* out, void* stream);

/* ---- a1: gluon.loss.TripletLoss(margin)(pred, positive, negative) --------------------------------
 * (pre-trained_efm_v3.py:183,210; train_efm.py:210,241; final_efm.py:220,259)
 * loss_i = relu(sum_d[(p-a)^2 - (n-a)^2] + margin), (B,) fp32, no mean. */
int efm_triplet_fwd(const void* a, const void* p, const void* n, long B, long D, int dtype, float margin,
                    float* loss, void* stream);
/* gradients of sum_i head_grad_i*loss_i (head_grad NULL = ones, i.e. loss.backward());
 * da/dp/dn are (B,D) in `dtype`, any of them may be NULL (tensor not in the autograd graph). */
int efm_triplet_bwd(const void* a, const void* p, const void* n, const float* head_grad, long B, long D,
                    int dtype, float margin, void* da, void* dp, void* dn, void* stream);

/* ---- a4: cosine_dist (train_efm.py:26-34): out_i = dot(a_i,b_i)/(|a_i|*|b_i|), (P,) fp32 -------- */
int efm_cosine_pairs(const void* a, const void* b, long P, long D, int dtype, float* out, void* stream);
/* gathered form for LFW-style pair lists: rows ia[k], ib[k] of one embedding matrix */
int efm_cosine_pairs_indexed(const void* E, const int32_t* ia, const int32_t* ib, long P, long D, int dtype,
                             float* out, void* stream);
/* all-pairs cosine matrix out[i*ld_out + j] = cos(A_i, B_j), fp32, ld_out >= N (a multiple of 32 keeps the
 * stores sector-aligned).  Operands are rounded to bf16 and run on tcgen05 (fp32 accumulate); needs workspace
 * for padded operands + norms. */
size_t efm_cosine_allpairs_workspace_bytes(long M, long N, long D, int dtype);
int efm_cosine_allpairs(const void* A, const void* B, long M, long N, long D, int dtype, float* out, long ld_out,
                        void* workspace, size_t ws_bytes, void* stream);

/* ---- a2': fused batch-hard triplet loss fwd+bwd (north_star; extends a1/a2) --------------------------
 * E_all: (B_all, D) embeddings (the all-gathered global batch; B_all == B_local on one GPU).
 * Anchors are rows [row_offset, row_offset+B_local).  For anchor i:
 *   p* = argmax_{j != i, y_j == y_i} d_ij,  n* = argmin_{y_j != y_i} d_ij,  d = squared Euclidean,
 *   lowest index on ties, -1 when no candidate; loss_i = relu(d_ip* - d_in* + margin) (0 if either
 *   index is -1).  dE_all (B_all, D) fp32 receives d(sum_i head_grad_i*loss_i)/dE_all (NULL = forward
 *   only).  The B x B distance matrix is never written to memory. */
size_t efm_bh_workspace_bytes(long B_local, long B_all, long D, int dtype, int flags);
int efm_bh_triplet_fwd_bwd(const void* E_all, const int32_t* labels_all, long B_all, long row_offset,
                           long B_local, long D, int dtype, float margin, const float* head_grad,
                           int flags, float* loss, int32_t* idx_p, int32_t* idx_n, float* dE_all,
                           void* workspace, size_t ws_bytes, void* stream);

/* The mining kernel alone, on a workspace already prepared by efm_bh_triplet_fwd_bwd with the same shapes
 * (re-mining is idempotent).  Exists so the dominant kernel can be timed / profiled in isolation. */
int efm_bh_mine(const void* E_all, long B_all, long row_offset, long B_local, long D, int dtype, int flags,
                void* workspace, size_t ws_bytes, void* stream);

/* Backward only, for given mined indices and a per-anchor head gradient (autograd over the un-reduced
 * loss vector): dE_all = d(sum_i head_grad_i * loss_i)/dE_all.  dE_all is overwritten. */
size_t efm_bh_backward_workspace_bytes(long B_all);
int efm_bh_backward(const void* E_all, long B_all, long row_offset, long B_local, long D, int dtype,
                    float margin, const float* head_grad, int flags, const int32_t* idx_p,
                    const int32_t* idx_n, float* dE_all, void* workspace, size_t ws_bytes, void* stream);

/* Multi-GPU exact gradient without a dense reduce-scatter: given the all-gathered batch and the all-gathered
 * (idx_p, idx_n, wgt) of EVERY anchor (anchor j = row j; wgt_j = 2 * head_grad_j if loss_j > 0 else 0), compute
 * dE for rows [row_offset, row_offset+B_local) only -- anchor-role terms of the local anchors plus the positive /
 * negative-role terms of all anchors (of any rank) that mined one of these rows.  dE_local: (B_local, D) fp32. */
size_t efm_bh_grad_rows_workspace_bytes(long B_all, long B_local);
int efm_bh_grad_rows(const void* E_all, long B_all, long D, int dtype, int flags, const int32_t* idx_p_all,
                     const int32_t* idx_n_all, const float* wgt_all, long row_offset, long B_local,
                     float* dE_local, void* workspace, size_t ws_bytes, void* stream);

/* ---- multi-GPU global batch-hard step with the exchange fused into the Gram kernel (BASELINE config C3) ----
 * Replaces the data-parallel split of mutli_gpu_v3.py:153 for the loss head: one process per GPU, rank r owns B_local
 * rows; the global batch is world * B_local rows, rank r's rows are global rows [r*B_local, (r+1)*B_local).
 *
 * Exchange regions: every rank allocates one with efm_xchg_alloc(efm_bh_xchg_bytes(...)) -- plain cudaMalloc memory,
 * zero-filled; the ONE place where the library allocates, because CUDA IPC cannot export sub-allocations of a caching
 * allocator -- exports it with efm_ipc_export (64-byte handle, sent to the peers by any host channel) and maps every
 * peer's region with efm_ipc_import.  `regions[r]` (HOST array of `world` device pointers) is rank r's region as
 * mapped in this process, regions[rank] the own allocation.  All ranks must have mapped all regions before the first
 * step and must call efm_bh_dist_fwd_bwd with the same shapes, flags and `step` (1, 2, 3, ... -- it selects one of two
 * buffers and the value the arrival counters are compared with; nothing is reset between steps).
 *
 * The step enqueues prep (local rows) -> Gram + mining kernel whose extra warps COPY the peers' rows into the own
 * region with loads over NVLink (block by block, in the order the tile schedule needs them) and build the label table
 * of the global batch (see csrc/bh.cu) -> finalize -> (exact mode) chain kernel that fetches 12 bytes per anchor from
 * the owners over NVLink -> gradient of the local rows.  No NCCL call, no
 * host synchronisation.  Results equal efm_bh_triplet_fwd_bwd on the concatenated batch: loss / idx_p / idx_n
 * (B_local; indices are global rows), dE_local (B_local, D) fp32 or NULL (forward only).  EFM_BH_DETACH_GATHERED:
 * gradient only through roles played by local rows of local anchors (no result exchange).  Mining always runs on the
 * tcgen05 path (bf16-rounded rows).  Requires B_local % 256 == 0, D % 8 == 0, D <= 512, world <= 8.
 * A peer that never arrives makes the waiting kernel trap after 8 s (launch failure) instead of hanging. */
size_t efm_bh_xchg_bytes(long B_local, int world, long D, int dtype);
size_t efm_bh_dist_workspace_bytes(long B_local, int world);
int efm_xchg_alloc(size_t bytes, void** region);
int efm_xchg_free(void* region);
int efm_ipc_export(const void* region, void* handle64 /* 64 bytes, host */);
int efm_ipc_import(const void* handle64, void** mapped);
int efm_ipc_release(void* mapped);
int efm_bh_dist_fwd_bwd(const void* E_local, const int32_t* labels_local, long B_local, long D, int dtype,
                        float margin, const float* head_grad, int flags, int world, int rank,
                        void* const* regions, unsigned long long step, float* loss, int32_t* idx_p,
                        int32_t* idx_n, float* dE_local, void* workspace, size_t ws_bytes, void* stream);
/* Profiling aid: the following efm_bh_triplet_fwd_bwd / efm_bh_dist_fwd_bwd calls of this thread record
 * events[i] (cudaEvent_t passed as void*, NULL entries skipped) on the call's stream at stage boundary i:
 * 0 step start, 1 after memset + prep, 2 after the Gram + mining (+ exchange) kernel, 3 unused, 4 after finalize, 5 after the chain kernel (multi-GPU exact gradient only), 6 after
 * the gradient kernel.  n = 0 switches it off. */
#define EFM_BH_PROFILE_EVENTS 7
int efm_bh_profile_events(void* const* events, int n);

/* ---- a8-a11: gallery scoring ---------------------------------------------------------------------
 * simd_dot (Feature.hpp:273-293) + Compare_Face_From_DB (:345-392 argmax, :295-343 top-K) +
 * Compare_Face_Person/Compare_Face_DB (:763-804).  G: (N,D), gnorm: (N,) stored row norms
 * (`sqrt_simdot`), Q: (nq,D) queries.
 *
 * efm_gallery_scores_f32: out[q][r] = simd_dot(G_r,Q_q)/(gnorm_r*sqrt(simd_dot(Q_q,Q_q))) with the
 * reference's exact fp32 operation order (bit-identical to the SSE code). */
int efm_gallery_scores_f32(const float* G, const float* gnorm, long N, long D, const float* Q, long nq,
                           float* out, void* stream);
/* out[q][r] = simd_dot(G_r, Q_q) alone (Feature.hpp:273-293), same exact operation order */
int efm_simd_dot_rows(const float* G, long N, long D, const float* Q, long nq, float* out, void* stream);
/* out[r] = sqrt(simd_dot(G_r, G_r)): the stored `sqrt_simdot` of a registered row (Feature.hpp:476,535) */
int efm_simd_norm_rows(const float* G, long N, long D, float* out, void* stream);
/* Top-K search.  dtype F32: exact-order fp32 scores (as above); dtype BF16: tcgen05 Gram with fp32
 * accumulate.  NaN scores skipped, score < sim_th skipped, (score desc, index asc) order, strict
 * replacement (earlier row survives ties).  out_score/out_idx: (nq,K); unused slots idx -1, score 0.
 * idx_base is added to every index (gallery shard offset). K <= EFM_GAL_MAX_K. */
#define EFM_GAL_MAX_K 16
size_t efm_gallery_topk_workspace_bytes(long N, long D, long nq, int K, int dtype);
int efm_gallery_topk(const void* G, const float* gnorm, long N, long D, const void* Q, long nq, int K,
                     float sim_th, int flags, int dtype, long idx_base, float* out_score, int32_t* out_idx,
                     void* workspace, size_t ws_bytes, void* stream);
/* K-way merge of `parts` partial lists per query (layout (parts, nq, K)) -> (nq, K) */
int efm_topk_merge(const float* part_score, const int32_t* part_idx, int parts, long nq, int K,
                   float* out_score, int32_t* out_idx, void* stream);
/* per-identity reduction for Compare_Face_DB: best row index per query -> its person id (host side
 * is a gather; provided for completeness): out_pid[q] = idx[q] < 0 ? -1 : person_ids[idx[q]] */
int efm_gather_pid(const int32_t* idx, const int32_t* person_ids, long nq, int32_t* out_pid, void* stream);

/* ---- a5: cosine histogram (draw_cos_dis_real.py:4-7): np.histogram(x, bins) counts ---------------
 * edges: (bins+1,) float64 device array (np.linspace(min, max, bins+1)); the kernel reproduces numpy's
 * uniform-bin index rule incl. its edge corrections; counts: (bins,) zeroed by the caller. */
int efm_histogram(const float* x, long n, const double* edges, int bins, unsigned long long* counts,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EFMTL_H_ */
