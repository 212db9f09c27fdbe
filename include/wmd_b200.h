/*
 * wmd_b200.h -- C ABI of libwmd_b200.so, the B200 (sm_100a) Word Mover's Distance engine.
 *
 * This is the drop-in boundary for the content-preservation scoring path of
 * iptmt/consistent__style_transfer.  The reference has no FFI of its own on this path (it is
 * Python calling gensim -> pyemd); each entry point below names the reference call it stands
 * behind.  Paths are relative to /root/reference.
 *
 *   reference call                                              entry point here
 *   ---------------------------------------------------------  --------------------------------
 *   WMDdistance.load / init_sims(replace=True)                   wmd_create (+ wmd_normalize_rows)
 *     src/wmd.py:50-55, evaluate/auto/content_preserve.py:38-41
 *   tokenizer.ids_to_tokens + gensim OOV filter                  wmd_set_token_map, wmd_set_rank
 *     src/vocab.py:26-27, src/wmd.py:40
 *   WMDdistance.cal_wmd_label (the per-batch python loop)        wmd_pairs_host / wmd_pairs_dev
 *     src/wmd.py:34-45  (caller: src/loader.py:60)
 *   WMDdistance.cal_wmd -> wv.wmdistance -> pyemd.emd            one pair of the above
 *     src/wmd.py:31-32
 *   calculate_wmd_scores (per-pair loop)                         wmd_pairs_host
 *     evaluate/auto/content_preserve.py:43-50 (caller: evaluate/eval.py:42)
 *   gensim nbow() / Dictionary.doc2bow  [third party]            wmd_nbow_host
 *   pyemd.emd(p, q, distance_matrix) called directly             wmd_emd_batch_host
 *     evaluate/auto/transfer_intensity.py:8-11 (calculate_emd)
 *   (not in the reference: Kusner et al. 2015 lower bound)       wmd_rwmd_pairs_host
 *   (not in the reference: all-pairs top-k with RWMD pruning,    wmd_allpairs_topk_host
 *    BASELINE.json configs[3])
 *
 * Conventions
 *   - every function returns 0 on success and a negative WMD_E* code on failure; the message
 *     is available from wmd_last_error() (thread-local).  Nothing throws.
 *   - "documents" are CSR-packed int32 id lists: ids[off[p] .. off[p+1]) is document p,
 *     off has npairs+1 int64 entries.  ids are embedding-table rows, or tokenizer ids when a
 *     token map is installed; anything that does not map to a row in [0, V) is out of
 *     vocabulary and is dropped (gensim's `token in self.vocab` filter).
 *   - *_host entries take HOST pointers (pinned or pageable) and do their own chunked
 *     host<->device copies; *_dev entries take DEVICE pointers and are ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream) with no host sync.
 *   - the caller owns every buffer; the handle owns its device copy of the table, the maps
 *     and a grow-only workspace.  A handle is bound to one device and is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device wmd_create fails with WMD_ENODEV.
 *
 * Per-pair status codes (gensim's early-outs):
 *   0 = ok, finite distance                      1 = +inf, a document is empty after OOV removal
 *   2 = 0.0, the union vocabulary has one token   3 = +inf, all-zero distance matrix
 */
#ifndef WMD_B200_H
#define WMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wmd_engine *wmd_handle;

#define WMD_OK        0
#define WMD_EINVAL   -1   /* bad argument (null pointer, negative size, document longer than WMD_MAX_DOC_LEN, ...) */
#define WMD_ENODEV   -2   /* no usable CUDA device */
#define WMD_ECUDA    -3   /* a CUDA runtime call failed; see wmd_last_error() */
#define WMD_ENOMEM   -4

#define WMD_MAX_DOC_LEN 256      /* tokens per document (before OOV removal) */

/* distance definition */
#define WMD_MODE_PYEMD 0   /* pyemd emd_hat_gd_metric: 1e6-grid integer optimum (bit-faithful to the reference) */
#define WMD_MODE_EXACT 1   /* additive: the real-valued transportation optimum in FP64 (no grid, no cancellation);
                              differs from the reference's value by ~1e-6 relative */
/* flag, OR-ed into `mode`: the ids of THIS call are embedding-table rows even though a token map is installed
 * (gensim's wv.wmdistance(tokens) and cal_wmd_label(tokenizer ids) served by one handle, src/wmd.py:31-45) */
#define WMD_IDS_ARE_ROWS 0x100

/* lifetime ------------------------------------------------------------------------------------ */

/* Copies the [V, d] float32 table (HOST pointer, row_stride floats between rows) to `device`.
 * normalize != 0 applies gensim init_sims(replace=True) on the device: each row divided by
 * sqrt(sum(row**2)) in float32 with numpy's summation order.  */
int wmd_create(const float *table_host, int64_t V, int32_t d, int64_t row_stride,
               int32_t normalize, int32_t device, wmd_handle *out);
int wmd_destroy(wmd_handle h);
const char *wmd_last_error(void);
const char *wmd_version(void);

/* tokenizer id -> table row (-1 = OOV), n entries; NULL/0 removes the map (ids are rows). */
int wmd_set_token_map(wmd_handle h, const int32_t *id_to_row_host, int64_t n);
/* rank[row] = position of the row's token in gensim Dictionary (python string sort) order.
 * Fixes the canonical order of nBOW outputs and of the FP64 mass summation. NULL = row order. */
int wmd_set_rank(wmd_handle h, const int32_t *rank_host, int64_t V);
/* copy the (possibly normalised) device table back to the host, [V, d] dense */
int wmd_get_table(wmd_handle h, float *out_host);

/* scoring ------------------------------------------------------------------------------------- */

/* WMD of npairs document pairs. out[npairs] float64, status[npairs] int32 (may be NULL). */
int wmd_pairs_host(wmd_handle h, const int32_t *ids1, const int64_t *off1,
                   const int32_t *ids2, const int64_t *off2, int64_t npairs,
                   int32_t mode, double *out, int32_t *status);

/* wmd_pairs_host whose results stay on the device: documents in HOST memory (chunked copies overlapped with the
 * kernels, as wmd_pairs_host), out_dev[npairs] / status_dev[npairs] (may be NULL) DEVICE arrays, complete on return.
 * The multi-GPU layer scores a rank's slice this way and hands the device array to the NCCL gather. */
int wmd_pairs_host_in_dev_out(wmd_handle h, const int32_t *ids1, const int64_t *off1,
                              const int32_t *ids2, const int64_t *off2, int64_t npairs,
                              int32_t mode, double *out_dev, int32_t *status_dev);

/* Asynchronous host entry for batches of the in-loop caller's size (src/loader.py:60 -> src/wmd.py:34-45): the
 * label of batch k+1 is computed while training step k runs.  wmd_pairs_submit copies the documents (pageable or
 * pinned HOST memory; the caller's buffers are free again on return) into pinned staging the handle owns, queues
 * copies and kernels on the handle's streams and returns; wmd_pairs_wait blocks until the job is done and copies
 * out[npairs] / status[npairs] (may be NULL) to the caller.  One job per handle may be in flight; every other
 * scoring entry fails with WMD_EINVAL until it has been waited for. */
int wmd_pairs_submit(wmd_handle h, const int32_t *ids1, const int64_t *off1,
                     const int32_t *ids2, const int64_t *off2, int64_t npairs, int32_t mode);
int wmd_pairs_wait(wmd_handle h, double *out, int32_t *status);

/* Same on device buffers, stream-ordered.  max_len1/max_len2 bound the document lengths of
 * each side (needed to size the launch without a host sync); total1/total2 are off1[npairs],
 * off2[npairs].  */
int wmd_pairs_dev(wmd_handle h, const int32_t *ids1, const int64_t *off1, int64_t total1, int32_t max_len1,
                  const int32_t *ids2, const int64_t *off2, int64_t total2, int32_t max_len2,
                  int64_t npairs, int32_t mode, double *out, int32_t *status, void *stream);

/* Padded [npairs, L1] / [npairs, L2] int32 id matrices on the device (pad_id entries are
 * skipped wherever they occur), stream-ordered, no host sync.  For a validation hook on
 * src/main_optimize.py:127-141 (tokens = sample_p.argmax(-1), x padded with PAD_ID=0). */
int wmd_pairs_padded_dev(wmd_handle h, const int32_t *a, int32_t L1, const int32_t *b, int32_t L2,
                         int64_t npairs, int32_t pad_id, int32_t mode,
                         double *out, int32_t *status, void *stream);

/* nBOW of ndocs documents: unique in-vocabulary rows in canonical order, int32 counts and
 * FP64 weights count/len, written at the document's own CSR offset (first u entries of its
 * slot); uniq[ndocs] receives u. */
int wmd_nbow_host(wmd_handle h, const int32_t *ids, const int64_t *off, int64_t ndocs,
                  int32_t *rows, int32_t *counts, double *weights, int32_t *uniq);

/* wmd_nbow_host on device buffers: ids / off / outputs are DEVICE pointers, max_len bounds the
 * document length, the launch is ordered on `stream`, no host sync.  weights may be NULL. */
int wmd_nbow_dev(wmd_handle h, const int32_t *ids, const int64_t *off, int64_t ndocs, int32_t max_len,
                 int32_t *rows, int32_t *counts, double *weights, int32_t *uniq, void *stream);

/* Relaxed WMD lower bound per pair: lb = max(l1, l2), l1 = sum_i w1[i] min_j c[i][j] (FP64,
 * canonical order), with the argmin column of every doc1 row / argmin row of every doc2
 * column (lowest index on ties) written at the documents' CSR offsets.  l1/l2/argmins may be NULL. */
int wmd_rwmd_pairs_host(wmd_handle h, const int32_t *ids1, const int64_t *off1,
                        const int32_t *ids2, const int64_t *off2, int64_t npairs,
                        double *lb, double *l1, double *l2,
                        int32_t *argmin_rows, int32_t *argmin_cols, int32_t *status);

/* wmd_rwmd_pairs_host on device buffers, stream-ordered, no host sync; arguments as wmd_pairs_dev.  The argmin
 * arrays are indexed by the documents' CSR offsets (total1 / total2 entries). */
int wmd_rwmd_pairs_dev(wmd_handle h, const int32_t *ids1, const int64_t *off1, int64_t total1, int32_t max_len1,
                       const int32_t *ids2, const int64_t *off2, int64_t total2, int32_t max_len2, int64_t npairs,
                       double *lb, double *l1, double *l2, int32_t *argmin_rows, int32_t *argmin_cols,
                       int32_t *status, void *stream);

/* pyemd.emd(first_histogram, second_histogram, distance_matrix, extra_mass_penalty) for nprob
 * independent problems of n <= 31 bins each: P, Q are [nprob, n] float64 HOST arrays, D is one
 * [n, n] matrix shared by every problem (shared_d != 0) or [nprob, n, n].  Same arithmetic as the
 * WMD path (emd_hat_gd_metric<double>: bin-wise cancellation, 1e6-grid quantisation, exact integer
 * optimum, un-normalisation); extra_mass_penalty = -1 means max(D), as in pyemd.  Preconditions as
 * upstream: non-negative entries, a positive total mass and a positive max(D). */
int wmd_emd_batch_host(wmd_handle h, const double *P, const double *Q, const double *D, int64_t nprob,
                       int32_t n, int32_t shared_d, double extra_mass_penalty, double *out);

/* All-pairs mode (BASELINE.json configs[3]; not in the reference).  For every document i of set A
 * in [row_begin, row_end), the k documents j of set B with the smallest WMD(i, j), ordered by
 * (distance, j); the distances are the same values wmd_pairs_host returns for (A_i, B_j), +inf
 * included.  Candidates are pruned with the relaxed lower bound (Kusner et al. 2015) evaluated
 * through a V x V table of word distances and per-corpus word-to-document minima (LC-RWMD, Atasu
 * et al. 2017); every surviving pair is solved exactly.  out_idx / out_dist are
 * [(row_end - row_begin), k] HOST arrays.  Row blocks are independent, so ranks of a multi-GPU job
 * call this with disjoint [row_begin, row_end).
 * stats[8] (may be NULL): {bounds evaluated, exact solves round 1, exact solves round 2, query
 * blocks, 0...};  ms[4] (may be NULL): device milliseconds spent on {word-distance table (first
 * call only), corpus index Z_B + nBOW, bounds + selection, exact solves + merges}. */
int wmd_allpairs_topk_host(wmd_handle h, const int32_t *idsA, const int64_t *offA, int64_t nA,
                           const int32_t *idsB, const int64_t *offB, int64_t nB, int32_t k,
                           int64_t row_begin, int64_t row_end, int32_t *out_idx, double *out_dist,
                           int64_t *stats, double *ms);

/* wmd_allpairs_topk_host whose result stays on the device: the documents are HOST arrays as above
 * (they are staged once), out_idx_dev / out_dist_dev are DEVICE arrays of [(row_end - row_begin), k] entries, complete
 * on return -- a multi-GPU job hands them straight to the NCCL all-gather of the per-row results. */
int wmd_allpairs_topk_dev(wmd_handle h, const int32_t *idsA, const int64_t *offA, int64_t nA,
                          const int32_t *idsB, const int64_t *offB, int64_t nB, int32_t k,
                          int64_t row_begin, int64_t row_end, int32_t *out_idx_dev, double *out_dist_dev,
                          int64_t *stats, double *ms);

/* Word-distance table of the pair entries (the reference has no counterpart: gensim's wmdistance,
 * models/keyedvectors.py [gensim 3.8], recomputes every word distance on every call).  A handle keeps the float32
 * distance of every two rows of its embedding table -- V * V * 4 bytes of device memory, computed ONCE by the pair
 * path's own cost kernels, so each entry is bit-identical to what they produce for a pair -- and every pair call
 * takes its costs from it: one fused kernel per chunk does nBOW, cost gather and the exact solve of a pair in one
 * warp (csrc/fused.cuh), nothing in between touches device memory.  Results do not change.
 * Default policy: ON when the table fits the budget (4 GiB and at most a quarter of the free device memory;
 * environment WMD_DTAB_BUDGET_MB, WMD_DTAB=0/1), built by the first scoring call of the handle (that call is
 * host-synchronous once, also through a *_dev entry).  enabled != 0 forces it on and builds it now (WMD_ENOMEM
 * when it does not fit); enabled == 0 returns to the direct path that recomputes every distance from the embedding
 * rows (the table, if built, is kept).  The all-pairs entry builds and uses the same table on its own. */
int wmd_set_distance_table(wmd_handle h, int32_t enabled);
/* bytes the table takes (V * V * 4), wall milliseconds its one-off build took (0 until built), whether the pair
 * entries use it, whether it is resident.  Any pointer may be NULL. */
int wmd_distance_table_info(wmd_handle h, int64_t *bytes, double *build_ms, int32_t *enabled, int32_t *resident);

/* Device memory of a handle.  *estimate: upper bound of the bytes a pair call of npairs documents
 * of at most max_len1 / max_len2 tokens needs in total (embedding table, maps, word-distance table under the current
 * policy, both workspace slots); *resident: what the handle holds right now.  Either pointer may be NULL. */
int wmd_workspace_bytes(wmd_handle h, int64_t npairs, int32_t max_len1, int32_t max_len2,
                        int64_t *estimate, int64_t *resident);

/* multi-GPU: the score gather fused into the kernels -------------------------------------------
 * Not in the reference (single GPU, /root/reference/job.yaml:29-32).  SURVEY.md 8(e): pairs are independent, the only
 * exchange of a sharded job is the gather of the final scores; instead of an all-gather after the kernels, every
 * rank's kernels store each score / status straight into the peers' copies of the global result over NVLink
 * (fire-and-forget stores that overlap the solves), and the ranks only need a barrier before reading. */

#define WMD_IPC_HANDLE_BYTES 64
#define WMD_MAX_FANOUT 7

/* A device buffer other processes of the same box can map: cudaMalloc + zero fill, its CUDA IPC handle in
 * ipc_handle[WMD_IPC_HANDLE_BYTES] (hand it to the peers by any means, e.g. torch.distributed.all_gather_object). */
int wmd_peer_alloc(wmd_handle h, int64_t bytes, void **dev_ptr, unsigned char *ipc_handle);
/* Maps a peer's buffer into this process (peer access is enabled on first use); *dev_ptr is valid on h's device. */
int wmd_peer_open(wmd_handle h, const unsigned char *ipc_handle, void **dev_ptr);
/* opened != 0: unmaps a buffer of wmd_peer_open; 0: frees a buffer of wmd_peer_alloc. */
int wmd_peer_close(wmd_handle h, void *dev_ptr, int32_t opened);
/* From now on every score / status the pair entries (wmd_pairs_dev, wmd_pairs_padded_dev, wmd_pairs_host,
 * wmd_pairs_host_in_dev_out; WMD_MODE_PYEMD) store at pair index p ALSO goes to out_ptrs[k][p] / status_ptrs[k][p],
 * k < n <= WMD_MAX_FANOUT (device pointers valid on h's device, e.g. from wmd_peer_open, already advanced to this rank's
 * first pair).  n == 0 switches it off.  The calls stay stream-ordered; a cross-rank barrier after them makes the peers'
 * copies complete. */
int wmd_set_fanout(wmd_handle h, int32_t n, double *const *out_ptrs, int32_t *const *status_ptrs);

/* instrumentation ----------------------------------------------------------------------------- */

/* When enabled, every kernel launch is bracketed by CUDA events on its own stream. */
int wmd_set_profiling(wmd_handle h, int32_t enabled);
/* When enabled, all chunks of a call run on ONE internal stream (no overlap between the solver of one
 * chunk and the cost kernels of the next): slower, but the per-kernel event times are the kernels' own. */
int wmd_set_serial(wmd_handle h, int32_t enabled);
#define WMD_K_NBOW   0
#define WMD_K_COST   1
#define WMD_K_SOLVE  2
#define WMD_K_RWMD   3
#define WMD_K_MISC   4
#define WMD_K_FUSED  5   /* table mode: nBOW + cost gather + solve of one pair in one warp (fused.cuh) */
#define WMD_K_COUNT  6
/* accumulated since the last reset: ms[k] device milliseconds, launches[k] launch count */
int wmd_get_profile(wmd_handle h, double *ms, int64_t *launches, int32_t reset);
/* totals of the last wmd_pairs_* call: sum over pairs of tokens, unique rows and tile cells
 * (for the algorithmic-bytes figure of the roofline). values[0..5] =
 * {tokens1+tokens2, uniq1+uniq2, cells, solved_pairs, max_rows, max_cols} */
int wmd_get_last_stats(wmd_handle h, int64_t *values);

#ifdef __cplusplus
}
#endif
#endif /* WMD_B200_H */
