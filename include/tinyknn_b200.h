/*
 * tinyknn_b200.h -- C ABI of the B200-native tinyknn query hot path.
 *
 * This is the drop-in boundary: plain C types, raw pointers + extents, no torch / numpy types.
 * Every entry point returns an int status (TKB_OK == 0) and never throws; tkb_last_error() gives
 * the message of the last failure on the calling thread. Paths cited as "ref:" are relative to
 * the upstream repository thomasahle/tinyknn.
 *
 * Two families:
 *   *_host : the exact surface of the reference's Cython kernel modules (tinyknn._fast_pq,
 *            tinyknn._fast_pq_avx). Pointers are HOST memory owned by the caller (numpy buffers);
 *            the call copies in, runs the CUDA kernels on the current device, copies out and
 *            returns when the result is in the caller's buffer. This is what a reference
 *            maintainer binds.
 *   *_dev  : the same operations (and their batched forms) on DEVICE pointers and a CUDA stream
 *            (cudaStream_t passed as void*). Asynchronous; used by the tinyknn_b200 host layer to
 *            keep indexes resident in HBM.
 *
 * Data layouts (ref: tinyknn/_transform.py:53-77, :114-138; pinned by tests/test_transform.py:80-101)
 *   codes  uint64[n_chunks][M] : chunk = 16 vectors; the 16 bytes at &codes[c][2p] hold one byte per
 *          vector of the chunk, low nibble = code of sub-quantizer 2p, high nibble = 2p+1.
 *   tables uint64[2M] == uint8[M][16] : row j = LUT of sub-quantizer j (int8 when signd).
 *   est    one byte per (padded) vector position, int8 when signd, else uint8.
 *   heap   indices int64[R] + vals int32[R], array-layout binary max-heap (root = largest value).
 */
#ifndef TINYKNN_B200_H
#define TINYKNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TKB_API __attribute__((visibility("default")))
#else
#define TKB_API
#endif

#define TKB_OK            0
#define TKB_ERR_INVALID   1   /* bad argument (null pointer, M not a multiple of 2/4, ...) */
#define TKB_ERR_CUDA      2   /* CUDA runtime/driver error, see tkb_last_error() */
#define TKB_ERR_NO_DEVICE 3   /* no usable CUDA device: there is NO CPU fallback */

#define TKB_ORDER_SSE 0       /* ref: tinyknn/_fast_pq.pyx:209-236  (one accumulator)            */
#define TKB_ORDER_AVX 1       /* ref: tinyknn/_fast_pq_256.pyx:126-156 (two lanes, default)      */

#define TKB_DTYPE_F32 0
#define TKB_DTYPE_F64 1

TKB_API int         tkb_version(void);
TKB_API const char *tkb_last_error(void);
TKB_API int         tkb_device_count(int *count);
/* number of CUDA kernels this library has launched so far in this process (all threads) */
TKB_API long long   tkb_launch_count(void);

/* ------------------------------------------------------------------------------------------ */
/* Host-buffer surface == the reference's Cython modules                                       */
/* ------------------------------------------------------------------------------------------ */

/* replaces estimate_pq_sse (ref: tinyknn/_fast_pq.pyx:101-111) and
 *          estimate_pq_avx (ref: tinyknn/_fast_pq_256.pyx:52-62).
 * data: uint64[n_chunks][M]; tables: uint64[2M]; out: uint64[2*n_chunks]. */
TKB_API int tkb_estimate_pq_host(const uint64_t *data, int64_t n_chunks, int M, const uint64_t *tables,
                         uint64_t *out, int order, int signd);

/* replaces query_pq_sse (ref: tinyknn/_fast_pq.pyx:114-206) and
 *          query_pq_avx (ref: tinyknn/_fast_pq_256.pyx:65-123).
 * indices/vals: caller-owned heap of R slots, updated in place (state persists across calls,
 * ref: tinyknn/ivf.py:137-150). labels: int64[>= n] or NULL (label = position). */
TKB_API int tkb_query_pq_host(const uint64_t *data, int64_t n_chunks, int M, int n, const uint64_t *tables,
                      int64_t *indices, int32_t *vals, int R, int order, int signd,
                      const int64_t *labels);

/* replaces init_heap / insert / insert_is (ref: tinyknn/_fast_pq.pyx:240-252, :274-307, :256-271).
 * Pure host functions on the caller's arrays (they are O(R) scalar updates of host memory; the
 * same insert code is compiled as the device function the replay kernel uses). */
TKB_API int tkb_init_heap(int64_t *indices, int32_t *vals, int R, int signd);
TKB_API int tkb_insert(int64_t *indices, int32_t *vals, int R, int64_t label, int v);
TKB_API int tkb_insert_is(int64_t *indices, int32_t *vals, int R, int64_t label, int v);

/* ------------------------------------------------------------------------------------------ */
/* Device surface (all pointers are device memory unless stated; stream = cudaStream_t)        */
/* ------------------------------------------------------------------------------------------ */

/* Batched LUT construction: replaces FastPQ.distance_table (ref: tinyknn/fast_pq.py:186-222) when
 * signd != 0 and FastPQ.udistance_table (ref: tinyknn/fast_pq.py:224-252) when signd == 0, for Q
 * queries at once, including pad1 (ref: tinyknn/utils.py:6-11), the optional rotation q @ R.T and
 * transform_tables (ref: tinyknn/_transform.py:114-138).
 *   queries  f32[Q][d]            raw queries
 *   normalize != 0                first q /= ||q|| in f32 (ref: tinyknn/ivf.py:126-127); the
 *                                 normalised query is written to q_out f32[Q][d] (may alias queries)
 *   centers  f32[16][Dp]          FastPQ.centers
 *   R        f64[Dp][Dpad] or NULL
 *   sqrt_n_blocks, log_n_blocks   host-computed np.sqrt(M) / np.log(M) (f64)
 *   tables   uint8[Q][M][16]      out (== uint64[Q][2M])
 *   q_rot    f64[Q][Dp] or NULL   out: rotated (padded) query, as f64
 *   shift, scale f64[Q] or NULL   out: _FastDistanceTable.mean / .scale
 */
TKB_API int tkb_lut_build_dev(const float *queries, int Q, int d, int normalize, float *q_out,
                      const float *centers, int Dp, int dpb, const double *R, int Dpad,
                      double sqrt_n_blocks, double log_n_blocks, int signd,
                      uint8_t *tables, double *q_rot, double *shift, double *scale, void *stream);

/* Batched estimate: est[q][pos] for Q LUTs over the same codes.
 *   tables uint8[Q][M][16]; est uint8[Q][est_stride], est_stride >= 16*n_chunks. */
TKB_API int tkb_estimate_dev(const uint64_t *codes, int64_t n_chunks, int M, const uint8_t *tables, int Q,
                     uint8_t *est, int64_t est_stride, int order, int signd, void *stream);

/* Where the estimates of an IVF scan live. Segment (q, s) = the estimates of list probes[q][s] under LUT q,
 * one byte per vector position, 16 * ceil(list_size / 16) bytes (the reference pads a list to whole chunks,
 * ref: tinyknn/fast_pq.py:165-169). Two addressing modes, chosen by `seg_off`:
 *   seg_off == NULL : est + (q * P + s) * slot_stride            (slot_stride >= 16 * chunks of the largest list)
 *   seg_off != NULL : est + seg_off[q * P + s], int64[Q][P] from tkb_ivf_plan_dev; a negative offset means the
 *                     segment is not in this buffer (list owned by another rank) and is skipped.
 * Lists live in one codes array: list l = chunks [list_chunk_off[l], list_chunk_off[l+1]); list_size int32[n_lists]
 * (may be NULL for the scans: then every stored chunk of the list is scanned, padding included).
 * probes int32[Q][P]: negative entries index from the end, like a Python list (ref: tinyknn/ivf.py:141);
 * INT32_MIN marks a probe slot that does not exist. */

/* IVF scan on the reference layout (step-by-step kernel). max_list_chunks: chunks of the largest list (0: slot_stride/16). */
TKB_API int tkb_ivf_scan_dev(const uint64_t *codes, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                     const uint8_t *tables, const int32_t *probes, int Q, int P,
                     uint8_t *est, int64_t slot_stride, const int64_t *seg_off, int64_t max_list_chunks,
                     int order, int signd, void *stream);

/* Segment planning (compact estimate buffers; the exchange of a list-sharded index, DESIGN.md "multi-GPU").
 *   list_owner int32[n_lists] or NULL (n_ranks == 1): the rank that stores each list
 *   TKB_PLAN_SEND: all Q queries, segments whose list this rank owns, grouped by the home rank of the query
 *                  (q / q_per_rank), then (q, s)         -> seg_off int64[Q][P]
 *   TKB_PLAN_RECV: the home queries [rank*q_per_rank, (rank+1)*q_per_rank), every segment, grouped by the owner
 *                  of the list, then (q, s)              -> seg_off int64[q_per_rank][P]  (row 0 = first home query)
 *   group_bytes int64[2*n_ranks+1] out: bytes per group (the all-to-all split sizes), the total, the group bases
 *   workspace: 8 * queries * n_ranks bytes */
#define TKB_PLAN_SEND 0
#define TKB_PLAN_RECV 1
#define TKB_PLAN_PUSH 2   /* tkb_ivf_plan_push_dev only */
TKB_API int tkb_ivf_plan_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                     int n_lists, int mode, int rank, int n_ranks, int q_per_rank,
                     int64_t *seg_off, int64_t *group_bytes, void *workspace, int64_t workspace_bytes, void *stream);

/* Push plan (fused scan + exchange over NVLink peer memory, DESIGN.md "multi-GPU"): the home rank of a query keeps the
 * estimates of ALL its segments packed in (query, probe slot) order -- exactly the single-GPU layout
 * (tkb_ivf_plan_dev, n_ranks == 1, over its own block of queries) -- and every scanning rank writes the segments of the
 * lists it owns straight into that buffer through a peer mapping. For all Q queries:
 *   seg_addr int64[Q][P] = home_base[q / q_per_rank] + offset of (q, s) inside the home buffer when this rank owns the
 *                          list, else -1. home_base int64[n_ranks] (device): the address of every rank's receive buffer
 *                          as mapped in THIS process (tkb_peer_open; the local pointer for rank == own).
 *   group_bytes[g] = bytes of home rank g's buffer (all segments), [n_ranks] = their sum.
 * Pass seg_addr as `seg_off` with est == NULL to tkb_ivf_scan_native_dev: the offsets are then absolute addresses. */
TKB_API int tkb_ivf_plan_push_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                          int n_lists, int rank, int n_ranks, int q_per_rank, const int64_t *home_base,
                          int64_t *seg_addr, int64_t *group_bytes, void *workspace, int64_t workspace_bytes, void *stream);

/* Pull exchange (DESIGN.md "multi-GPU"): every rank scans the lists it OWNS into its OWN estimate buffer (TKB_PLAN_SEND layout:
 * grouped by the home rank of the query, then (q, s); local stores at full speed, chunk minima next to them), the ranks
 * exchange the G + 1 numbers (total, group bases) of their layouts, and the home rank's replay reads the chunk minima of its
 * queries from a local copy (tkb_ivf_pull_minima_dev) and fetches only the chunks that can hold a candidate -- a few percent --
 * from the owners' buffers through the peer mappings.
 *   tkb_ivf_plan_pull_owner_dev: TKB_PLAN_SEND with a capacity guard: a segment that would end past `capacity` bytes gets
 *       offset -1 (never written); group_bytes[n_ranks] still holds the full total, so an overflow is visible to every rank
 *       after the exchange of those numbers.
 *   tkb_ivf_plan_pull_home_dev: for the home queries [rank*q_per_rank, ...): seg_addr int64[q_per_rank][P] = the ABSOLUTE
 *       address of segment (q, s) inside its owner's buffer = owner_base[o] + owner_groups[o][1 + rank] + offset inside the
 *       (owner o, home rank) group. owner_base int64[n_ranks]: every rank's buffer as mapped in this process;
 *       owner_groups int64[n_ranks][n_ranks + 1]: row o = group_bytes[n_ranks .. 2 n_ranks] of owner o's plan.
 *   tkb_ivf_pull_minima_dev: copies the minima of every segment of the home queries from the owners
 *       (cm_table[o] + (seg_addr >> 4)) to cmin_local[seg_local >> 4], seg_local = the compact single-GPU plan of the home
 *       queries (tkb_ivf_plan_dev, n_ranks == 1).
 *   tkb_ivf_replay_fresh_pull_dev: tkb_ivf_replay_fresh_cm_dev whose estimate reads go to seg_addr (absolute addresses),
 *       while the minima are addressed through cm_seg_off (= seg_local). cmin == NULL: no minima, every chunk is fetched. */
TKB_API int tkb_ivf_plan_pull_owner_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                                int n_lists, int rank, int n_ranks, int q_per_rank, int64_t capacity,
                                int64_t *seg_off, int64_t *group_bytes, void *workspace, int64_t workspace_bytes, void *stream);
TKB_API int tkb_ivf_plan_pull_home_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                               int n_lists, int rank, int n_ranks, int q_per_rank, const int64_t *owner_base,
                               const int64_t *owner_groups, int64_t *seg_addr, int64_t *group_bytes, void *workspace,
                               int64_t workspace_bytes, void *stream);
TKB_API int tkb_ivf_pull_minima_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner, int n_lists,
                            const int64_t *seg_addr, const int64_t *seg_local, const int64_t *cm_table, uint8_t *cmin_local,
                            void *stream);
TKB_API int tkb_ivf_replay_fresh_pull_dev(const int64_t *seg_addr, const int64_t *cm_seg_off, const uint8_t *cmin,
                                  const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, const int64_t *ids,
                                  const int32_t *probes, int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                                  int unique_labels, int32_t *fallback, void *stream);

/* Peer-visible device buffers for the push exchange: one process per GPU, so the buffers are shared as CUDA IPC handles
 * (64 bytes) that the caller moves between the processes of a box itself (torch.distributed all_gather in the host layer).
 * tkb_peer_alloc: cudaMalloc on the current device + its handle. tkb_peer_open: map another process's buffer into this one
 * (peer access over NVLink is enabled lazily); not valid in the allocating process. close / free undo them. */
#define TKB_PEER_HANDLE_BYTES 64
TKB_API int tkb_peer_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle);
TKB_API int tkb_peer_open(const unsigned char *handle, void **dev_ptr);
TKB_API int tkb_peer_close(void *dev_ptr);
TKB_API int tkb_peer_free(void *dev_ptr);

/* Batched PQ encoder (build time): replaces FastPQ.transform (ref: tinyknn/fast_pq.py:147-184) with
 * pad2 (ref: tinyknn/utils.py:14-19), the rotation `data @ R.T`, the per-block nearest-of-16 search of knn_brute
 * (ref: tinyknn/utils.py:66-86) and the nibble packing of transform_data (ref: tinyknn/_transform.py:4-77).
 * The arithmetic mirrors numpy operation by operation (csrc/tkb_encode.cu), so codes equal the reference's wherever
 * the nearest codeword is unique.
 *   rows      f32/f64 [n_rows][d]      vectors (already normalised for angular indexes, ref: tinyknn/ivf.py:80-82)
 *   row_index int64[n_out] or NULL     output position i encodes rows[row_index[i]]; an index outside [0, n_rows) is
 *                                      the zero vector (the reference's padding rows). NULL: position i = row i.
 *   n_out     multiple of 16           positions to encode (chunks of 16, lists padded by the caller)
 *   centers   f32[16][Dp], cnorm f32[Dp/dpb][16] = per block |c|^2 in f32 (np.einsum on the f32 codebook)
 *   R         f64[Dp][Dpad] or NULL    Dpad = d rounded up to dpad * dpb (ref: fast_pq.py:161-169); Dp <= 128 when R
 *   codes     uint64[n_out/16][Dp/dpb] out, reference layout (feed tkb_codes_to_native_dev for the scan layout) */
TKB_API int tkb_encode_dev(const void *rows, int rows_dtype, int64_t n_rows, int d, const int64_t *row_index, int64_t n_out,
                   const float *centers, const float *cnorm, int Dp, int dpb, const double *R, int Dpad,
                   uint64_t *codes, void *stream);

/* Coarse assignment (build time): replaces the arithmetic of knn_brute(data, all_centers, k) in
 * IVF.build (ref: tinyknn/ivf.py:84-86, tinyknn/utils.py:66-86): part = (|x|^2 + |c|^2) - (2x).c with the dot product as
 * the FMA chain over the dimension that the reference's sgemm/dgemm computes (bit-identical for d <= 384), the k in
 * {1, 2} smallest per row in ascending (part, index) order. rows/centers/xnorm/cnorm share one dtype (f32 or f64).
 *   xnorm [n], cnorm [C] : |x|^2 / |c|^2 as the caller's numpy computed them (np.einsum; pass them for bit parity), or
 *                          NULL: computed here left to right into scratch ((n + C) elements)
 *   nearest int32[n][k]  : out. k = 1 reproduces the reference's assignment wherever the minimum is unique; for k = 2 the
 *                          SET per row is the reference's, the order inside a row is not (np.argpartition's is unspecified) */
TKB_API int tkb_assign_dev(const void *rows, int dtype, int64_t n, int d, const void *centers, int C, const void *xnorm,
                   const void *cnorm, int k, int32_t *nearest, void *scratch, int64_t scratch_bytes, void *stream);

/* k-means on the device for the two `fit` steps (build time; replaces the arithmetic of sklearn.cluster.KMeans in
 * tinyknn/ivf.py:19-51 and tinyknn/fast_pq.py:106-145). PARITY UNPINNED by nature (the reference's k-means++ seeding draws from
 * numpy's global random state); deterministic given (rows, initial centers): cluster sums are fixed-point integer atomics.
 *   tkb_kmeans_dev: rows f32 [n][d], centers f32 [k][d] in (initial) / out, Lloyd iterations until no row changes its cluster or
 *       max_iters; assign int32[n] out = nearest centre of every row for the returned centers (tkb_assign_dev's exact chains);
 *       *iters_done (host) = iterations run; absmax >= max |rows[i][j]| (scales the fixed point). Synchronises the stream once
 *       per iteration.
 *   tkb_kmeans_pq_dev: the M = D / dpb independent 16-centre problems of FastPQ in one pass over the rows per iteration;
 *       centers f32 [16][D] in FastPQ's layout (centre c of block m at [c][m*dpb .. (m+1)*dpb)) in / out; dpb <= 8;
 *       workspace >= 16 KB + 160 * D bytes. */
TKB_API int tkb_kmeans_workspace(int64_t n, int d, int k, int64_t *bytes);
TKB_API int tkb_kmeans_dev(const float *rows, int64_t n, int d, int k, float *centers, int max_iters, double absmax, int32_t *assign,
                   int *iters_done, void *workspace, int64_t workspace_bytes, void *stream);
TKB_API int tkb_kmeans_pq_dev(const float *rows, int64_t n, int D, int dpb, float *centers, int iters, double absmax, void *workspace,
                      int64_t workspace_bytes, void *stream);

/* Device-native code layout for the fast scan (chosen at upload, round-trips to the reference layout).
 * tile = 8 chunks; the 16 bytes of (tile t, pair p, chunk slot s) sit at ((t*M/2 + p)*8 + s)*16 and hold
 * 8 halfwords: halfword g = codes of sub-quantizer 2p for vectors 4g..4g+3 (nibble i = vector 4g+i),
 * halfword 4+g = the same for sub-quantizer 2p+1. `native` needs ceil(n_chunks/8)*8 * M*8 bytes; padding
 * chunks are zero codes. In a native IVF array every list starts on a tile boundary (list_chunk_off % 8 == 0). */
TKB_API int tkb_codes_to_native_dev(const uint64_t *codes, int64_t n_chunks, int M, void *native, void *stream);
TKB_API int tkb_codes_from_native_dev(const void *native, int64_t n_chunks, int M, uint64_t *codes, void *stream);

/* Fast scan on the native layout: same results as tkb_estimate_dev / tkb_ivf_scan_dev, bit for bit.
 * LUT rows are looked up with PRMT; sums are accumulated without per-step clamps and a per-vector certificate decides
 * which chunks are recomputed with the reference's step-by-step fold -- inside the same kernel, by the CTA that found
 * them.
 * workspace: optional 16-byte aligned device scratch of >= 16 bytes; its first 8 bytes return the number of chunks whose
 * certificate failed (a statistic, reset by every call). NULL is allowed.
 * max_chunks_per_query (ivf): upper bound used to size the grid (0: P * slot_stride / 16).
 * tkb_ivf_scan_native_dev with est == NULL and seg_off != NULL: seg_off holds absolute device addresses (possibly
 * peer-mapped memory of another GPU), negative = skip (tkb_ivf_plan_push_dev). */
TKB_API int tkb_estimate_native_dev(const void *native, int64_t n_chunks, int M, const uint8_t *tables, int Q,
                            uint8_t *est, int64_t est_stride, int order, int signd,
                            void *workspace, int64_t workspace_bytes, void *stream);
TKB_API int tkb_ivf_scan_native_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                            const uint8_t *tables, const int32_t *probes, int Q, int P,
                            uint8_t *est, int64_t slot_stride, const int64_t *seg_off, int64_t max_chunks_per_query,
                            int order, int signd, void *workspace, int64_t workspace_bytes, void *stream);

/* Exact replay of the reference heap (ref: tinyknn/_fast_pq.pyx:153-206, :274-307) over
 * precomputed estimates, one heap per query.
 *   heap_idx int64[Q][R], heap_val int32[Q][R]: in/out (call tkb_heap_fill_dev first for a fresh heap)
 * tkb_replay_dev : one segment per query: est[q][0..16*n_chunks), true size n, labels int64[>=n] or NULL
 * tkb_ivf_replay_dev : P segments per query in probe order; segment s = list probes[q][s] with
 *   n = list_size[l] and labels = ids + 16*list_chunk_off[l] (ids are stored padded like the codes; here
 *   list_chunk_off addresses `ids`, which a rank of a sharded index keeps for ALL lists). */
TKB_API int tkb_heap_fill_dev(int64_t *heap_idx, int32_t *heap_val, int64_t count, int signd, void *stream);
TKB_API int tkb_replay_dev(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n,
                   int64_t *heap_idx, int32_t *heap_val, int Q, int R, int signd,
                   const int64_t *labels, void *stream);
TKB_API int tkb_ivf_replay_dev(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                       const int32_t *list_size, int n_lists, const int64_t *ids,
                       const int32_t *probes, int Q, int P,
                       int64_t *heap_idx, int32_t *heap_val, int R, int signd, void *stream);

/* Same replays for a FRESH heap (init_heap + query_pq in one call; the heap arrays are outputs only).
 * These use the queue-replay kernel (producer warps filter the estimates, one lane per query replays; heaps
 * in shared memory) when its preconditions hold and the warp-per-query kernel otherwise; results are identical.
 * unique_labels != 0 certifies that no label occurs twice among the lists (an index built with one list per
 * point), which makes the reference's label dedupe a no-op. fallback: int32[Q] device scratch. */
TKB_API int tkb_replay_fresh_dev(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n,
                         int64_t *heap_idx, int32_t *heap_val, int Q, int R, int signd, void *stream);
TKB_API int tkb_ivf_replay_fresh_dev(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                             const int32_t *list_size, int n_lists, const int64_t *ids,
                             const int32_t *probes, int Q, int P,
                             int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                             int unique_labels, int32_t *fallback, void *stream);

/* Chunk minima for long probe lists. The scan additionally stores, for every 16-vector chunk, the smallest
 * of its estimates: cmin[o / 16] for the chunk whose estimates sit at est + o (cmin: est bytes / 16 + 16 bytes, 16-byte
 * aligned, same signedness as est). The replay of a query whose segments lie back to back in `est` (the compact plan of
 * tkb_ivf_plan_dev with n_ranks == 1) then tests 16 chunks per 16-byte load and fetches a chunk's estimates only when its
 * minimum is below the (stale, hence conservative) bound of the round; other queries take the ordinary path. Results are
 * identical to tkb_ivf_scan_native_dev + tkb_ivf_replay_fresh_dev, bit for bit (the heap arrays included). */
TKB_API int tkb_ivf_scan_native_cm_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                               const uint8_t *tables, const int32_t *probes, int Q, int P,
                               uint8_t *est, const int64_t *seg_off, uint8_t *cmin, int64_t max_chunks_per_query,
                               int order, int signd, void *workspace, int64_t workspace_bytes, void *stream);
/* The same scan on the tensor cores, LIST-MAJOR (csrc/tkb_scan_tc.cu; tcgen05.mma kind::i8, one-hot codes in tensor memory):
 * the (query, probe slot) pairs of the batch are grouped by list, a 128-vector tile of a list is expanded to one-hot form once
 * and multiplied with the LUTs of all the queries that probe the list. Same bytes in `est` (and `cmin`, optional) as
 * tkb_ivf_scan_native_cm_dev: signed tables, avx accumulation order (ref: _fast_pq_256.pyx:126-156), M = 32. Pairs the
 * tensor-core path cannot certify are refolded step by step; queries whose LUT fails the per-query precondition are scanned
 * by the CUDA-core kernel inside the same call. workspace: tkb_ivf_scan_tc_workspace bytes, 16-byte aligned; afterwards its
 * third int32 holds the number of refolded (vector, query) pairs. Worth it when several queries probe the same list.
 * Inside the push exchange: est == NULL, seg_off = absolute addresses (tkb_ivf_plan_push_dev), cm_home / q_per_rank as for
 * tkb_ivf_scan_native_push_cm_dev (cm_home may be NULL: no minima); otherwise cm_home == NULL, q_per_rank == 0. */
TKB_API int tkb_ivf_scan_tc_supported(void);      /* 1: the current device has tcgen05 tensor cores (compute capability 10.x) */
TKB_API int tkb_ivf_scan_tc_workspace(int Q, int P, int n_lists, int64_t *bytes);
TKB_API int tkb_ivf_scan_tc_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                        const uint8_t *tables, const int32_t *probes, int Q, int P,
                        uint8_t *est, const int64_t *seg_off, uint8_t *cmin, const int64_t *cm_home, int q_per_rank,
                        int64_t max_chunks_per_query, void *workspace, int64_t workspace_bytes, void *stream);
/* The same scan inside the push exchange (est == NULL, seg_addr = absolute addresses from tkb_ivf_plan_push_dev): the
 * minima of a segment go to the home rank's minima region. cm_home int64[n_ranks] (device): for home rank h,
 * (address of its minima region) - (address of its estimate buffer >> 4), both as mapped in THIS process; the home rank of
 * query q is q / q_per_rank. */
TKB_API int tkb_ivf_scan_native_push_cm_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                                    const uint8_t *tables, const int32_t *probes, int Q, int P,
                                    const int64_t *seg_addr, const int64_t *cm_home, int q_per_rank, int64_t max_chunks_per_query,
                                    int order, int signd, void *workspace, int64_t workspace_bytes, void *stream);
TKB_API int tkb_ivf_replay_fresh_cm_dev(const uint8_t *est, const int64_t *seg_off, const uint8_t *cmin, const int64_t *list_chunk_off,
                                const int32_t *list_size, int n_lists, const int64_t *ids,
                                const int32_t *probes, int Q, int P,
                                int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                                int unique_labels, int32_t *fallback, void *stream);

/* Exact rescoring distances: replaces the arithmetic of knn_brute1 (ref: tinyknn/utils.py:89-92).
 *   dists[q][r] = sum_i (rows[idx[q][r]][i] - queries[q][i])^2, computed in the rows' dtype
 *   (f32 rows: f32 like numpy; f64 rows: f64). Negative idx index from the end (numpy semantics:
 *   the reference rescoring may see -1 heap padding, ref: tinyknn/fast_pq.py:311). */
TKB_API int tkb_gather_dists_dev(const void *rows, int rows_dtype, int64_t n_rows, int d,
                         const float *queries, const int64_t *idx, int Q, int R,
                         void *dists, void *stream);

/* Device-side selection (deterministic order: ascending distance, ties by slot).
 * tkb_select_probes_dev replaces `indices[best]` of _FastDistanceTable.top (ref: tinyknn/fast_pq.py:307-312)
 *   for the probe lists: out int32[Q][P]; when R <= P the raw heap is returned (first R slots).
 * tkb_select_topk_dev replaces ivf.py:154-163: drops -1, returns all survivors when <= k, else the
 *   k nearest. out_ids int64[Q][k] (padded with -1), out_dists f32/f64[Q][k], out_count int32[Q]. */
TKB_API int tkb_select_probes_dev(const int64_t *heap_idx, const void *dists, int dists_dtype, int Q, int R,
                          int P, int32_t *probes, void *stream);
TKB_API int tkb_select_topk_dev(const int64_t *heap_idx, const void *dists, int dists_dtype, int Q, int R,
                        int k, int64_t *out_ids, void *out_dists, int32_t *out_count, void *stream);

/* Probe selection as ONE kernel: replaces `dtable.top(pq_transformed_centers, active_centers, k=n_probes)` of IVF.query
 * (ref: tinyknn/ivf.py:131 -> tinyknn/fast_pq.py:284-312) for Q queries: the scan of the PQ-encoded centroids, the exact
 * replay of the reference heap of R = min(2 * n_probes + 10, C) slots (signed tables, label = centroid position), the exact
 * distances of the candidates to the raw centroids (knn_brute1) and the P nearest in device order. Same results as
 * tkb_estimate_native_dev + tkb_replay_fresh_dev + tkb_gather_dists_dev + tkb_select_probes_dev, bit for bit (R <= P: the
 * raw heap, INT32_MIN in the missing slots). R <= 1024.
 *   native_centers : encoded centroids in the native layout (n_chunks chunks, C real vectors); centers f32[C][d]
 *   tables uint8[Q][M][16]; queries f32[Q][d] (normalised for angular); probes int32[Q][P] out
 *   heap_idx int64[Q][R], heap_val int32[Q][R], dists f32[Q][R] : optional outputs (NULL to skip) */
TKB_API int tkb_coarse_probes_dev(const void *native_centers, int64_t n_chunks, int C, int M, const uint8_t *tables, int Q,
                          const float *centers, int d, const float *queries, int R, int P, int order, int32_t *probes,
                          int64_t *heap_idx, int32_t *heap_val, float *dists, void *stream);

/* The IVF.query body after probe selection as ONE kernel (ref: tinyknn/ivf.py:135-163): for each query, scan of
 * the probed lists (query_pq_*'s scan half), exact replay of the reference heap of R slots in probe order
 * (query_pq_*'s heap half with labels = ids), removal of the -1 padding, exact distances of the candidates
 * (knn_brute1, ref: tinyknn/utils.py:89-92) and the k nearest (device order: ascending distance, ties by heap
 * slot). Same results as tkb_ivf_scan_native_dev + tkb_ivf_replay_fresh_dev + tkb_gather_dists_dev +
 * tkb_select_topk_dev, bit for bit; signed tables only (ref: tinyknn/ivf.py:138,148).
 * Requires labels that are unique across the lists (one list per point; see tkb_ivf_replay_fresh_dev). Negative
 * probe entries index from the end like a Python list and may visit a list twice: the reference's label dedupe
 * is reproduced.
 *   native/list_chunk_off/list_size/ids : the index, as for tkb_ivf_scan_native_dev / tkb_ivf_replay_dev
 *   tables uint8[Q][M][16], probes int32[Q][P], queries f32[Q][d] (normalised for angular), rows [n_rows][d]
 *   max_list_chunks : ceil(size/16) of the largest list
 *   out_ids int64[Q][k] (-1 padded), out_dists f32/f64[Q][k] or NULL, out_count int32[Q]
 *   heap_idx int64[Q][R] / heap_val int32[Q][R] : optional outputs, the reference's final heap arrays
 *   workspace : 16-byte aligned device scratch of tkb_ivf_query_fused_workspace() bytes (a smaller one limits
 *               the number of resident CTAs); its second 8-byte word returns the number of recomputed chunks */
TKB_API int tkb_ivf_query_fused_workspace(int Q, int P, int R, int M, int order, int rows_dtype,
                                  int64_t max_list_chunks, int64_t *bytes);
TKB_API int tkb_ivf_query_fused_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists,
                            int M, const uint8_t *tables, const int32_t *probes, int Q, int P, const int64_t *ids,
                            const void *rows, int rows_dtype, int64_t n_rows, int d, const float *queries,
                            int R, int k, int order, int64_t max_list_chunks,
                            int64_t *out_ids, void *out_dists, int32_t *out_count,
                            int64_t *heap_idx, int32_t *heap_val,
                            void *workspace, int64_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYKNN_B200_H */
