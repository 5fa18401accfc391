/*
 * absb200.h — C ABI of libabsb200.so, the B200-native (sm_100a) implementation of the
 * abstracts-search hot path: stella_en_1.5B_v5 bulk embedding + faiss-style IVF/Flat
 * inner-product top-k search.
 *
 * The reference (colonelwatch/abstracts-search) reaches this path only through the Python
 * surfaces `SentenceTransformer.encode()` and `faiss.Index.{train,add,search}`:
 *   - bulk encode          /root/reference/Makefile:65   (sidecar-search build -b 32)
 *   - Index.train          /root/reference/Makefile:38-39 (sidecar-search index train, -c 65536 README.md:60)
 *   - Index.add            /root/reference/Makefile:24-25 (sidecar-search index fill)
 *   - Index.search         /root/reference/Makefile:31-32 (sidecar-search index tune), README.md:16,28 (app.py)
 * Every entry point below names the reference interface it stands in for.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only. Every function returns 0 (ABSB_OK) or a negative
 *     error code; the message is available from absb_last_error() (thread-local).
 *   - No exceptions cross the boundary. The caller owns all host buffers; they are borrowed for
 *     the duration of the call. Opaque handles own device memory.
 *   - Functions without a suffix take HOST pointers (numpy arrays in the Python binding) and do the
 *     host<->device copies themselves. `_dev` variants take DEVICE pointers on the handle's device
 *     plus a cudaStream_t passed as void* (NULL = the CUDA legacy default stream, as in any CUDA
 *     API) and do not synchronise unless stated.
 *   - One handle per host thread (thread-compatible, no hidden globals besides the error string).
 *   - There is NO CPU fallback: without a usable sm_100 device the create calls fail.
 */
#ifndef ABSB200_H
#define ABSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABSB_OK 0
#define ABSB_ERR_INVALID (-1)     /* bad argument (shape, dtype contract, k too large, ...)   */
#define ABSB_ERR_CUDA (-2)        /* CUDA runtime / driver failure                              */
#define ABSB_ERR_STATE (-3)       /* index not trained, encoder weights missing, ...            */
#define ABSB_ERR_UNSUPPORTED (-4) /* valid in faiss but outside this path (e.g. METRIC_L2)      */
#define ABSB_ERR_OOM (-5)         /* device allocation failed                                   */

/* faiss MetricType values (faiss/MetricType.h); only inner product is on the path. */
#define ABSB_METRIC_INNER_PRODUCT 0
#define ABSB_METRIC_L2 1

/* Largest k / nprobe the fused warp top-k supports. */
#define ABSB_MAX_K 256

typedef struct absb_ivf_s* absb_ivf_t;
typedef struct absb_flat_s* absb_flat_t;
typedef struct absb_enc_s* absb_enc_t;
typedef struct absb_peer_s* absb_peer_t;

/* ---------------------------------------------------------------- library ------------------ */
int absb_version(void);
const char* absb_last_error(void);
/* Number of usable CUDA devices; fails (ABSB_ERR_CUDA) when there is no driver/device. */
int absb_device_count(int* count);
/* Name + SM count + compute capability of a device (buffer of at least 128 bytes). */
int absb_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- synthetic corpora -------- */
/* Counter-based integer-lattice generator shared bit-for-bit with oracle/synth.py.
 * kind 0: corpus rows   x[r] = (mu[c(r)] + eps(r)) / 128
 * kind 1: centroids     c[j] =  mu[j] / 128           (row index = list number)
 * kind 2: queries       q[i] = clamp(mu[c(s)] + eps(s) + delta(i)) / 128, s = source row of query i
 * kind | 4: the same integer row divided by its L2 norm instead of by 128 — real-valued unit-length
 *           fp32 rows (what `index fill` stores, /root/reference/Makefile:24-25), not representable in
 *           fp16, still bit-identical to oracle/synth.py (exact integer norm, IEEE sqrt and divide).
 * out is a DEVICE pointer to [n, d] float32. */
int absb_synth_fill_dev(int kind, uint64_t seed, int64_t row0, int64_t n, int d, int nlist,
                        int64_t corpus_rows, float* out_dev, void* stream);
/* Same generator for an explicit list of row numbers rows_dev [n] (DEVICE int64): lets a shard
 * materialise only the rows whose inverted list it owns. */
int absb_synth_fill_rows_dev(int kind, uint64_t seed, const int64_t* rows_dev, int64_t n, int d,
                             int nlist, int64_t corpus_rows, float* out_dev, void* stream);
/* cluster id c(r) of corpus rows [row0,row0+n) -> DEVICE int64 (feeds add_preassigned) */
int absb_synth_cluster_dev(uint64_t seed, int64_t row0, int64_t n, int nlist, int64_t* out_dev,
                           void* stream);

/* ---------------------------------------------------------------- IndexFlatIP -------------- */
/* faiss.IndexFlatIP(d): exact inner-product search; the IVF coarse quantiser and config 1. */
int absb_flat_create(int d, int metric, int device, absb_flat_t* out);
int absb_flat_destroy(absb_flat_t h);
int absb_flat_reset(absb_flat_t h);
int absb_flat_ntotal(absb_flat_t h, int64_t* ntotal);
/* Index.add(x): x [n,d] float32 C-contiguous. */
int absb_flat_add(absb_flat_t h, int64_t n, const float* x);
int absb_flat_add_dev(absb_flat_t h, int64_t n, const float* x_dev, void* stream);
/* Index.search(x,k) -> D [n,k] f32 descending, I [n,k] i64; missing: I=-1, D=-FLT_MAX. */
int absb_flat_search(absb_flat_t h, int64_t n, const float* q, int k, float* D, int64_t* I);
int absb_flat_search_dev(absb_flat_t h, int64_t n, const float* q_dev, int k, float* D_dev,
                         int64_t* I_dev, void* stream);
/* Index.reconstruct_n(i0, n) */
int absb_flat_reconstruct(absb_flat_t h, int64_t i0, int64_t n, float* x);

/* ---------------------------------------------------------------- IndexIVFFlat ------------- */
/* faiss.index_factory(d, "IVF<nlist>,Flat", METRIC_INNER_PRODUCT)  (SURVEY §8a a4). */
int absb_ivf_create(int d, int nlist, int metric, int device, absb_ivf_t* out);
int absb_ivf_destroy(absb_ivf_t h);
/* Index.reset(): drop inverted lists, keep centroids. */
int absb_ivf_reset(absb_ivf_t h);
int absb_ivf_ntotal(absb_ivf_t h, int64_t* ntotal);
int absb_ivf_is_trained(absb_ivf_t h, int* trained);

/* faiss ClusteringParameters (index.cp): niter=10, max_points_per_centroid=256,
 * min_points_per_centroid=39, seed=1234 by default. */
int absb_ivf_set_clustering(absb_ivf_t h, int niter, int max_points_per_centroid,
                            int min_points_per_centroid, int64_t seed);
/* faiss ClusteringParameters.spherical: L2-normalise the centroids after the initial draw and after
 * every Lloyd iteration (Clustering::post_process_centroids -> fvec_renorm_L2).  The struct default is
 * false; faiss's index_factory switches it on for METRIC_INNER_PRODUCT IVF indexes — the
 * `index train` call of /root/reference/Makefile:38-39 goes through index_factory (external, faiss
 * unpinned: the Python index_factory mirrors that default, IndexIVFFlat(...) keeps false). */
int absb_ivf_set_clustering_spherical(absb_ivf_t h, int spherical);
/* In-place fvec_renorm_L2 of [n, d] device rows (zero rows stay zero): the step above, exported for
 * the distributed trainer. */
int absb_renorm_rows_dev(int device, int64_t n, int d, float* x_dev, void* stream);
/* Index.train(x)  (SURVEY §8a a5): subsample, random-row init, niter Lloyd iterations with
 * arg-max-IP assignment, mean update, empty-cluster split. x is a HOST pointer [n,d]. */
int absb_ivf_train(absb_ivf_t h, int64_t n, const float* x);
/* Same with x already on the device (the sample gather happens on the device). */
int absb_ivf_train_dev(absb_ivf_t h, int64_t n, const float* x_dev, void* stream);
/* quantizer.add(centroids) / quantizer.reconstruct_n: import / export the coarse centroids
 * [nlist,d] (host). Importing marks the index trained. */
int absb_ivf_set_centroids(absb_ivf_t h, const float* centroids);
int absb_ivf_get_centroids(absb_ivf_t h, float* centroids);
int absb_ivf_set_centroids_dev(absb_ivf_t h, const float* centroids_dev, void* stream);

/* Index.add(x) (ids == NULL: id = ntotal + i) / Index.add_with_ids(x, ids)  (SURVEY §8a a6). */
int absb_ivf_add(absb_ivf_t h, int64_t n, const float* x, const int64_t* ids);
int absb_ivf_add_dev(absb_ivf_t h, int64_t n, const float* x_dev, const int64_t* ids_dev,
                     void* stream);
/* IndexIVF::add_core(n, x, xids, precomputed list numbers): list_ids [n] int64, -1 = skip row. */
int absb_ivf_add_preassigned(absb_ivf_t h, int64_t n, const float* x, const int64_t* ids,
                             const int64_t* list_ids);
int absb_ivf_add_preassigned_dev(absb_ivf_t h, int64_t n, const float* x_dev,
                                 const int64_t* ids_dev, const int64_t* list_ids_dev,
                                 void* stream);
/* Make every list's pages physically consecutive (optional; changes no result, only the number of
 * scan work items after an index was filled by many small add() calls, e.g. streamed row groups
 * of `sidecar-search index fill`, /root/reference/Makefile:24-25).  In place: the pages are permuted
 * window by window through a bounded scratch (<= 4 GB, or `scratch_pages` pages), so a 106 GB
 * shard compacts inside 180 GB.  Synchronises. */
int absb_ivf_compact(absb_ivf_t h);
int absb_ivf_compact_scratch(absb_ivf_t h, int64_t scratch_pages);
/* Test hook (host only): the compaction schedule for a page table `src` [n] (a permutation;
 * content_new[t] = content_old[src[t]]).  moves [n_moves, 2] = (from, to), >= 0 pool page, < 0
 * scratch slot -1 - v; phase_end [n_phases] = exclusive end of each phase in `moves`; the copies
 * of one phase are independent, phases run in order. */
int absb_plan_page_compaction(int64_t n, const int32_t* src, int64_t scratch_pages, int32_t* moves,
                              int64_t moves_cap, int64_t* phase_end, int64_t phases_cap,
                              int64_t* n_moves, int64_t* n_phases);

/* quantizer.search(x, nprobe): Dc [n,nprobe] f32, Ic [n,nprobe] i64, best first. */
int absb_ivf_coarse(absb_ivf_t h, int64_t n, const float* q, int nprobe, float* Dc, int64_t* Ic);
int absb_ivf_coarse_dev(absb_ivf_t h, int64_t n, const float* q_dev, int nprobe, float* Dc_dev,
                        int64_t* Ic_dev, void* stream);
/* quantizer.assign(x): list number per row. */
int absb_ivf_assign(absb_ivf_t h, int64_t n, const float* x, int64_t* list_ids);

/* Index.search(x, k) with IndexIVF.nprobe = nprobe  (SURVEY §8a a7, a8). */
int absb_ivf_search(absb_ivf_t h, int64_t n, const float* q, int k, int nprobe, float* D,
                    int64_t* I);
int absb_ivf_search_dev(absb_ivf_t h, int64_t n, const float* q_dev, int k, int nprobe,
                        float* D_dev, int64_t* I_dev, void* stream);
/* IndexIVF::search_preassigned(n, x, k, assign, ...): coarse ids [n,nprobe] int64 (-1 = none). */
int absb_ivf_search_preassigned(absb_ivf_t h, int64_t n, const float* q, int k, int nprobe,
                                const int64_t* coarse_ids, float* D, int64_t* I);
int absb_ivf_search_preassigned_dev(absb_ivf_t h, int64_t n, const float* q_dev, int k,
                                    int nprobe, const int64_t* coarse_ids_dev, float* D_dev,
                                    int64_t* I_dev, void* stream);

/* invlists.list_size(l) for all lists -> sizes [nlist] int64 (host). */
int absb_ivf_list_sizes(absb_ivf_t h, int64_t* sizes);
/* invlists.get_codes(l)/get_ids(l): codes [list_size,d] f32 and ids [list_size] i64 in insertion
 * order (host buffers sized from absb_ivf_list_sizes; either may be NULL). */
int absb_ivf_get_list(absb_ivf_t h, int64_t list_no, float* codes, int64_t* ids);

/* Building blocks of Index.train when the training rows are spread over ranks (k-means
 * row): the LOCAL half of Clustering::compute_centroids — per-list fp32 sums of the rows assigned to
 * each list (ascending row order) and member counts, to be all-reduced by the caller — plus the two
 * host-side pieces of faiss's Clustering that need a sequential std::mt19937: rand_perm (subsample
 * and initial centroids) and split_clusters (empty-cluster repair; returns the number of splits). */
int absb_ivf_centroid_sums_dev(absb_ivf_t h, int64_t n, const float* x_dev, const int64_t* list_ids_dev,
                               float* sums_dev /* [nlist,d] */, float* counts_dev /* [nlist] */,
                               void* stream);
int absb_rand_perm(int64_t n, int64_t seed, int32_t* out);
int absb_kmeans_split_clusters(int d, int64_t k, int64_t n, float* hassign, float* centroids,
                               int64_t* nsplit);

/* Sharding by inverted list: this handle keeps only lists l with
 * l % world == rank; add() silently drops rows of other lists but ntotal still counts only kept
 * rows. Must be called before the first add. Coarse quantisation stays replicated. */
int absb_ivf_set_shard(absb_ivf_t h, int rank, int world);
/* Merge per-shard partial results (scores f32 / ids i64 [n,k] per rank, DEVICE) into the final
 * [n,k] with the same (score desc, id asc) order as a single-shard search.  Rank w's arrays start
 * at D_all_dev + w*rank_stride_bytes and I_all_dev + w*rank_stride_bytes — the layout of ONE
 * all-gather of a packed per-rank record {I [n,k] i64, D [n,k] f32}.  rank_stride_bytes == 0 means
 * two dense [world, n, k] arrays. */
int absb_merge_shards_dev(int device, int world, int64_t n, int k, const float* D_all_dev,
                          const int64_t* I_all_dev, int64_t rank_stride_bytes, float* D_dev,
                          int64_t* I_dev, void* stream);

/* Two-stage fine scan (d = 1024): keep an fp16 shadow copy of the list codes (+50% memory) and
 * answer Index.search from half the HBM traffic — a shortlist of `shortlist` (32, 64 or 128, >= k)
 * candidates per query from the fp16 codes, exact fp32 re-score of the shortlist, and a per-query
 * error-bound check that sends every query it cannot prove through the single-pass fp32 scan.
 * Results (ids and scores) are identical to the single-pass scan.  Must be enabled before the first
 * add(); shortlist = 0 returns to the single-pass scan.  absb_ivf_two_stage_fallbacks returns (and
 * clears) the number of queries that took the fallback; it synchronises. */
int absb_ivf_set_two_stage(absb_ivf_t h, int shortlist);
int absb_ivf_two_stage_fallbacks(absb_ivf_t h, int64_t* queries);
/* Tunables (chunk = vectors per scan work item; coarse_impl 0 = fp32 SIMT, 1 = tcgen05
 * split-bf16). Values < 0 leave a setting unchanged. */
int absb_ivf_set_tunables(absb_ivf_t h, int scan_chunk, int coarse_impl, int scan_ctas_per_sm);
/* Fine-scan kernel: impl 1 (default, d = 1024) = list vectors staged in shared memory by cp.async.bulk
 * into per-warp rings of `ring_depth` stages x `ring_stage_vecs` fp32 vectors (the fp16 pass stages twice
 * as many), `ring_warps` warps per CTA (ivf_scan_ring.cu); impl 0 = vectors held in registers
 * (ivf_scan.cu); impl 2 = the ring scan in its co-resident shape: ONE CTA of 8 warps x 2 stages x 4 KB per SM,
 * capped at 96 registers, which fits on an SM next to a GEMM CTA of the encoder when the GEMMs run under
 * absb_gemm_set_smem_budget(161 KB) — the list scan of batch i then overlaps the encode of batch i+1.
 * Results are bit-identical.  -1 / 0 keeps a field. */
int absb_ivf_set_scan_impl(absb_ivf_t h, int impl, int ring_warps, int ring_depth, int ring_stage_vecs);
/* Order of the scan work queue: 1 (default) = list-major — the (query, probe) pairs are sorted by list
 * number, so probes of several queries into one list are scanned at the same time and the repeats hit
 * L2; 0 = query-major.  Results are identical. */
int absb_ivf_set_scan_order(absb_ivf_t h, int list_major);
/* Statistics of the most recent search call on this handle: number of list vectors scanned,
 * algorithmic bytes (vectors * (4d+8)), number of scan work items, number of kernel launches. */
int absb_ivf_last_stats(absb_ivf_t h, int64_t* vectors_scanned, int64_t* bytes_scanned,
                        int64_t* work_items, int64_t* launches);
/* Per-phase device timing with CUDA events recorded on the search's own stream: on = 1 start,
 * 0 stop, 2 start with counters reset.  get_profile synchronises and returns the accumulated
 * milliseconds of the fine-scan kernel, of the coarse GEMM and of everything else (plan, selects,
 * merges) plus the number of fine-scan launches covered — bench.py's roofline.achieved for the
 * scan comes from here (bytes from absb_ivf_last_stats). */
int absb_ivf_set_profile(absb_ivf_t h, int on);
int absb_ivf_get_profile(absb_ivf_t h, double* scan_ms, double* coarse_gemm_ms, double* other_ms,
                         int64_t* scan_launches);
/* Timeline of the kernel spans recorded since absb_*_set_profile(.., 1|2) (before absb_*_get_profile folds them):
 * out[i] = {kind, start_ms, stop_ms} relative to base_event (a cudaEvent_t the caller recorded).  kinds: index 0
 * fine scan, 1 coarse GEMM, 2 other; encoder 0 GEMM, 1 attention, 2 other.  Used to show which kernels of the
 * two QueryPipeline streams run at the same time. */
int absb_ivf_profile_spans(absb_ivf_t h, void* base_event, float* out, int64_t cap, int64_t* n);
int absb_enc_profile_spans(absb_enc_t e, void* base_event, float* out, int64_t cap, int64_t* n);
/* Time and launch count of the fp16 shortlist pass of the two-stage scan alone (it is also part of scan_ms of
 * absb_ivf_get_profile, which adds the fp32 re-score and the fallback launch). */
int absb_ivf_get_profile_scan16(absb_ivf_t h, double* scan16_ms, int64_t* scan16_launches);
/* Replays ONLY the fine-scan kernel of the most recent *_dev search (same work items) `iters`
 * times on `stream` and returns the mean duration in ms measured with CUDA events on that
 * stream — used by bench.py for roofline.achieved. */
int absb_ivf_time_scan(absb_ivf_t h, int iters, void* stream, float* ms_mean);

/* ---------------------------------------------------------------- encoder ------------------ */
/* stella_en_1.5B_v5 = Qwen2-1.5B backbone (bidirectional) -> mean pool -> Dense 1536->1024 ->
 * optional L2 normalise; the arithmetic behind SentenceTransformer.encode(). */
typedef struct absb_enc_config {
  int32_t vocab_size;        /* 151646 */
  int32_t hidden_size;       /* 1536   */
  int32_t num_layers;        /* 28     */
  int32_t num_heads;         /* 12     */
  int32_t num_kv_heads;      /* 2      */
  int32_t head_dim;          /* 128    */
  int32_t intermediate_size; /* 8960   */
  int32_t embed_dim;         /* 1024 (Dense out_features) */
  int32_t max_seq_len;       /* 512    */
  int32_t causal;            /* 0 = bidirectional (stella), 1 = stock Qwen2 causal */
  float rms_eps;             /* 1e-6   */
  float rope_theta;          /* 1e6    */
} absb_enc_config;

int absb_enc_create(const absb_enc_config* cfg, int device, absb_enc_t* out);
int absb_enc_destroy(absb_enc_t e);
/* Load one parameter by its Hugging Face name ("embed_tokens.weight",
 * "layers.3.self_attn.q_proj.weight", "norm.weight", "dense.weight", "dense.bias", ...).
 * data is a HOST pointer; dtype 0 = float32, 1 = bfloat16 (raw uint16). Weights are stored bf16
 * on the device (norm weights and biases fp32). */
int absb_enc_load_weight(absb_enc_t e, const char* name, const void* data, int dtype,
                         const int64_t* shape, int ndim);
/* Initialise all weights on the device from the counter-based generator (normal(0, std)) —
 * the random-init model of the true architecture used offline (no checkpoint available). */
int absb_enc_init_random(absb_enc_t e, uint64_t seed, float std);
/* Read a parameter back (fp32) for the oracle to load the very same weights. */
int absb_enc_get_weight(absb_enc_t e, const char* name, float* out, int64_t numel);
/* Forward: input_ids [B,S] int64, attention_mask [B,S] int32 (1 = token, 0 = pad; right padding)
 * -> out [B, embed_dim] float32 (L2-normalised when normalize != 0). Host pointers. */
int absb_enc_forward(absb_enc_t e, int B, int S, const int64_t* input_ids,
                     const int32_t* attention_mask, int normalize, float* out);
int absb_enc_forward_dev(absb_enc_t e, int B, int S, const int64_t* input_ids_dev,
                         const int32_t* attention_mask_dev, int normalize, float* out_dev,
                         void* stream);
/* Debug / parity taps: last_hidden_state [B,S,hidden] fp32 of the most recent forward. */
int absb_enc_last_hidden(absb_enc_t e, float* out, int64_t numel);
/* FLOP count and kernel launches of the most recent forward. */
int absb_enc_last_stats(absb_enc_t e, double* flops, int64_t* launches);
/* Attention kernel: 1 (default) = tcgen05 for every sequence length up to the model's 512 tokens (persistent
 * warp-specialised kernel up to 256 tokens, two-key-block kernel for 257-512); 2 = the one-tile-per-CTA
 * tcgen05 kernel up to 256 tokens; 0 = always the mma.sync kernel (test hooks). */
int absb_enc_set_attention_impl(absb_enc_t e, int impl);
/* Per-phase device timing with CUDA events recorded on the forward's own stream around every
 * kernel: on = 1 start, 0 stop, 2 start with counters reset.  get_profile synchronises and returns
 * the accumulated milliseconds of the tcgen05 GEMM launches (with their FLOPs), of the attention
 * kernel and of everything else, plus the number of forwards covered — bench.py's roofline.achieved
 * for the GEMM comes from here. */
int absb_enc_set_profile(absb_enc_t e, int on);
int absb_enc_get_profile(absb_enc_t e, double* gemm_ms, double* gemm_flops, double* attention_ms,
                         double* other_ms, int64_t* forwards);

/* Stand-alone GEMM entry used by tests and the micro-benchmark: C[M,N] (fp32) = A[M,K] * B[N,K]^T
 * with bf16 operands on tcgen05 (DEVICE pointers; K % 64 == 0). */
/* Shared memory one GEMM CTA may take (bytes; 0 = all 227 KB): the operand ring gets as many stages as fit.
 * 161 KB (4-5 stages instead of 5-7) leaves room for one co-resident scan CTA (absb_ivf_set_scan_impl 2). */
int absb_gemm_set_smem_budget(int bytes);
/* Test hook: force the tile shape of the tcgen05 GEMM (0 = automatic; 1 = one CTA, 128x256 tiles;
 * 2 = CTA pair (cta_group::2), 256x256 tiles; 3 = CTA pair, 256x192 tiles; 4 / 5 = "quad": a cluster of two CTA
 * pairs stacked in M that share the B tile through TMA multicast, 512x256 / 512x192 super tiles). Process-wide. */
int absb_gemm_set_variant(int variant);
/* Split K of the residual-add GEMMs (epilogue 2: O-proj, FFN-down): 0 / 1 = never (default: measured slower on
 * B200 at every encoder shape), n = force n slices (test / micro-benchmark hook).  The slices
 * of a tile add into the output in slice order: the result does not depend on timing.  Process-wide. */
int absb_gemm_set_ksplit(int slices);
/* Same GEMM with one of the fused epilogues (gemm_tc.cuh): 0 = bf16 out (+bias), 1 = f32 out (+bias),
 * 2 = f32 out += acc (residual stream), 3 = bf16 SwiGLU (B rows interleaved per 256-row tile, out
 * [M, N/2]).  Test / micro-benchmark hook. */
int absb_gemm_bf16_epi_dev(int device, int epi, int M, int N, int K, const void* A_dev, const void* B_dev,
                           void* out_dev, int64_t ldc, const float* bias_dev, void* stream);
int absb_gemm_bf16_dev(int device, int M, int N, int K, const void* A_dev, const void* B_dev,
                       float* C_dev, void* stream);

/* ---------------------------------------------------------------- NVLink peer exchange ------ */
/* The exchange step of the list-sharded search without NCCL: every rank of one
 * box owns a small device buffer — a two-deep ring of [world] records plus [world] arrival flags —
 * and maps every other rank's buffer (CUDA IPC over NVLink/NVSwitch).  An all-gather is remote
 * stores by the producing kernel into slot [rank] of every rank's buffer, a system-scope release
 * of flag [rank] everywhere, and an acquire of the local flags by the consuming kernel.  This
 * replaces faiss IndexShards' host-side merge for /root/reference/README.md:16,28 (app.py query
 * loop) at 8 GPUs.  All calls of one exchange must use one stream; every rank must issue the same
 * sequence of exchanges.
 *   create:    slot_bytes = capacity of one rank's record.  Collective by convention: every rank
 *              creates its exchange, publishes absb_peer_ipc_handle (64 bytes) to all ranks (any
 *              host channel), then calls absb_peer_connect with the [world][64] handle table.
 *   connect_ptrs: same-process variant taking raw device pointers (absb_peer_local_ptr of the
 *              other exchanges) — several ranks emulated on one GPU, for tests.
 *   allgather: push `bytes` (multiple of 16) from src_dev, wait for all ranks, return the local
 *              [world][slot_bytes] entry of this epoch (valid until the next-but-one exchange).
 *              push / wait are its two halves.
 *   status:    1 after a wait gave up on a dead peer (20 s), else 0.  Synchronises the stream. */
int absb_peer_create(int device, int rank, int world, size_t slot_bytes, absb_peer_t* out);
int absb_peer_destroy(absb_peer_t p);
int absb_peer_ipc_handle(absb_peer_t p, void* handle64);
int absb_peer_local_ptr(absb_peer_t p, void** ptr_dev);
int absb_peer_connect(absb_peer_t p, const void* handles /* [world][64] */);
int absb_peer_connect_ptrs(absb_peer_t p, void* const* ptrs_dev /* [world] */);
int absb_peer_allgather_dev(absb_peer_t p, const void* src_dev, size_t bytes, void** gathered_dev,
                            void* stream);
int absb_peer_push_dev(absb_peer_t p, const void* src_dev, size_t bytes, void* stream);
int absb_peer_wait_dev(absb_peer_t p, void** gathered_dev, void* stream);
int absb_peer_status(absb_peer_t p, int* status);
/* Index.search fused with the exchange: the kernel that merges this shard's partial k-best lists
 * stores the result straight into every rank's buffer and raises the flags (no local D, I, no
 * separate copy or collective).  Record = I [n,k] i64 then D [n,k] f32 (16-byte aligned). */
int absb_ivf_search_push_dev(absb_ivf_t h, absb_peer_t p, int64_t n, const float* q_dev, int k,
                             int nprobe, void* stream);
/* absb_ivf_search_push_dev with the coarse result given (IndexIVF::search_preassigned): for query batches whose
 * coarse top-nprobe was computed where the query was encoded and travelled with the embedding in the all-gather,
 * instead of being recomputed for the whole batch on every rank. */
int absb_ivf_search_preassigned_push_dev(absb_ivf_t h, absb_peer_t p, int64_t n, const float* q_dev, int k, int nprobe,
                                         const int64_t* coarse_ids_dev, void* stream);
/* Same exchange for a result that already sits in local (D, I) [n,k] (e.g. from the two-stage scan):
 * packs the record and pushes it to every rank; absb_peer_merge_shards_dev consumes it. */
int absb_peer_push_results_dev(absb_peer_t p, int64_t n, int k, const float* D_dev, const int64_t* I_dev,
                               void* stream);
/* The consuming half: waits (inside the kernel) for every rank's record of the last
 * absb_ivf_search_push_dev and merges world x k candidates per query with the single-index order
 * (score desc, id asc) -> D_dev [n,k], I_dev [n,k] on this rank. */
int absb_peer_merge_shards_dev(absb_peer_t p, int64_t n, int k, float* D_dev, int64_t* I_dev,
                               void* stream);

/* ---------------------------------------------------------------- OpenAlex front end -------- */
/* The stage in front of bulk encode.  Replaces the reference's `./oa_jsonl`
 * executable — /root/reference/Makefile:64, main loop /root/reference/oa_jsonl.c:351-414 — which
 * turns OpenAlex `works` JSON lines into {"id","document"} JSON lines (English records with a
 * non-empty abstract_inverted_index; document = title + ' ' + un-inverted abstract, strings left
 * JSON-escaped).  HOST code only, no GPU needed; bytes out are identical to the reference
 * program's on well-formed input.
 *   in[0:in_len)  one block of the stream.  final_chunk = 0: only its complete lines are
 *                 converted and *consumed tells where the unfinished tail starts; final_chunk = 1:
 *                 a last line without '\n' is converted too (oa_jsonl.c:333-349).
 *   threads       line ranges converted concurrently (0 = all hardware threads).
 *   *out          malloc'ed result of *out_len bytes (+ a NUL), released with absb_oa_jsonl_free.
 *   stats[4]      optional: lines read, records written, records dropped, 1 if an empty line
 *                 ended the conversion (oa_jsonl.c:363-366; *consumed = in_len then).
 * A malformed record returns ABSB_ERR_INVALID naming the line (the reference assert()s). */
int absb_oa_jsonl_convert(const char* in, size_t in_len, int final_chunk, int threads, char** out,
                          size_t* out_len, size_t* consumed, int64_t* stats);
int absb_oa_jsonl_free(char* out);

#ifdef __cplusplus
}
#endif
#endif /* ABSB200_H */
