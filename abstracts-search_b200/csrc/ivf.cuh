// ivf.cuh — host-side state of the B200 IndexIVFFlat / IndexFlatIP replacements.
#pragma once

#include <memory>
#include <vector>

#include "common.cuh"
#include "ivf_scan.cuh"
#include "peer.cuh"

namespace absb {

// Slab pool of list pages.  Page g -> slab g >> shift.  Slabs are plain cudaMalloc blocks that are
// never moved, so growing the index never copies list data (180 GB HBM: no room for a 2x copy of a
// 106 GB shard).
struct PagePool {
  int d = 0;
  int page_vecs = 0;
  int slab_shift = 0;  // pages per slab = 1 << slab_shift
  int64_t pages_used = 0;
  std::vector<float*> code_slabs;
  std::vector<long long*> id_slabs;
  std::vector<unsigned short*> half_slabs;  // fp16 shadow codes (two-stage scan), same page numbering
  DBuf<float*> d_code_slabs;
  DBuf<long long*> d_id_slabs;
  DBuf<unsigned short*> d_half_slabs;
  bool shadow = false;  // keep fp16 shadow codes for pages allocated from now on
  size_t table_cap = 0;

  void configure(int d_, int page_vecs_);
  void ensure_pages(int64_t total_pages, cudaStream_t st);
  void release();
  ~PagePool() { release(); }
};

struct ClusteringParams {
  int niter = 10;
  int max_points_per_centroid = 256;
  int min_points_per_centroid = 39;
  int64_t seed = 1234;
  bool spherical = false;  // faiss ClusteringParameters::spherical (index_factory sets it for inner product)
};

// Where a search sends its merged top-k when it feeds the peer exchange instead of a local (D, I).
struct SearchPush {
  PeerPush pp;       // this epoch's descriptor (PeerExchange::begin_push)
  int64_t q_base;    // first query of the current call within the record
  int64_t nq_total;  // queries of the whole record: the launch that completes it raises the flags
};

struct SearchStats {
  int64_t vectors = 0, bytes = 0, items = 0, launches = 0;
};

struct IvfIndex {
  int d, nlist, device;
  DeviceProps props;
  cudaStream_t own_stream = nullptr;
  bool trained = false;
  ClusteringParams cp;
  int shard_rank = 0, shard_world = 1;

  DBuf<float> centroids;  // [nlist, d]
  DBuf<uint16_t> centroids3;  // bf16 [nlist, 3d] = [hi | mid | lo] split for the tcgen05 coarse GEMM
  DBuf<uint16_t> ws_q3;
  DBuf<float> ws_amax;  // fused arg-max partials of assign_dev
  DBuf<int> ws_aidx;
  bool c3_dirty = true;

  // inverted lists
  PagePool pool;
  DBuf<long long> list_size;  // [nlist]
  DBuf<long long> pt_off;     // [nlist+1]
  DBuf<int> pt_pages;
  int64_t pt_total_pages = 0;
  std::vector<int64_t> h_list_size;
  // prefix sums of the per-list work-item counts (what the plan emits under scan_chunk), sorted
  // descending: entry [p] bounds the items of any p probes.  Rebuilt lazily (one small D2H copy) on
  // the first search after add / compact / reset or a change of scan_chunk.
  std::vector<int64_t> h_items_prefix_desc;
  int items_bound_chunk = -1;  // scan_chunk the table was built for; -1 = stale
  DBuf<int> ws_list_items;
  int64_t ntotal = 0;
  int64_t rows_seen = 0;  // rows offered to add() so far, kept or not: the next default id

  // tunables
  int scan_chunk = 128;
  int coarse_impl = 1;  // 1 = tcgen05 split-bf16 (falls back to the FFMA GEMM for shapes it cannot take)
  int scan_ctas_per_sm = 0;
  int two_stage_k = 0;  // shortlist length of the two-stage scan (0 = single-pass fp32 scan)
  int scan_impl = 1;  // 1 = shared-memory ring scan (cp.async.bulk staging, ivf_scan_ring.cu; d = 1024), 2 = its small
                      // co-resident variant (one 8-warp CTA per SM beside the encoder's GEMM CTAs), 0 = register scan
  ScanRing ring;      // geometry of the ring scan
  int scan_order = 1;  // 1 = list-major work queue (probes of one list scanned together: L2 reuse), 0 = query-major

  // workspaces (single stream at a time)
  DBuf<float> ws_scores;
  DBuf<float> ws_coarse_s;
  DBuf<long long> ws_coarse_i;
  DBuf<ScanItem> ws_items;
  DBuf<float> ws_part_s;
  DBuf<long long> ws_part_id;
  DBuf<int> ws_q_begin;
  DBuf<int> ws_pair_counts, ws_pair_offs;  // plan: items per (query, probe) pair and their prefix sums
  DBuf<unsigned char> ws_plan_tmp;
  DBuf<unsigned> ws_okeys, ws_okeys_sorted;  // list-major queue order (scan_order = 1)
  DBuf<int> ws_ovals, ws_ovals_sorted, ws_ocounts, ws_oqoffs, ws_order;
  // two-stage scan (ivf_scan16.cu)
  DBuf<float> ws_maxima;  // [0] max |x - fp16(x)|, [1] max |x| over everything added
  DBuf<float> ws_short_s, ws_part2_s;
  DBuf<long long> ws_short_g, ws_part2_id;
  DBuf<ScanItem> ws_items2;
  DBuf<int> ws_q_begin2, ws_counters2;
  DBuf<unsigned char> ws_flags;
  DBuf<int> ws_nflag;  // queries sent to the single-pass fallback since the last reset of the counter
  DBuf<unsigned long long> ws_stats2;
  DBuf<int> ws_counters;  // [0] n_items, [1] queue counter
  DBuf<unsigned long long> ws_stats;
  DBuf<unsigned char> ws_cub;
  DBuf<float> ws_x;  // staged host rows / queries
  DBuf<long long> ws_ids;
  DBuf<long long> ws_list_ids;
  DBuf<float> ws_D;
  DBuf<long long> ws_I;
  DBuf<float> ws_push_D;  // local (D, I) of a two-stage search whose final merge pushes to the peers
  DBuf<long long> ws_push_I;

  // replay info for absb_ivf_time_scan
  ScanLaunch last_scan{};
  bool have_last_scan = false;
  bool use_ring() const { return scan_impl >= 1 && d == 1024; }
  void run_scan(const ScanLaunch& a, cudaStream_t st) {
    if (use_ring()) launch_scan_ring(a, ring, st);
    else launch_scan(a, st);
  }
  SearchStats stats;
  bool stats_pending = false;  // ws_stats holds device-side numbers not yet folded into `stats`

  // optional per-phase device timing (absb_ivf_set_profile): CUDA events recorded on the search's
  // own stream around the fine-scan kernel (0), the coarse GEMM (1) and everything else (2)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  size_t ev_used = 0;
  std::vector<int> ev_kind;
  double prof_ms[4] = {0, 0, 0, 0};  // [3] = the fp16 shortlist pass of the two-stage scan alone (also counted in [0])
  int64_t prof_scan16_launches = 0;
  int64_t prof_scan_launches = 0;
  int64_t prof_vectors = 0;
  struct Span {
    IvfIndex* ix;
    cudaStream_t st;
    bool on;
    cudaEvent_t stop = nullptr;
    Span(IvfIndex* ix_, cudaStream_t st_, int kind);
    ~Span() {
      if (on) cudaEventRecord(stop, st);
    }
  };
  void fold_profile();

  IvfIndex(int d, int nlist, int device);
  ~IvfIndex();

  ListTable table() const;
  void reset_stats() {
    stats = SearchStats{};
    stats_pending = false;
  }

  void reset();
  void set_centroids_dev(const float* c, cudaStream_t st);
  void train_dev(int64_t n, const float* x, cudaStream_t st);
  void train_host(int64_t n, const float* x);
  void centroid_sums_dev(int64_t n, const float* x, const long long* assign, float* sums, float* counts,
                         cudaStream_t st);
  // scores -> top-k coarse (k = nprobe) for nq <= chunk rows; Ic int64 [nq, nprobe]
  void coarse_dev(int64_t nq, const float* q, int nprobe, float* Dc, long long* Ic, bool finalize,
                  cudaStream_t st);
  void assign_dev(int64_t n, const float* x, long long* list_ids, cudaStream_t st);
  void add_core_dev(int64_t n, const float* x, const long long* ids, const long long* list_ids,
                    cudaStream_t st);
  void add_dev(int64_t n, const float* x, const long long* ids, cudaStream_t st);
  void search_preassigned_dev(int64_t nq, const float* q, int k, int nprobe, const long long* coarse,
                              float* D, long long* I, cudaStream_t st, const SearchPush* push = nullptr);
  void search_dev(int64_t nq, const float* q, int k, int nprobe, float* D, long long* I,
                  cudaStream_t st, const SearchPush* push = nullptr);
  void fold_stats();
  void coarse_scores(int M, const float* q, float* S, cudaStream_t st);
  void get_list(int64_t list_no, float* codes, long long* ids);
  // Physically reorders the pages so that every list's pages are consecutive (in place, through a
  // bounded scratch of `scratch_pages` pages; <= 0 picks a size from the free memory).
  void compact(int64_t scratch_pages, cudaStream_t st);
  void set_two_stage(int shortlist);
  int64_t two_stage_fallbacks(cudaStream_t st);
  int64_t items_bound_per_query(int nprobe, cudaStream_t st);
  void refresh_host_sizes(cudaStream_t st);
};

// One page copy of the in-place compaction: locations >= 0 are pool pages, < 0 scratch slot -1 - v.
struct PageMove {
  int from, to;
};
// Plans content_new[t] = content_old[src[t]] (src a permutation of [0, n)) as phases of mutually
// independent page copies that run in order, using at most `scratch_pages` scratch pages.
void plan_page_compaction(std::vector<int> src, int64_t scratch_pages, std::vector<PageMove>& moves,
                          std::vector<int64_t>& phase_end);

void rand_perm_export(int64_t n, int64_t seed, int* out);
void renorm_rows(int64_t n, int d, float* x, cudaStream_t st);  // fvec_renorm_L2
int64_t split_clusters_export(int d, int64_t k, int64_t n, float* hassign, float* centroids);

struct FlatIndex {
  int d, device;
  DeviceProps props;
  cudaStream_t own_stream = nullptr;
  DBuf<float> xb;
  int64_t ntotal = 0;
  DBuf<float> ws_scores, ws_part_s, ws_x, ws_D;
  DBuf<long long> ws_part_id, ws_I;
  DBuf<int> ws_q_begin;
  DBuf<uint16_t> ws_q3, ws_x3;  // bf16 [rows, 3d] splits of the query block / the database chunk (tcgen05 path)

  FlatIndex(int d, int device);
  ~FlatIndex();
  void add_dev(int64_t n, const float* x, cudaStream_t st);
  void search_dev(int64_t nq, const float* q, int k, float* D, long long* I, cudaStream_t st);
};

}  // namespace absb
