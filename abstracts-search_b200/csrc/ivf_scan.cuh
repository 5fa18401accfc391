// ivf_scan.cuh — data layout shared by the inverted-list storage (ivf.cu) and the scan kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace absb {

constexpr int kMaxPlanQueries = 1024;  // queries per plan/scan launch (host splits larger batches)
constexpr int kScanThreads = 128;      // 4 warps per CTA, every warp an independent worker
constexpr int kScanUnroll = 2;         // list vectors in flight per warp (2 x 4 KB)

// One unit of scan work: `len` physically contiguous vectors of one list for query `q`.
struct __align__(16) ScanItem {
  const float* codes;    // [len, d]
  const long long* ids;  // [len]
  int len;
  int q;
  long long pad;  // global slot number (page * page_vecs) of the first vector: locates the fp16 shadow codes
};
static_assert(sizeof(ScanItem) == 32, "ScanItem must be 32 bytes");

// Device view of the paged inverted lists.
//   page g lives in slab g >> slab_shift at slot g & (2^slab_shift - 1);
//   a page holds page_vecs vectors ([page_vecs, d] f32) and page_vecs ids;
//   list l owns pages pt_pages[pt_off[l] .. pt_off[l] + ceil(list_size[l] / page_vecs)) in order.
struct ListTable {
  int nlist;
  int d;
  int page_vecs;
  int slab_shift;
  const long long* list_size;  // [nlist]
  const long long* pt_off;     // [nlist + 1]
  const int* pt_pages;         // [pt_off[nlist]]
  float* const* code_slabs;    // device array of slab base pointers
  long long* const* id_slabs;
  unsigned short* const* half_slabs;  // fp16 shadow codes of the two-stage scan (nullptr = none)
};

struct ScanLaunch {
  const float* Q;  // [nq, d]
  int d;
  int k;
  const ScanItem* items;
  const int* n_items;
  int* queue_counter;
  const int* order;    // queue position -> item index (list-major order), or nullptr = identity
  float* part_s;       // [max_items, k]
  long long* part_id;  // [max_items, k]
  int sm_count;
  int ctas_per_sm;  // <= 0: occupancy query
};

// Geometry of the shared-memory ring scan (ivf_scan_ring.cu): every warp of a CTA owns `depth` stages of
// `stage_vecs` vectors (fp32: 1 or 2 = 4 / 8 KB; fp16 shadow codes: 2 or 4 = 4 / 8 KB) filled by cp.async.bulk.
struct ScanRing {
  int warps = 4;
  int depth = 4;
  int stage_vecs = 2;  // fp32 vectors per stage; the fp16 pass stages twice as many (same bytes)
  bool l2_evict_first = false;  // stream the list codes through L2 with evict-first priority
  bool small = false;  // co-resident variant: 4 KB stages, <= 96 registers (fits beside a GEMM CTA of the encoder)
};
void launch_scan_ring(const ScanLaunch& a, const ScanRing& r, cudaStream_t st);

// Scratch of the list-major queue order (all [npairs + 1] except order [max_items]).
struct PlanOrderWs {
  unsigned *keys, *keys_sorted;
  int *vals, *vals_sorted, *counts_sorted, *qoffs, *order;
};
// active (optional, [nq] bytes): queries with active[q] == 0 get no work items.
void launch_plan(const ListTable& lt, const long long* coarse, int nq, int nprobe, int chunk,
                 int max_items, ScanItem* items, int* q_begin, int* n_items, int* queue_counter,
                 unsigned long long* stats, int* pair_counts, int* pair_offs, void* scan_tmp,
                 size_t scan_tmp_bytes, const PlanOrderWs* order_ws, cudaStream_t st,
                 const unsigned char* active = nullptr);
size_t plan_scan_tmp_bytes(int max_pairs);
// out[l] = number of work items the plan emits for list l with this chunk length
void launch_count_list_items(const ListTable& lt, int chunk, int* out, cudaStream_t st);
void launch_scan(const ScanLaunch& a, cudaStream_t st);

// ---- two-stage scan (ivf_scan16.cu): fp16 shortlist pass + exact fp32 re-score, d = 1024 ----
struct Scan16Launch {
  const float* Q;  // [nq, 1024]
  int K;           // shortlist length per item and per query (32, 64 or 128)
  const ScanItem* items;
  const int* n_items;
  int* queue_counter;
  const int* order;
  float* part_s;      // [max_items, K] approximate scores
  long long* part_g;  // [max_items, K] global slot numbers
  const unsigned short* const* half_slabs;
  int slab_shift, page_vecs;
  int sm_count, ctas_per_sm;
};
void launch_scan16(const Scan16Launch& a, cudaStream_t st);
void launch_scan16_ring(const Scan16Launch& a, const ScanRing& r, cudaStream_t st);
// shortlist G [nq, K] (global slot numbers, -1 = none) -> one single-vector fp32 work item each
void launch_rescore_items(const ListTable& lt, int nq, int K, const long long* G, ScanItem* items, int* q_begin,
                          int* n_items, int* queue_counter, cudaStream_t st);
// flags[q] = 1 where the error bound cannot prove that the shortlist holds the exact top-k
void launch_two_stage_check(int nq, int d, int k, int K, const float* Q, const float* D, const long long* I,
                            const float* Dp, const long long* G, const float* maxima, unsigned char* flags,
                            int* n_flagged, cudaStream_t st);

// dense.cu
void gemm_nt_f32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                 int ldc, cudaStream_t st);
void select_rows(const float* S, int64_t ld, int64_t nrows, int ncols, long long id_offset, int k,
                 float* out_s, long long* out_id, int64_t out_ld, bool finalize, cudaStream_t st);
// active (optional, [nq] bytes): rows of queries with active[q] == 0 are left untouched.
void merge_partials(int nq, int k, const int* q_begin, const float* part_s, const long long* part_id,
                    float* D, long long* I, cudaStream_t st, const unsigned char* active = nullptr);
void merge_shards(int world, int64_t nq, int k, const float* D_all, const long long* I_all,
                  int64_t d_stride_bytes, int64_t i_stride_bytes, float* D, long long* I, cudaStream_t st);

// synth.cu
void synth_fill(int kind, uint64_t seed, int64_t row0, const long long* row_ids, int64_t n, int d,
                int nlist, int64_t corpus_rows, float* out, cudaStream_t st);
void synth_cluster(uint64_t seed, int64_t row0, int64_t n, int nlist, long long* out, cudaStream_t st);

}  // namespace absb
