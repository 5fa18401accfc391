// peer.cuh — NVLink peer-memory exchange for the sharded search.
//
// Every rank owns one device buffer laid out as a two-deep ring of [world] records plus [world]
// arrival flags, and maps the buffers of all other ranks of the box into its address space (CUDA
// IPC; NVSwitch gives every pair full bandwidth).  An all-gather is then plain remote stores: the
// producing kernel writes its record into slot [rank] of EVERY rank's ring entry and, when its
// last CTA is done, raises flag [rank] on every rank with a system-scope release; the consuming
// kernel acquires the `world` flags of its own buffer and reads locally.  No NCCL call, no
// separate copy, no host involvement on the query path.
//
// Ring safety (two entries suffice): rank A pushes epoch e+2 only after its own wait(e+1) has
// completed in stream order, which needed rank B's push(e+1), which B issued after the kernels
// that read entry e & 1 on B.  All exchange traffic of one PeerExchange must stay on one stream.
#pragma once

#include <vector>

#include "common.cuh"

namespace absb {

// By-value kernel argument: where the merged top-k of a search launch goes.
struct PeerPush {
  int world = 0;                           // 0 = plain local output
  char* const* slot_ptrs = nullptr;        // [world] base of this rank's record on every rank (this epoch)
  unsigned long long* const* flag_ptrs = nullptr;  // [world] &flags[rank] on every rank
  unsigned long long epoch = 0;            // value to raise when the launch completes (0 = not yet)
  unsigned* done_counter = nullptr;        // local: CTAs finished in this launch
  long long i_off = 0, d_off = 0;          // byte offsets of I [nq_total, k] i64 and D [nq_total, k] f32 in the record
  long long q_off = 0;                     // first query of this launch within the record
};

struct PeerExchange {
  int device, rank, world;
  size_t slot_bytes;                   // capacity of one rank's record
  size_t data_bytes;                   // 2 * world * slot_bytes rounded up
  char* local = nullptr;               // [2][world][slot_bytes] | flags [world] u64
  std::vector<char*> base;             // [world] mapped buffers (base[rank] == local)
  std::vector<bool> opened;            // mapped through cudaIpcOpenMemHandle (to be closed)
  bool connected = false;
  DBuf<char*> d_slot_ptrs;             // [2][world]
  DBuf<unsigned long long*> d_flag_ptrs;  // [world]
  DBuf<unsigned> done_counter;
  // 0 ok, 1 = a wait timed out.  Lives in mapped pinned host memory: the waiting kernel writes it
  // through the device alias, the host polls it after every search WITHOUT a synchronisation.
  int* status_host = nullptr;
  int* status_dev = nullptr;
  DBuf<char> staging;                  // packed {I, D} record of absb_peer_push_results_dev
  unsigned long long epoch = 0;        // last epoch pushed
  int64_t rec_n = 0;                   // shape of the last pushed search record
  int rec_k = 0;

  PeerExchange(int device, int rank, int world, size_t slot_bytes);
  ~PeerExchange();
  void ipc_handle(void* blob64) const;
  void connect_ipc(const void* blobs);  // [world][64]
  void connect_ptrs(void* const* ptrs);  // same-process peers (tests): raw device pointers
  void finish_connect();
  unsigned long long* local_flags() const { return reinterpret_cast<unsigned long long*>(local + data_bytes); }
  char* local_entry(unsigned long long e) const { return local + (e & 1) * world * slot_bytes; }

  // next epoch's push descriptor (does not advance the epoch; commit() does)
  PeerPush begin_push(long long i_off, long long d_off);
  void commit() { ++epoch; }
  // generic all-gather: bytes from src into slot [rank] everywhere, then wait for everyone
  void push(const void* src, size_t bytes, cudaStream_t st);
  char* allgather(const void* src, size_t bytes, cudaStream_t st);
  void wait(cudaStream_t st);
  int read_status() const { return *const_cast<volatile int*>(status_host); }
};

// dense.cu
// active (optional, [nq] bytes): only queries with active[q] != 0 are merged from the partials; the
// others push row q of (Dbase, Ibase) as it is.
void merge_partials_push(int nq, int k, const int* q_begin, const float* part_s, const long long* part_id,
                         const PeerPush& pp, cudaStream_t st, const unsigned char* active = nullptr,
                         const float* Dbase = nullptr, const long long* Ibase = nullptr);
void merge_shards_wait(int world, int64_t nq, int k, const char* entry, int64_t slot_bytes, int64_t i_off,
                       int64_t d_off, const unsigned long long* flags, unsigned long long epoch, int* status,
                       float* D, long long* I, cudaStream_t st);

}  // namespace absb
