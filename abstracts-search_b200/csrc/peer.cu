// peer.cu — the peer-memory exchange behind ShardedIndexIVFFlat.search (see peer.cuh).
#include "peer.cuh"

namespace absb {
namespace {

constexpr unsigned long long kWaitTimeoutNs = 20ull * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// grid = (chunks, world): CTA (c, w) copies 16-byte words of the record to rank w; the last CTA to
// finish raises this rank's flag everywhere.
__global__ void peer_push_kernel(const uint4* __restrict__ src, size_t n16, int world, char* const* __restrict__ slot_ptrs,
                                 unsigned long long* const* __restrict__ flag_ptrs, unsigned long long epoch,
                                 unsigned* __restrict__ done_counter) {
  uint4* dst = reinterpret_cast<uint4*>(slot_ptrs[blockIdx.y]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    if (atomicAdd(done_counter, 1u) == total - 1) {
      *done_counter = 0;
      __threadfence_system();
      for (int w = 0; w < world; ++w) st_release_sys(flag_ptrs[w], epoch);
    }
  }
}

__global__ void peer_wait_kernel(int world, const unsigned long long* __restrict__ flags, unsigned long long epoch,
                                 int* __restrict__ status) {
  if ((int)threadIdx.x < world) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
      if (globaltimer_ns() - t0 > kWaitTimeoutNs) {
        *status = 1;
        break;
      }
      __nanosleep(64);
    }
  }
}

}  // namespace

PeerExchange::PeerExchange(int device_, int rank_, int world_, size_t slot_bytes_)
    : device(device_), rank(rank_), world(world_), slot_bytes((slot_bytes_ + 255) & ~(size_t)255) {
  ABSB_CHECK(world >= 1 && world <= 64 && rank >= 0 && rank < world, ABSB_ERR_INVALID, "rank %d of %d", rank, world);
  ABSB_CHECK(slot_bytes_ > 0, ABSB_ERR_INVALID, "empty record");
  data_bytes = 2 * (size_t)world * slot_bytes;
  ABSB_CUDA(cudaMalloc(&local, data_bytes + sizeof(unsigned long long) * world));
  ABSB_CUDA(cudaMemset(local, 0, data_bytes + sizeof(unsigned long long) * world));
  done_counter.alloc_exact(1);
  ABSB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&status_host), sizeof(int), cudaHostAllocMapped));
  *status_host = 0;
  ABSB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&status_dev), status_host, 0));
  ABSB_CUDA(cudaMemset(done_counter.p, 0, sizeof(unsigned)));
  ABSB_CUDA(cudaDeviceSynchronize());  // zeroed before anybody can learn the handle
  base.assign(world, nullptr);
  opened.assign(world, false);
  base[rank] = local;
}

PeerExchange::~PeerExchange() {
  cudaSetDevice(device);
  cudaDeviceSynchronize();
  for (int w = 0; w < world; ++w)
    if (opened[w] && base[w]) cudaIpcCloseMemHandle(base[w]);
  if (local) cudaFree(local);
  if (status_host) cudaFreeHost(status_host);
}

void PeerExchange::ipc_handle(void* blob64) const {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  ABSB_CUDA(cudaIpcGetMemHandle(&h, local));
  memcpy(blob64, &h, sizeof(h));
}

void PeerExchange::connect_ipc(const void* blobs) {
  ABSB_CHECK(!connected, ABSB_ERR_STATE, "already connected");
  for (int w = 0; w < world; ++w) {
    if (w == rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(blobs) + 64 * (size_t)w, sizeof(h));
    void* p = nullptr;
    ABSB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    base[w] = static_cast<char*>(p);
    opened[w] = true;
  }
  finish_connect();
}

void PeerExchange::connect_ptrs(void* const* ptrs) {
  ABSB_CHECK(!connected, ABSB_ERR_STATE, "already connected");
  for (int w = 0; w < world; ++w)
    if (w != rank) base[w] = static_cast<char*>(ptrs[w]);
  finish_connect();
}

void PeerExchange::finish_connect() {
  std::vector<char*> slots(2 * (size_t)world);
  std::vector<unsigned long long*> flags((size_t)world);
  for (int w = 0; w < world; ++w) {
    ABSB_CHECK(base[w] != nullptr, ABSB_ERR_STATE, "peer %d not mapped", w);
    for (int e = 0; e < 2; ++e) slots[(size_t)e * world + w] = base[w] + ((size_t)e * world + rank) * slot_bytes;
    flags[w] = reinterpret_cast<unsigned long long*>(base[w] + data_bytes) + rank;
  }
  d_slot_ptrs.alloc_exact(slots.size());
  d_flag_ptrs.alloc_exact(flags.size());
  ABSB_CUDA(cudaMemcpy(d_slot_ptrs.p, slots.data(), sizeof(char*) * slots.size(), cudaMemcpyHostToDevice));
  ABSB_CUDA(cudaMemcpy(d_flag_ptrs.p, flags.data(), sizeof(void*) * flags.size(), cudaMemcpyHostToDevice));
  connected = true;
}

PeerPush PeerExchange::begin_push(long long i_off, long long d_off) {
  ABSB_CHECK(connected, ABSB_ERR_STATE, "peer exchange is not connected");
  PeerPush pp;
  const unsigned long long e = epoch + 1;
  pp.world = world;
  pp.slot_ptrs = d_slot_ptrs.p + (e & 1) * world;
  pp.flag_ptrs = d_flag_ptrs.p;
  pp.epoch = e;
  pp.done_counter = done_counter.p;
  pp.i_off = i_off;
  pp.d_off = d_off;
  return pp;
}

void PeerExchange::wait(cudaStream_t st) {
  peer_wait_kernel<<<1, 64, 0, st>>>(world, local_flags(), epoch, status_dev);
  ABSB_CUDA(cudaGetLastError());
}

void PeerExchange::push(const void* src, size_t bytes, cudaStream_t st) {
  ABSB_CHECK(bytes <= slot_bytes && bytes % 16 == 0, ABSB_ERR_INVALID, "record of %zu bytes (capacity %zu, multiple of 16)",
             bytes, slot_bytes);
  ABSB_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0, ABSB_ERR_INVALID, "source must be 16-byte aligned");
  const PeerPush pp = begin_push(0, 0);
  const size_t n16 = bytes / 16;
  const int chunks = (int)std::max<size_t>(1, std::min<size_t>(16, (n16 + 1023) / 1024));
  peer_push_kernel<<<dim3(chunks, world), 256, 0, st>>>(static_cast<const uint4*>(src), n16, world, pp.slot_ptrs,
                                                        pp.flag_ptrs, pp.epoch, pp.done_counter);
  ABSB_CUDA(cudaGetLastError());
  commit();
}

char* PeerExchange::allgather(const void* src, size_t bytes, cudaStream_t st) {
  push(src, bytes, st);
  wait(st);
  return local_entry(epoch);
}

}  // namespace absb
