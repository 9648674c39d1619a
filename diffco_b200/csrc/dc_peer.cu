// Peer memory for the multi-GPU path: buffers every rank of one box can store into over NVLink (CUDA IPC),
// and a device-side barrier on flags living in those buffers.  With them the tensor-core score kernel writes each tile's
// [score | grad] records straight into EVERY rank's gathered buffer from its epilogue — the all-gather is fused into
// the kernel and overlaps the arithmetic tile by tile; what remains of the collective is one flag round trip.
#include <cstring>

#include "dc_common.cuh"

namespace dc {

double g_peer_timeout_s = 120.0;  // DC_OPT_PEER_TIMEOUT_S

// Thread r: tell rank r that this rank has finished epoch `epoch` (release, system scope), then wait until rank r has
// told us the same (acquire).  Launched after the kernel whose peer stores it publishes, on the same stream.
__global__ void __launch_bounds__(32) peer_barrier_kernel(dc_peer_table flags, int rank, int world, uint32_t epoch,
                                                          long long timeout_cycles) {
  const int r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();
  uint32_t* theirs = static_cast<uint32_t*>(flags.ptr[r]) + rank;  // my slot in rank r's flag array
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const uint32_t* mine = static_cast<const uint32_t*>(flags.ptr[rank]) + r;  // rank r's slot in my flag array
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    // a rank that never arrives (it died, or its launch failed) must not hang the GPU for ever: after the configured
    // time (default 120 s — host-side skew such as a first-call build or data loading is far below that) give up
    if (clock64() - t0 > timeout_cycles) __trap();
  }
}

}  // namespace dc

extern "C" {

int dc_peer_alloc(int64_t bytes, void** ptr, dc_peer_handle* handle) {
  if (bytes < 1 || !ptr || !handle) return DC_ERR_INVALID_ARG;
  void* p = nullptr;
  DC_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    (void)cudaFree(p);
    (void)cudaGetLastError();
    return DC_ERR_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(dc_peer_handle), "dc_peer_handle too small");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    (void)cudaFree(p);
    (void)cudaGetLastError();
    return DC_ERR_CUDA;
  }
  memset(handle, 0, sizeof(*handle));
  memcpy(handle->bytes, &h, sizeof(h));
  *ptr = p;
  return DC_OK;
}

int dc_peer_open(const dc_peer_handle* handle, void** ptr) {
  if (!handle || !ptr) return DC_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle->bytes, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    (void)cudaGetLastError();
    return DC_ERR_CUDA;
  }
  *ptr = p;
  return DC_OK;
}

int dc_peer_close(void* ptr) {
  if (!ptr) return DC_OK;
  if (cudaIpcCloseMemHandle(ptr) != cudaSuccess) {
    (void)cudaGetLastError();
    return DC_ERR_CUDA;
  }
  return DC_OK;
}

int dc_peer_free(void* ptr) {
  if (!ptr) return DC_OK;
  if (cudaFree(ptr) != cudaSuccess) {
    (void)cudaGetLastError();
    return DC_ERR_CUDA;
  }
  return DC_OK;
}

int dc_peer_barrier(const dc_peer_table* flags, int32_t rank, int32_t world, uint32_t epoch, dc_stream_t stream) {
  if (!flags || world < 1 || world > DC_MAX_PEERS || rank < 0 || rank >= world) return DC_ERR_INVALID_ARG;
  for (int r = 0; r < world; ++r)
    if (!flags->ptr[r]) return DC_ERR_INVALID_ARG;
  const long long cycles = (long long)(dc::g_peer_timeout_s * 2.0e9);
  dc::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*flags, rank, world, epoch, cycles);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
