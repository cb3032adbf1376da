// Peer-memory frame sink: finished frames travel to the gathering rank as copy-engine writes into
// a buffer that rank exported over CUDA IPC — no kernel runs on either side, so the gather takes
// no SMs and no HBM bandwidth from the rasterizer of the receiving rank (an NCCL gather funnels
// 7 x 68 MB per step through receive kernels on rank 0).  The reference has no counterpart: its
// ranks write PNGs to disk and rank 0 globs them (engines/visualizer_pgdvs.py:160-177).
#include "common.cuh"

#include <string.h>

using namespace pgdvs;

static_assert(sizeof(cudaIpcMemHandle_t) == PGDVS_IPC_HANDLE_BYTES, "handle size");

extern "C" int pgdvs_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out) {
  if (dev_ptr == nullptr || handle_out == nullptr || bytes == 0) return PGDVS_E_BADARG;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return PGDVS_OK;
}

extern "C" int pgdvs_ipc_open(const unsigned char* handle, void** dev_ptr) {
  if (handle == nullptr || dev_ptr == nullptr) return PGDVS_E_BADARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  *dev_ptr = p;
  return PGDVS_OK;
}

extern "C" int pgdvs_ipc_close(void* dev_ptr) {
  if (dev_ptr == nullptr) return PGDVS_E_BADARG;
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  return e == cudaSuccess ? PGDVS_OK : (int)e;
}

extern "C" int pgdvs_ipc_free(void* dev_ptr) {
  if (dev_ptr == nullptr) return PGDVS_E_BADARG;
  cudaError_t e = cudaFree(dev_ptr);
  return e == cudaSuccess ? PGDVS_OK : (int)e;
}

extern "C" int pgdvs_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return PGDVS_OK;
  if (dst == nullptr || src == nullptr) return PGDVS_E_BADARG;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  return e == cudaSuccess ? PGDVS_OK : (int)e;
}
