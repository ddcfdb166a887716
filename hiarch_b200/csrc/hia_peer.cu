// Peer-memory plumbing of the row-sharded path: one process per GPU, every
// rank maps the plane and output buffers of the other ranks of the box into its own address space
// (CUDA IPC; NVLink 5 / NVSwitch carries the loads and stores), so that
//   * the planes of a peer's row block are pulled into local HBM by the COPY ENGINES
//     (hia_peer_copy) while the tensor cores contract the previous block -- no SM is taken from the
//     persistent tcgen05 kernel, unlike a send/recv kernel -- and
//   * the mirrored half of every block leaves the contraction's epilogue as plain stores into the
//     peer's rows (hia_sim_gemm_block, out_t).
// The handles travel between the ranks through torch.distributed (host side, hiarch_b200/peer.py).
#include "hia_common.cuh"

typedef int (*PFN_getAddressRange)(unsigned long long*, size_t*, unsigned long long);

// handle64: 64 bytes (cudaIpcMemHandle_t). offset: dev_ptr minus the base of its allocation (the
// handle names the whole cudaMalloc allocation the pointer lies in).
HIA_API int hia_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset) {
  using namespace hia;
  if (!dev_ptr || !handle64 || !offset) return HIA_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  static PFN_getAddressRange fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p) { set_last_cuda_error(e); return HIA_ERR_CUDA; }
    fn = reinterpret_cast<PFN_getAddressRange>(p);
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0) {
    set_last_cuda_error(cudaErrorInvalidValue);
    return HIA_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  HIA_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
  memcpy(handle64, &h, 64);
  *offset = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return HIA_OK;
}

// Maps a peer's allocation; *base_out is valid in THIS process until hia_ipc_close(base).
HIA_API int hia_ipc_open(const void* handle64, void** base_out) {
  using namespace hia;
  if (!handle64 || !base_out) return HIA_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  HIA_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out = p;
  return HIA_OK;
}

HIA_API int hia_ipc_close(void* base) {
  using namespace hia;
  if (!base) return HIA_ERR_BAD_ARG;
  HIA_CUDA(cudaIpcCloseMemHandle(base));
  return HIA_OK;
}

// Stream-ordered copy between any two device pointers of the box (local or peer-mapped): the
// driver routes it through a copy engine over NVLink.
HIA_API int hia_peer_copy(void* dst, const void* src, int64_t bytes, void* stream_) {
  using namespace hia;
  if (!dst || !src || bytes < 0) return HIA_ERR_BAD_ARG;
  if (bytes == 0) return HIA_OK;
  HIA_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream_)));
  return HIA_OK;
}
