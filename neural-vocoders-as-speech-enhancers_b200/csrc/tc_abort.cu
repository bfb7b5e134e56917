// Host side of the bounded-wait failure flag of the tensor-core kernels (tc_ptx.cuh).
//
// Every translation unit with tcgen05 kernels holds its own device copy of the flag; a kernel whose mbarrier wait
// times out raises it and, through a device pointer to ONE pinned + mapped host word, lets the host see the failure
// with a plain memory read.  The product entry points (nvse_generator_forward / _forward_train / _backward, the
// layer-level tensor-core calls) poll that word on entry: if a kernel of an EARLIER call timed out, the flags are
// cleared (stream-ordered, so later kernels are not aborted by a stale flag) and the call returns NVSE_ERR_STATE --
// the results produced since the failing call are invalid and the caller must not use them.
#include <mutex>

#include "generator.cuh"
#include "grad.cuh"
#include "resblock_tc.cuh"
#include "ups_tc.cuh"

namespace nvse {

namespace {
std::mutex g_mu;
unsigned int* g_host_word = nullptr;  // pinned, mapped, portable
bool g_bound[64] = {};
}  // namespace

int tc_abort_bind_device() {
  int dev = 0;
  NVSE_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_mu);
  if (dev >= 0 && dev < 64 && g_bound[dev]) return NVSE_OK;
  if (!g_host_word) {
    NVSE_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&g_host_word), sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable));
    *g_host_word = 0u;
  }
  unsigned int* dptr = nullptr;
  NVSE_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), g_host_word, 0));
  if (int rc = tc_abort_bind(dptr)) return rc;
  if (int rc = rb_abort_bind(dptr)) return rc;
  if (int rc = pair_abort_bind(dptr)) return rc;
  if (int rc = wgrad_abort_bind(dptr)) return rc;
  if (int rc = ups_abort_bind(dptr)) return rc;
  if (dev >= 0 && dev < 64) g_bound[dev] = true;
  return NVSE_OK;
}

int tc_abort_poll(cudaStream_t st) {
  if (int rc = tc_abort_bind_device()) return rc;
  if (*reinterpret_cast<volatile unsigned int*>(g_host_word) == 0u) return NVSE_OK;
  *reinterpret_cast<volatile unsigned int*>(g_host_word) = 0u;
  if (int rc = tc_abort_clear(st)) return rc;
  if (int rc = rb_abort_clear(st)) return rc;
  if (int rc = pair_abort_clear(st)) return rc;
  if (int rc = wgrad_abort_clear(st)) return rc;
  if (int rc = ups_abort_clear(st)) return rc;
  return fail(NVSE_ERR_STATE, "a tensor-core kernel of an earlier call gave up on a bounded wait (> 0.2 s: GPU time-slicing, a debugger, or a "
                              "protocol bug): every result since that call is invalid; the failure flags have been cleared");
}

void tc_abort_host_clear() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_host_word) *reinterpret_cast<volatile unsigned int*>(g_host_word) = 0u;
}

}  // namespace nvse
