// Error registry + device probing for the C ABI (include/frcnn_b200.h).
#include "common.cuh"
#include "../../include/frcnn_b200.h"
#include <stdarg.h>
#include <string.h>

namespace frcnn {
static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return ERR_CUDA;
}
}  // namespace frcnn

extern "C" int frcnn_version(void) { return 100; }

extern "C" int frcnn_last_error(char* buf, size_t buflen) {
  const size_t n = strlen(frcnn::g_err);
  if (buf && buflen) {
    const size_t c = n < buflen - 1 ? n : buflen - 1;
    memcpy(buf, frcnn::g_err, c);
    buf[c] = 0;
  }
  return (int)n;
}

extern "C" int frcnn_check_device(int device_id) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device_id < 0 || device_id >= count) {
    frcnn::set_error("no CUDA device %d (count=%d, %s)", device_id, count, cudaGetErrorString(e));
    return frcnn::ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  FRCNN_CUDA(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) {
    frcnn::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device_id, prop.major, prop.minor);
    return frcnn::ERR_NO_DEVICE;
  }
  return frcnn::OK;
}

extern "C" int frcnn_zero_async(void* dev_ptr, size_t bytes, void* stream) {
  FRCNN_REQUIRE(dev_ptr, "zero_async: null pointer");
  if (bytes) FRCNN_CUDA(cudaMemsetAsync(dev_ptr, 0, bytes, (cudaStream_t)stream));
  return frcnn::OK;
}

// ---- CUDA-graph capture of a launch sequence (the host layer records its stage calls once per shape and replays them) ----------
struct frcnn_graph { cudaGraphExec_t exec; };

extern "C" int frcnn_graph_begin(void* stream) {
  FRCNN_REQUIRE(stream, "graph_begin: capture needs a non-default stream");
  FRCNN_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
  return frcnn::OK;
}

extern "C" int frcnn_graph_end(void* stream, frcnn_graph** out) {
  FRCNN_REQUIRE(stream && out, "graph_end: null argument");
  *out = nullptr;
  cudaGraph_t g = nullptr;
  FRCNN_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &g));
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return frcnn::cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
  frcnn_graph* h = new frcnn_graph{exec};
  *out = h;
  return frcnn::OK;
}

extern "C" int frcnn_graph_launch(const frcnn_graph* g, void* stream) {
  FRCNN_REQUIRE(g && g->exec, "graph_launch: null graph");
  FRCNN_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  return frcnn::OK;
}

extern "C" void frcnn_graph_destroy(frcnn_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  delete g;
}
