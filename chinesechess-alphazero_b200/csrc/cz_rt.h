// cz_rt.h — launch / memory shims so the integer-kernel translation units build both with
// nvcc (the product) and with g++ -DCZ_EMUL (tests/simt_emul, CPU test tier only).
#pragma once
#include "cz_simt.h"
#include <stdio.h>
#include <string.h>

#if defined(CZ_EMUL)
#include <functional>
namespace czs { void emul_launch(int nblocks, int nwarps, size_t smem_bytes, const std::function<void()>& body); }
typedef void* cz_stream_t;
#define CZ_LAUNCH(kern, nblocks, nwarps, smem, stream, ...) \
  czs::emul_launch((nblocks), (nwarps), (smem), [=]() { kern(__VA_ARGS__); })
static inline int czrt_last_error(const char** msg) { *msg = ""; return 0; }
static inline int czrt_memset(void* p, int v, size_t n, cz_stream_t) { memset(p, v, n); return 0; }
static inline int czrt_copy(void* d, const void* s, size_t n, cz_stream_t) { memcpy(d, s, n); return 0; }
static inline int czrt_sync(cz_stream_t) { return 0; }
#else
typedef cudaStream_t cz_stream_t;
#define CZ_LAUNCH(kern, nblocks, nwarps, smem, stream, ...) \
  kern<<<(nblocks), (nwarps) * 32, (smem), (stream)>>>(__VA_ARGS__)
static inline int czrt_last_error(const char** msg) {
  cudaError_t e = cudaGetLastError();
  *msg = cudaGetErrorString(e);
  return e == cudaSuccess ? 0 : (int)e;
}
static inline int czrt_memset(void* p, int v, size_t n, cz_stream_t s) { return (int)cudaMemsetAsync(p, v, n, s); }
static inline int czrt_copy(void* d, const void* s, size_t n, cz_stream_t st) {
  return (int)cudaMemcpyAsync(d, s, n, cudaMemcpyDefault, st);
}
static inline int czrt_sync(cz_stream_t s) { return (int)cudaStreamSynchronize(s); }
#endif
