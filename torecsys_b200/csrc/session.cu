// Host-buffer entry point: indices in host memory -> logits in host memory (the e2e path bench.py times).
//
// The batch is cut into `chunks` slices that ping-pong over two streams so that the H2D copy of slice k+1
// overlaps the kernel of slice k and the D2H copy of slice k-1 (PCIe is the bound here: 312 B of int64
// indices per sample against ~3 KB of HBM traffic).  Pinned user buffers are copied from directly; pageable
// ones are staged through the session's pinned buffer.
#include <string.h>

#include <new>

#include "common.cuh"

struct trs_session {
  int64_t max_batch;
  int fields;
  int chunks;
  void* idx_pinned;
  void* idx_dev;
  float* logits_pinned;
  float* logits_dev;
  int32_t* status_dev;
  int32_t* status_pinned;
  cudaStream_t streams[2];
};

using namespace trs;

static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

extern "C" int trs_session_create(int64_t max_batch, int fields, int chunks, trs_session** out_session) {
  TRS_REQUIRE(out_session, "trs_session_create: null out_session");
  TRS_REQUIRE(max_batch > 0 && fields > 0 && chunks > 0, "trs_session_create: bad sizes");
  trs_session* s = new (std::nothrow) trs_session();
  TRS_REQUIRE(s, "trs_session_create: out of host memory");
  memset(s, 0, sizeof(*s));
  s->max_batch = max_batch;
  s->fields = fields;
  s->chunks = chunks;
  const size_t idx_bytes = (size_t)max_batch * fields * sizeof(int64_t);
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMallocHost(&s->idx_pinned, idx_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&s->idx_dev, idx_bytes);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&s->logits_pinned, (size_t)max_batch * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->logits_dev, (size_t)max_batch * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->status_dev, TRS_STATUS_WORDS * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&s->status_pinned, TRS_STATUS_WORDS * sizeof(int32_t));
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&s->streams[i], cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("trs_session_create: %s", cudaGetErrorString(e));
    trs_session_destroy(s);
    return TRS_ERR_CUDA;
  }
  *out_session = s;
  return TRS_OK;
}

extern "C" int trs_session_destroy(trs_session* s) {
  if (!s) return TRS_OK;
  for (int i = 0; i < 2; ++i)
    if (s->streams[i]) cudaStreamDestroy(s->streams[i]);
  if (s->idx_pinned) cudaFreeHost(s->idx_pinned);
  if (s->idx_dev) cudaFree(s->idx_dev);
  if (s->logits_pinned) cudaFreeHost(s->logits_pinned);
  if (s->logits_dev) cudaFree(s->logits_dev);
  if (s->status_dev) cudaFree(s->status_dev);
  if (s->status_pinned) cudaFreeHost(s->status_pinned);
  delete s;
  return TRS_OK;
}

static int session_run(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets, int64_t batch,
                       int fields, const float* w_feat, const float* w_emb, const float* packed, int64_t rows,
                       int embed, const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                       const float* const* mlp_b, int activation, float* logits_host, int64_t* oob_count) {
  TRS_REQUIRE(s && idx_host && logits_host, "trs_session_deepfm_forward_host: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_session_deepfm_forward_host: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && batch <= s->max_batch && fields == s->fields,
              "trs_session_deepfm_forward_host: batch/fields exceed the session (%lld x %d)",
              (long long)s->max_batch, s->fields);
  if (batch == 0) {
    if (oob_count) *oob_count = 0;
    return TRS_OK;
  }
  const size_t isz = idx_bits / 8;
  const bool src_pinned = is_pinned(idx_host);
  const bool dst_pinned = is_pinned(logits_host);
  TRS_CUDA(cudaMemsetAsync(s->status_dev, 0, TRS_STATUS_WORDS * sizeof(int32_t), s->streams[0]));
  cudaEvent_t ready;
  TRS_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  TRS_CUDA(cudaEventRecord(ready, s->streams[0]));
  TRS_CUDA(cudaStreamWaitEvent(s->streams[1], ready, 0));
  const int chunks = (int)(batch < s->chunks ? batch : s->chunks);
  const int64_t per = (((batch + chunks - 1) / chunks + 15) / 16) * 16;  // slices start on 16-byte aligned indices
  int rc = TRS_OK;
  for (int c = 0; c < chunks && rc == TRS_OK; ++c) {
    const int64_t b0 = c * per;
    const int64_t nb = batch - b0 < per ? batch - b0 : per;
    if (nb <= 0) break;
    cudaStream_t st = s->streams[c & 1];
    const size_t off = (size_t)b0 * fields * isz, bytes = (size_t)nb * fields * isz;
    const char* src = static_cast<const char*>(idx_host) + off;
    if (!src_pinned) {
      memcpy(static_cast<char*>(s->idx_pinned) + off, src, bytes);
      src = static_cast<const char*>(s->idx_pinned) + off;
    }
    TRS_CUDA(cudaMemcpyAsync(static_cast<char*>(s->idx_dev) + off, src, bytes, cudaMemcpyHostToDevice, st));
    if (packed != nullptr)
      rc = trs_deepfm_forward_packed(static_cast<char*>(s->idx_dev) + off, idx_bits, offsets, nb, fields, packed, rows,
                                     mlp_dims, mlp_layers, mlp_w, mlp_b, activation, s->logits_dev + b0,
                                     s->status_dev, st);
    else
      rc = trs_deepfm_forward(static_cast<char*>(s->idx_dev) + off, idx_bits, offsets, nb, fields, w_feat, w_emb,
                              rows, embed, mlp_dims, mlp_layers, mlp_w, mlp_b, activation, s->logits_dev + b0,
                              s->status_dev, st);
    if (rc != TRS_OK) break;
    float* dst = dst_pinned ? logits_host + b0 : s->logits_pinned + b0;
    TRS_CUDA(cudaMemcpyAsync(dst, s->logits_dev + b0, (size_t)nb * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  cudaError_t e0 = cudaStreamSynchronize(s->streams[1]);
  cudaError_t e1 = cudaSuccess;
  if (rc == TRS_OK && e0 == cudaSuccess)
    e1 = cudaMemcpyAsync(s->status_pinned, s->status_dev, TRS_STATUS_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost,
                         s->streams[0]);
  cudaError_t e2 = cudaStreamSynchronize(s->streams[0]);
  cudaEventDestroy(ready);
  if (rc != TRS_OK) return rc;
  if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) {
    cudaError_t e = e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2);
    set_error("trs_session_deepfm_forward_host: %s", cudaGetErrorString(e));
    return TRS_ERR_CUDA;
  }
  if (!dst_pinned) memcpy(logits_host, s->logits_pinned, (size_t)batch * sizeof(float));
  if (oob_count) *oob_count = s->status_pinned[0];
  return TRS_OK;
}

extern "C" int trs_session_deepfm_forward_host(trs_session* s, const void* idx_host, int idx_bits,
                                               const int64_t* offsets, int64_t batch, int fields,
                                               const float* w_feat, const float* w_emb, int64_t rows, int embed,
                                               const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                               const float* const* mlp_b, int activation, float* logits_host,
                                               int64_t* oob_count) {
  return session_run(s, idx_host, idx_bits, offsets, batch, fields, w_feat, w_emb, nullptr, rows, embed, mlp_dims,
                     mlp_layers, mlp_w, mlp_b, activation, logits_host, oob_count);
}

extern "C" int trs_session_deepfm_forward_host_packed(trs_session* s, const void* idx_host, int idx_bits,
                                                      const int64_t* offsets, int64_t batch, int fields,
                                                      const float* packed, int64_t rows, const int* mlp_dims,
                                                      int mlp_layers, const float* const* mlp_w,
                                                      const float* const* mlp_b, int activation,
                                                      float* logits_host, int64_t* oob_count) {
  TRS_REQUIRE(packed, "trs_session_deepfm_forward_host_packed: null packed table");
  return session_run(s, idx_host, idx_bits, offsets, batch, fields, nullptr, nullptr, packed, rows, 16, mlp_dims,
                     mlp_layers, mlp_w, mlp_b, activation, logits_host, oob_count);
}
