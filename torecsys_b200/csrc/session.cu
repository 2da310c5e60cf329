// Host-buffer entry points: indices in host memory -> logits in host memory (the e2e path bench.py times).
//
// A session owns kSlots independent "slots" (device index/logit buffers, pinned staging, status words, two streams,
// one completion event).  A submitted batch occupies one slot: its rows are cut into `chunks` slices that ping-pong
// over the slot's two streams so that the H2D copy of slice k+1 overlaps the kernel of slice k and the D2H copy of
// slice k-1, and -- because slots are independent -- the copies of batch k+1 overlap the kernel and the read-back of
// batch k.  PCIe is the bound here (312 B of int64 indices per sample against ~3 KB of HBM traffic), so the job of this
// file is to keep the host->device copy engine busy all the time.
//   trs_session_submit_*  enqueue one batch, return a ticket at once (no host synchronisation)
//   trs_session_wait      block until that batch's logits are in the caller's host buffer
//   trs_session_deepfm_forward_host[_packed] = submit + wait (the synchronous call)
// Pinned user buffers are copied from/to directly; pageable ones are staged through the slot's pinned buffers.
#include <string.h>

#include <new>

#include "common.cuh"

namespace {

constexpr int kSlots = 3;

struct Slot {
  void* idx_pinned;
  void* idx_dev;
  float* logits_pinned;
  float* logits_dev;
  int32_t* status_dev;
  int32_t* status_pinned;
  cudaStream_t streams[2];
  cudaEvent_t forked, joined, done;
  // the batch in flight (busy != 0)
  int busy;
  int64_t ticket;
  int64_t batch;
  float* user_logits;   // non-null when the logits have to be copied out of logits_pinned at wait time
};

}  // namespace

struct trs_session {
  int64_t max_batch;
  int fields;
  int chunks;
  int64_t next_ticket;
  Slot slots[kSlots];
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
  s->next_ticket = 1;
  const size_t idx_bytes = (size_t)max_batch * fields * sizeof(int64_t);
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < kSlots; ++k) {
    Slot& sl = s->slots[k];
    if (e == cudaSuccess) e = cudaMallocHost(&sl.idx_pinned, idx_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&sl.idx_dev, idx_bytes);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&sl.logits_pinned, (size_t)max_batch * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&sl.logits_dev, (size_t)max_batch * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&sl.status_dev, TRS_STATUS_WORDS * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&sl.status_pinned, TRS_STATUS_WORDS * sizeof(int32_t));
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&sl.streams[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.forked, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.joined, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
  }
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
  for (int k = 0; k < kSlots; ++k) {
    Slot& sl = s->slots[k];
    for (int i = 0; i < 2; ++i)
      if (sl.streams[i]) {
        cudaStreamSynchronize(sl.streams[i]);
        cudaStreamDestroy(sl.streams[i]);
      }
    if (sl.forked) cudaEventDestroy(sl.forked);
    if (sl.joined) cudaEventDestroy(sl.joined);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.idx_pinned) cudaFreeHost(sl.idx_pinned);
    if (sl.idx_dev) cudaFree(sl.idx_dev);
    if (sl.logits_pinned) cudaFreeHost(sl.logits_pinned);
    if (sl.logits_dev) cudaFree(sl.logits_dev);
    if (sl.status_dev) cudaFree(sl.status_dev);
    if (sl.status_pinned) cudaFreeHost(sl.status_pinned);
  }
  delete s;
  return TRS_OK;
}

extern "C" int trs_session_depth(void) { return kSlots; }

// Enqueues one batch on a free slot.  On any failure after the first enqueue the slot's streams are drained so that
// the slot is reusable.
static int session_submit(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets, int64_t batch,
                          int fields, const float* w_feat, const float* w_emb, const float* packed, int64_t rows,
                          int embed, const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                          const float* const* mlp_b, int activation, float* logits_host, int64_t* ticket) {
  TRS_REQUIRE(s && idx_host && logits_host && ticket, "trs_session_submit: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_session_submit: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && batch <= s->max_batch && fields == s->fields,
              "trs_session_submit: batch/fields exceed the session (%lld x %d)", (long long)s->max_batch, s->fields);
  Slot* slp = nullptr;
  for (int k = 0; k < kSlots && !slp; ++k)
    if (!s->slots[k].busy) slp = &s->slots[k];
  TRS_REQUIRE(slp, "trs_session_submit: all %d slots are in flight -- trs_session_wait() for a ticket first", kSlots);
  Slot& sl = *slp;
  sl.ticket = s->next_ticket++;
  sl.batch = batch;
  sl.user_logits = nullptr;
  sl.status_pinned[0] = 0;
  *ticket = sl.ticket;
  sl.busy = 1;
  if (batch == 0) {
    TRS_CUDA(cudaEventRecord(sl.done, sl.streams[0]));
    return TRS_OK;
  }
  const size_t isz = idx_bits / 8;
  const bool src_pinned = is_pinned(idx_host);
  const bool dst_pinned = is_pinned(logits_host);
  if (!dst_pinned) sl.user_logits = logits_host;
  int rc = TRS_OK;
  cudaError_t e = cudaMemsetAsync(sl.status_dev, 0, TRS_STATUS_WORDS * sizeof(int32_t), sl.streams[0]);
  if (e == cudaSuccess) e = cudaEventRecord(sl.forked, sl.streams[0]);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(sl.streams[1], sl.forked, 0);
  const int chunks = (int)(batch < s->chunks ? batch : s->chunks);
  const int64_t per = (((batch + chunks - 1) / chunks + 15) / 16) * 16;  // slices start on 16-byte aligned indices
  for (int c = 0; c < chunks && rc == TRS_OK && e == cudaSuccess; ++c) {
    const int64_t b0 = c * per;
    const int64_t nb = batch - b0 < per ? batch - b0 : per;
    if (nb <= 0) break;
    cudaStream_t st = sl.streams[c & 1];
    const size_t off = (size_t)b0 * fields * isz, bytes = (size_t)nb * fields * isz;
    const char* src = static_cast<const char*>(idx_host) + off;
    if (!src_pinned) {
      memcpy(static_cast<char*>(sl.idx_pinned) + off, src, bytes);
      src = static_cast<const char*>(sl.idx_pinned) + off;
    }
    e = cudaMemcpyAsync(static_cast<char*>(sl.idx_dev) + off, src, bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) break;
    if (packed != nullptr)
      rc = trs_deepfm_forward_packed(static_cast<char*>(sl.idx_dev) + off, idx_bits, offsets, nb, fields, packed, rows,
                                     mlp_dims, mlp_layers, mlp_w, mlp_b, activation, sl.logits_dev + b0,
                                     sl.status_dev, st);
    else
      rc = trs_deepfm_forward(static_cast<char*>(sl.idx_dev) + off, idx_bits, offsets, nb, fields, w_feat, w_emb,
                              rows, embed, mlp_dims, mlp_layers, mlp_w, mlp_b, activation, sl.logits_dev + b0,
                              sl.status_dev, st);
    if (rc != TRS_OK) break;
    float* dst = dst_pinned ? logits_host + b0 : sl.logits_pinned + b0;
    e = cudaMemcpyAsync(dst, sl.logits_dev + b0, (size_t)nb * sizeof(float), cudaMemcpyDeviceToHost, st);
  }
  // join stream 1 into stream 0, read the status words back, mark completion
  if (rc == TRS_OK && e == cudaSuccess) e = cudaEventRecord(sl.joined, sl.streams[1]);
  if (rc == TRS_OK && e == cudaSuccess) e = cudaStreamWaitEvent(sl.streams[0], sl.joined, 0);
  if (rc == TRS_OK && e == cudaSuccess)
    e = cudaMemcpyAsync(sl.status_pinned, sl.status_dev, TRS_STATUS_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost,
                        sl.streams[0]);
  if (rc == TRS_OK && e == cudaSuccess) e = cudaEventRecord(sl.done, sl.streams[0]);
  if (rc != TRS_OK || e != cudaSuccess) {
    cudaStreamSynchronize(sl.streams[0]);
    cudaStreamSynchronize(sl.streams[1]);
    sl.busy = 0;
    if (rc != TRS_OK) return rc;   // message already set by the failing entry point
    set_error("trs_session_submit: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TRS_ERR_CUDA;
  }
  return TRS_OK;
}

extern "C" int trs_session_wait(trs_session* s, int64_t ticket, int64_t* oob_count) {
  TRS_REQUIRE(s, "trs_session_wait: null session");
  Slot* slp = nullptr;
  for (int k = 0; k < kSlots && !slp; ++k)
    if (s->slots[k].busy && s->slots[k].ticket == ticket) slp = &s->slots[k];
  TRS_REQUIRE(slp, "trs_session_wait: ticket %lld is not in flight", (long long)ticket);
  Slot& sl = *slp;
  cudaError_t e = cudaEventSynchronize(sl.done);
  sl.busy = 0;
  if (e != cudaSuccess) {
    set_error("trs_session_wait: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TRS_ERR_CUDA;
  }
  if (sl.user_logits && sl.batch > 0) memcpy(sl.user_logits, sl.logits_pinned, (size_t)sl.batch * sizeof(float));
  if (oob_count) *oob_count = sl.batch > 0 ? sl.status_pinned[0] : 0;
  return TRS_OK;
}

extern "C" int trs_session_submit_deepfm(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets,
                                         int64_t batch, int fields, const float* w_feat, const float* w_emb,
                                         int64_t rows, int embed, const int* mlp_dims, int mlp_layers,
                                         const float* const* mlp_w, const float* const* mlp_b, int activation,
                                         float* logits_host, int64_t* ticket) {
  return session_submit(s, idx_host, idx_bits, offsets, batch, fields, w_feat, w_emb, nullptr, rows, embed, mlp_dims,
                        mlp_layers, mlp_w, mlp_b, activation, logits_host, ticket);
}

extern "C" int trs_session_submit_deepfm_packed(trs_session* s, const void* idx_host, int idx_bits,
                                                const int64_t* offsets, int64_t batch, int fields,
                                                const float* packed, int64_t rows, const int* mlp_dims,
                                                int mlp_layers, const float* const* mlp_w,
                                                const float* const* mlp_b, int activation, float* logits_host,
                                                int64_t* ticket) {
  TRS_REQUIRE(packed, "trs_session_submit_deepfm_packed: null packed table");
  return session_submit(s, idx_host, idx_bits, offsets, batch, fields, nullptr, nullptr, packed, rows, 16, mlp_dims,
                        mlp_layers, mlp_w, mlp_b, activation, logits_host, ticket);
}

extern "C" int trs_session_deepfm_forward_host(trs_session* s, const void* idx_host, int idx_bits,
                                               const int64_t* offsets, int64_t batch, int fields,
                                               const float* w_feat, const float* w_emb, int64_t rows, int embed,
                                               const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                               const float* const* mlp_b, int activation, float* logits_host,
                                               int64_t* oob_count) {
  int64_t ticket = 0;
  int rc = session_submit(s, idx_host, idx_bits, offsets, batch, fields, w_feat, w_emb, nullptr, rows, embed, mlp_dims,
                          mlp_layers, mlp_w, mlp_b, activation, logits_host, &ticket);
  if (rc != TRS_OK) return rc;
  return trs_session_wait(s, ticket, oob_count);
}

extern "C" int trs_session_deepfm_forward_host_packed(trs_session* s, const void* idx_host, int idx_bits,
                                                      const int64_t* offsets, int64_t batch, int fields,
                                                      const float* packed, int64_t rows, const int* mlp_dims,
                                                      int mlp_layers, const float* const* mlp_w,
                                                      const float* const* mlp_b, int activation,
                                                      float* logits_host, int64_t* oob_count) {
  TRS_REQUIRE(packed, "trs_session_deepfm_forward_host_packed: null packed table");
  int64_t ticket = 0;
  int rc = session_submit(s, idx_host, idx_bits, offsets, batch, fields, nullptr, nullptr, packed, rows, 16, mlp_dims,
                          mlp_layers, mlp_w, mlp_b, activation, logits_host, &ticket);
  if (rc != TRS_OK) return rc;
  return trs_session_wait(s, ticket, oob_count);
}
