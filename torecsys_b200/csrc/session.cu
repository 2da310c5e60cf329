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
#include <immintrin.h>
#include <sched.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kSlots = 3;

// ---- host-side index narrowing ------------------------------------------------------------------------------------
// The link is the bound of this path and int64 indices are half zeros: a small pool of host threads converts the
// caller's int64 (B, N) block into int32 in the slot's pinned staging buffer, slice by slice, while the copy engine
// is busy with the previous slices/batches; the kernels then run their int32 instantiation.  A value that does not
// fit int32 becomes INT32_MIN, which stays out of range after the (non-negative, < 2^31) field offset is added, so
// the lookup is reported exactly as in the int64 path (rows < 2^31 is a precondition of int32 indices anyway).
struct NarrowBlock {
  int64_t begin, end;   // element range
  int chunk;
};

// Scalar form (also the fix-up of a vector in which some value did not fit).
inline void narrow_scalar(const int64_t* __restrict__ s, int32_t* __restrict__ d, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const int64_t v = s[i];
    d[i] = static_cast<int64_t>(static_cast<int32_t>(v)) == v ? static_cast<int32_t>(v) : INT32_MIN;
  }
}
// Explicit SIMD forms: the loop above does not auto-vectorise into anything useful (64 -> 32-bit packing plus a select), and
// ordinary stores into the pinned staging buffer cost a read-for-ownership of every line -- measured slower than sending
// the int64 block as it is.  These convert 16 (AVX-512: vpmovqd) or 8 (AVX2: dword permutes) values per step, write with
// NON-TEMPORAL stores where the destination is 32 / 64-byte aligned, and re-do a vector in scalar form only when one of
// its values fails the sign-extension test (never, for valid indices).
__attribute__((target("avx512f"))) void narrow_avx512(const int64_t* __restrict__ s, int32_t* __restrict__ d, int64_t n) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 63) != 0) { narrow_scalar(s + i, d + i, 1); ++i; }
  for (; i + 16 <= n; i += 16) {
    const __m512i a = _mm512_loadu_si512(s + i), b = _mm512_loadu_si512(s + i + 8);
    const __m256i na = _mm512_cvtepi64_epi32(a), nb = _mm512_cvtepi64_epi32(b);
    const __mmask8 bad = _mm512_cmpneq_epi64_mask(_mm512_cvtepi32_epi64(na), a) |
                         _mm512_cmpneq_epi64_mask(_mm512_cvtepi32_epi64(nb), b);
    const __m512i both = _mm512_inserti64x4(_mm512_castsi256_si512(na), nb, 1);
    _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), both);
    if (bad) narrow_scalar(s + i, d + i, 16);
  }
  narrow_scalar(s + i, d + i, n - i);
  _mm_sfence();
}
__attribute__((target("avx2"))) void narrow_avx2(const int64_t* __restrict__ s, int32_t* __restrict__ d, int64_t n) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 31) != 0) { narrow_scalar(s + i, d + i, 1); ++i; }
  const __m256i even = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
  for (; i + 8 <= n; i += 8) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 4));
    // a value fits iff its high dword equals the sign fill of its low dword
    const __m256i sa = _mm256_shuffle_epi32(_mm256_srai_epi32(a, 31), _MM_SHUFFLE(2, 2, 0, 0));
    const __m256i sb = _mm256_shuffle_epi32(_mm256_srai_epi32(b, 31), _MM_SHUFFLE(2, 2, 0, 0));
    const int ok = _mm256_movemask_epi8(_mm256_and_si256(_mm256_cmpeq_epi32(a, sa), _mm256_cmpeq_epi32(b, sb)));
    const __m256i pa = _mm256_permutevar8x32_epi32(a, even), pb = _mm256_permutevar8x32_epi32(b, even);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), _mm256_permute2x128_si256(pa, pb, 0x20));
    if ((ok & 0xf0f0f0f0) != static_cast<int>(0xf0f0f0f0)) narrow_scalar(s + i, d + i, 8);   // (high dwords: bytes 4-7 of each value)
  }
  narrow_scalar(s + i, d + i, n - i);
  _mm_sfence();
}
void narrow_generic(const int64_t* __restrict__ s, int32_t* __restrict__ d, int64_t n) { narrow_scalar(s, d, n); }
typedef void (*narrow_fn)(const int64_t*, int32_t*, int64_t);
narrow_fn pick_narrow() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return narrow_avx512;
  return __builtin_cpu_supports("avx2") ? narrow_avx2 : narrow_generic;
}

class NarrowPool {
 public:
  explicit NarrowPool(int threads) {
    for (int i = 0; i < threads; ++i) workers_.emplace_back([this] { worker(); });
  }
  ~NarrowPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_.store(true, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int threads() const { return static_cast<int>(workers_.size()); }

  // starts narrowing src[0, n) -> dst; blocks never straddle a chunk boundary (chunk c = elements [c*per, (c+1)*per))
  void start(const int64_t* src, int32_t* dst, int64_t n, int64_t per, int chunks) {
    constexpr int64_t kBlock = 16384;   // 128 KB in, 64 KB out per block
    blocks_.clear();
    chunk_blocks_.assign(chunks, 0);
    for (int c = 0; c < chunks; ++c) {
      const int64_t lo = c * per, hi = (c + 1) * per < n ? (c + 1) * per : n;
      for (int64_t b = lo; b < hi; b += kBlock) {
        blocks_.push_back({b, b + kBlock < hi ? b + kBlock : hi, c});
        ++chunk_blocks_[c];
      }
    }
    for (int c = 0; c < chunks; ++c) chunk_done_[c].store(0, std::memory_order_relaxed);
    src_ = src;
    dst_ = dst;
    next_.store(0, std::memory_order_relaxed);
    active_.store(threads(), std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);
      seq_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
  }
  // the calling thread helps until chunk c is complete
  void finish_chunk(int c) {
    while (chunk_done_[c].load(std::memory_order_acquire) < chunk_blocks_[c])
      if (!run_one()) __builtin_ia32_pause();
  }
  // all workers have left the job (its block list may be rebuilt)
  void quiesce() {
    while (active_.load(std::memory_order_acquire) != 0) __builtin_ia32_pause();
  }
  static constexpr int kMaxChunks = 64;

 private:
  bool run_one() {
    const int64_t b = next_.fetch_add(1, std::memory_order_relaxed);
    if (b >= static_cast<int64_t>(blocks_.size())) return false;
    const NarrowBlock& blk = blocks_[b];
    fn_(src_ + blk.begin, dst_ + blk.begin, blk.end - blk.begin);
    chunk_done_[blk.chunk].fetch_add(1, std::memory_order_release);
    return true;
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      int spins = 0;
      while (seq_.load(std::memory_order_acquire) == seen && !stop_.load(std::memory_order_acquire)) {
        if (++spins < 20000) {
          __builtin_ia32_pause();
        } else {   // idle for a few hundred microseconds: sleep until the next job.  (Helpers that never sleep were
                   // measured: they take the cores from the submitting thread and the driver -- 235 -> 155 M samples/s.)
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return seq_.load(std::memory_order_acquire) != seen || stop_.load(); });
        }
      }
      if (stop_.load(std::memory_order_acquire)) return;
      seen = seq_.load(std::memory_order_acquire);
      while (run_one()) {
      }
      active_.fetch_sub(1, std::memory_order_release);
    }
  }

  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<uint64_t> seq_{0};
  std::atomic<bool> stop_{false};
  std::atomic<int> active_{0};
  std::atomic<int64_t> next_{0};
  std::atomic<int> chunk_done_[kMaxChunks];
  std::vector<int> chunk_blocks_;
  std::vector<NarrowBlock> blocks_;
  const int64_t* src_ = nullptr;
  int32_t* dst_ = nullptr;
  narrow_fn fn_ = pick_narrow();
};

int usable_cpus() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int n = CPU_COUNT(&set);
    if (n > 0) return n;
  }
  const unsigned hc = std::thread::hardware_concurrency();
  return hc > 0 ? static_cast<int>(hc) : 1;
}

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
  NarrowPool* narrow;   // null: int64 indices cross the link as they are
  cudaStream_t producer;   // stream on which the caller prepares tables / parameters (ordering, see submit)
  int has_producer;
  cudaEvent_t producer_ready;
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
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->producer_ready, cudaEventDisableTiming);
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
  delete s->narrow;
  if (s->producer_ready) cudaEventDestroy(s->producer_ready);
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

// test hook (CPU suite): runs one of the narrowing forms on host arrays.  which: 0 = the form the sessions use, 1 = scalar,
// 2 = AVX2, 3 = AVX-512 (TRS_ERR_UNSUPPORTED when the CPU lacks it)
extern "C" int trs_host_narrow_indices(const int64_t* src, int32_t* dst, int64_t n, int which) {
  TRS_REQUIRE(src && dst && n >= 0 && which >= 0 && which <= 3, "trs_host_narrow_indices: bad arguments");
  __builtin_cpu_init();
  if (which == 0) pick_narrow()(src, dst, n);
  else if (which == 1) narrow_generic(src, dst, n);
  else if (which == 2) {
    TRS_UNSUPPORTED(!__builtin_cpu_supports("avx2"), "trs_host_narrow_indices: no AVX2 on this CPU");
    narrow_avx2(src, dst, n);
  } else {
    TRS_UNSUPPORTED(!__builtin_cpu_supports("avx512f"), "trs_host_narrow_indices: no AVX-512 on this CPU");
    narrow_avx512(src, dst, n);
  }
  return TRS_OK;
}

// The sessions' narrowing POOL on plain host arrays (no device involved): `threads` threads (the caller included) convert
// src[0, n) -> dst `reps` times; returns the best time of one pass in nanoseconds (tools / tests: how the conversion
// scales over the host's cores), or a negative TRS_ERR_* code.
extern "C" int64_t trs_host_narrow_pool_ns(const int64_t* src, int32_t* dst, int64_t n, int threads, int reps) {
  if (!src || !dst || n < 0 || threads < 1 || threads > 256 || reps < 1) {
    set_error("trs_host_narrow_pool_ns: bad arguments");
    return TRS_ERR_INVALID_ARGUMENT;
  }
  NarrowPool pool(threads - 1);
  int64_t best = -1;
  for (int r = 0; r < reps; ++r) {
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pool.start(src, dst, n, n > 0 ? n : 1, 1);
    pool.finish_chunk(0);
    pool.quiesce();
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const int64_t ns = (t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec);
    if (best < 0 || ns < best) best = ns;
  }
  return best;
}

extern "C" int trs_session_set_index_narrowing(trs_session* s, int threads) {
  TRS_REQUIRE(s, "trs_session_set_index_narrowing: null session");
  TRS_REQUIRE(threads >= -1 && threads <= 256, "trs_session_set_index_narrowing: threads must be in [-1, 256]");
  for (int k = 0; k < kSlots; ++k)
    TRS_REQUIRE(!s->slots[k].busy, "trs_session_set_index_narrowing: batches are in flight");
  if (threads < 0) {   // auto: half of the CPUs this process may run on, at most 8 helpers
    const int cpus = usable_cpus();
    threads = cpus / 2 < 8 ? cpus / 2 : 8;
  }
  const int helpers = threads > 0 ? threads - 1 : 0;   // the submitting thread is one of the `threads`
  delete s->narrow;
  s->narrow = nullptr;
  if (threads > 0) {
    s->narrow = new (std::nothrow) NarrowPool(helpers);
    TRS_REQUIRE(s->narrow, "trs_session_set_index_narrowing: out of host memory");
  }
  return threads;
}

namespace {
// what a slice of a batch runs: any fused indices -> logits entry point of this library, bound to its parameters
struct DeepFmCall {
  const int64_t* offsets;
  int fields;
  const float *w_feat, *w_emb, *packed, *workspace;
  int64_t rows;
  int embed, variant;
  const int* mlp_dims;
  int mlp_layers;
  const float* const* mlp_w;
  const float* const* mlp_b;
  int activation;
};
int deepfm_call(void* ctx, int /*lane*/, const void* idx_dev, int idx_bits, int64_t nb, float* logits_dev,
                int32_t* status_dev, void* stream) {
  const DeepFmCall& c = *static_cast<const DeepFmCall*>(ctx);
  if (c.workspace != nullptr)
    return trs_deepfm_forward_tc(idx_dev, idx_bits, c.offsets, nb, c.fields, c.packed, c.rows, c.mlp_dims, c.mlp_layers,
                                 c.mlp_w, c.mlp_b, c.activation, c.workspace, c.variant, logits_dev, status_dev, 0u,
                                 stream);
  if (c.packed != nullptr)
    return trs_deepfm_forward_packed(idx_dev, idx_bits, c.offsets, nb, c.fields, c.packed, c.rows, c.mlp_dims,
                                     c.mlp_layers, c.mlp_w, c.mlp_b, c.activation, logits_dev, status_dev, stream);
  return trs_deepfm_forward(idx_dev, idx_bits, c.offsets, nb, c.fields, c.w_feat, c.w_emb, c.rows, c.embed, c.mlp_dims,
                            c.mlp_layers, c.mlp_w, c.mlp_b, c.activation, logits_dev, status_dev, stream);
}
struct FmCall {
  const int64_t* offsets;
  int fields;
  const float *w_feat, *w_emb, *packed, *bias;
  int64_t rows;
  int embed;
};
int fm_call(void* ctx, int, const void* idx_dev, int idx_bits, int64_t nb, float* logits_dev, int32_t* status_dev,
            void* stream) {
  const FmCall& c = *static_cast<const FmCall*>(ctx);
  if (c.packed != nullptr)
    return trs_fm_model_forward_packed(idx_dev, idx_bits, c.offsets, nb, c.fields, c.packed, c.rows, c.bias, logits_dev,
                                       status_dev, stream);
  return trs_fm_model_forward(idx_dev, idx_bits, c.offsets, nb, c.fields, c.w_feat, c.w_emb, c.rows, c.embed, c.bias,
                              logits_dev, status_dev, stream);
}
struct DcnCall {
  const int64_t* offsets;
  int fields;
  const float* w_emb;
  int64_t rows;
  int embed;
  const float *cross_w, *cross_b;
  int cross_layers;
  const int* mlp_dims;
  int mlp_layers;
  const float* const* mlp_w;
  const float* const* mlp_b;
  int activation;
  const float *fc_w, *fc_b;
};
int dcn_call(void* ctx, int, const void* idx_dev, int idx_bits, int64_t nb, float* logits_dev, int32_t* status_dev,
             void* stream) {
  const DcnCall& c = *static_cast<const DcnCall*>(ctx);
  return trs_dcn_forward(idx_dev, idx_bits, c.offsets, nb, c.fields, c.w_emb, c.rows, c.embed, c.cross_w, c.cross_b,
                         c.cross_layers, c.mlp_dims, c.mlp_layers, c.mlp_w, c.mlp_b, c.activation, c.fc_w, c.fc_b,
                         logits_dev, status_dev, stream);
}
struct XdfmCall {
  const int64_t* offsets;
  int fields;
  const float *w_feat, *w_emb;
  int64_t rows;
  int embed;
  const float* const* cin_w;
  const float* const* cin_scale;
  const float* const* cin_shift;
  const int* cin_sizes;
  int cin_layers, cin_is_direct, cin_act;
  const float *cin_fc_w, *cin_fc_b;
  const int* mlp_dims;
  int mlp_layers;
  const float* const* mlp_w;
  const float* const* mlp_b;
  int mlp_act;
  const float* bias;
  char* workspace;
  int64_t lane_bytes;
};
int xdeepfm_call(void* ctx, int lane, const void* idx_dev, int idx_bits, int64_t nb, float* logits_dev,
                 int32_t* status_dev, void* stream) {
  const XdfmCall& c = *static_cast<const XdfmCall*>(ctx);
  return trs_xdeepfm_forward(idx_dev, idx_bits, c.offsets, nb, c.fields, c.w_feat, c.w_emb, c.rows, c.embed, c.cin_w,
                             c.cin_scale, c.cin_shift, c.cin_sizes, c.cin_layers, c.cin_is_direct, c.cin_act, c.cin_fc_w,
                             c.cin_fc_b, c.mlp_dims, c.mlp_layers, c.mlp_w, c.mlp_b, c.mlp_act, c.bias, logits_dev,
                             c.workspace + (size_t)lane * c.lane_bytes, c.lane_bytes, status_dev, stream);
}
struct FfmCall {
  const int64_t* offsets;
  int fields;
  const float* w_feat;
  const float* const* tables;
  const float* packed;
  int64_t rows;
  int embed;
  const float* bias;
};
int ffm_call(void* ctx, int, const void* idx_dev, int idx_bits, int64_t nb, float* logits_dev, int32_t* status_dev,
             void* stream) {
  const FfmCall& c = *static_cast<const FfmCall*>(ctx);
  if (c.packed != nullptr)
    return trs_ffm_model_forward_interleaved(idx_dev, idx_bits, c.offsets, nb, c.fields, c.packed, c.rows, c.embed, c.bias,
                                             logits_dev, status_dev, stream);
  return trs_ffm_model_forward(idx_dev, idx_bits, c.offsets, nb, c.fields, c.w_feat, c.tables, c.rows, c.embed, c.bias,
                               logits_dev, status_dev, stream);
}
}  // namespace

// Enqueues one batch on a free slot.  On any failure after the first enqueue the slot's streams are drained so that
// the slot is reusable.  `rows` < 2^31 (or 0 = unknown) gates the host-side int64 -> int32 narrowing.
static int session_submit(trs_session* s, const void* idx_host, int idx_bits, int64_t batch, int fields,
                          trs_forward_fn fn, void* ctx, int64_t rows, float* logits_host, int64_t* ticket) {
  TRS_REQUIRE(s && idx_host && logits_host && ticket && fn, "trs_session_submit: null pointer");
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
  // int64 indices of a non-trivial batch are narrowed on the host when the session has a narrowing pool
  const bool narrow = idx_bits == 64 && s->narrow != nullptr && s->chunks <= NarrowPool::kMaxChunks && rows > 0 &&
                      rows < (int64_t(1) << 31) && batch * fields >= 4096;
  if (narrow) idx_bits = 32;
  const size_t isz = idx_bits / 8;
  const bool src_pinned = !narrow && is_pinned(idx_host);
  const bool dst_pinned = is_pinned(logits_host);
  if (!dst_pinned) sl.user_logits = logits_host;
  int rc = TRS_OK;
  cudaError_t e = cudaSuccess;
  // the slot's private streams are not ordered against the caller's own stream: whatever it enqueued there before
  // this call (packing the table, uploading offsets, an optimizer step, load_state_dict) is waited for on the device
  if (s->has_producer) {
    e = cudaEventRecord(s->producer_ready, s->producer);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(sl.streams[0], s->producer_ready, 0);
  }
  if (e == cudaSuccess) e = cudaMemsetAsync(sl.status_dev, 0, TRS_STATUS_WORDS * sizeof(int32_t), sl.streams[0]);
  if (e == cudaSuccess) e = cudaEventRecord(sl.forked, sl.streams[0]);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(sl.streams[1], sl.forked, 0);
  const int chunks = (int)(batch < s->chunks ? batch : s->chunks);
  const int64_t per = (((batch + chunks - 1) / chunks + 15) / 16) * 16;  // slices start on 16-byte aligned indices
  if (narrow)
    s->narrow->start(static_cast<const int64_t*>(idx_host), static_cast<int32_t*>(sl.idx_pinned), batch * fields,
                     per * fields, chunks);
  for (int c = 0; c < chunks && rc == TRS_OK && e == cudaSuccess; ++c) {
    const int64_t b0 = c * per;
    const int64_t nb = batch - b0 < per ? batch - b0 : per;
    if (nb <= 0) break;
    cudaStream_t st = sl.streams[c & 1];
    const size_t off = (size_t)b0 * fields * isz, bytes = (size_t)nb * fields * isz;
    const char* src = static_cast<const char*>(idx_host) + off;
    if (narrow) {
      s->narrow->finish_chunk(c);
      src = static_cast<const char*>(sl.idx_pinned) + off;
    } else if (!src_pinned) {
      memcpy(static_cast<char*>(sl.idx_pinned) + off, src, bytes);
      src = static_cast<const char*>(sl.idx_pinned) + off;
    }
    e = cudaMemcpyAsync(static_cast<char*>(sl.idx_dev) + off, src, bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) break;
    rc = fn(ctx, static_cast<int>(slp - s->slots) * 2 + (c & 1), static_cast<char*>(sl.idx_dev) + off, idx_bits, nb,
            sl.logits_dev + b0, sl.status_dev, st);
    if (rc != TRS_OK) break;
    float* dst = dst_pinned ? logits_host + b0 : sl.logits_pinned + b0;
    e = cudaMemcpyAsync(dst, sl.logits_dev + b0, (size_t)nb * sizeof(float), cudaMemcpyDeviceToHost, st);
  }
  if (narrow) s->narrow->quiesce();
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

extern "C" int trs_session_set_producer_stream(trs_session* s, void* stream, int enabled) {
  TRS_REQUIRE(s, "trs_session_set_producer_stream: null session");
  s->producer = static_cast<cudaStream_t>(stream);
  s->has_producer = enabled != 0;
  return TRS_OK;
}

extern "C" int trs_session_lanes(void) { return 2 * kSlots; }

extern "C" int trs_session_submit_fm(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets,
                                     int64_t batch, int fields, const float* w_feat, const float* w_emb,
                                     const float* packed, int64_t rows, int embed, const float* bias,
                                     float* logits_host, int64_t* ticket) {
  TRS_REQUIRE(packed || (w_feat && w_emb), "trs_session_submit_fm: need the packed table or both tables");
  FmCall c{offsets, fields, w_feat, w_emb, packed, bias, rows, embed};
  return session_submit(s, idx_host, idx_bits, batch, fields, fm_call, &c, rows, logits_host, ticket);
}

extern "C" int trs_session_submit_dcn(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets,
                                      int64_t batch, int fields, const float* w_emb, int64_t rows, int embed,
                                      const float* cross_w, const float* cross_b, int cross_layers,
                                      const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                      const float* const* mlp_b, int activation, const float* fc_w, const float* fc_b,
                                      float* logits_host, int64_t* ticket) {
  DcnCall c{offsets, fields, w_emb, rows, embed, cross_w, cross_b, cross_layers, mlp_dims, mlp_layers, mlp_w, mlp_b,
            activation, fc_w, fc_b};
  return session_submit(s, idx_host, idx_bits, batch, fields, dcn_call, &c, rows, logits_host, ticket);
}

extern "C" int trs_session_submit_xdeepfm(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets,
                                          int64_t batch, int fields, const float* w_feat, const float* w_emb,
                                          int64_t rows, int embed, const float* const* cin_w,
                                          const float* const* cin_scale, const float* const* cin_shift,
                                          const int* cin_layer_sizes, int cin_layers, int cin_is_direct,
                                          int cin_activation, const float* cin_fc_w, const float* cin_fc_b,
                                          const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                          const float* const* mlp_b, int mlp_activation, const float* bias,
                                          void* workspace, int64_t workspace_bytes, float* logits_host,
                                          int64_t* ticket) {
  TRS_REQUIRE(s && workspace, "trs_session_submit_xdeepfm: null session / workspace");
  // one workspace slice per lane: slices on different lanes run concurrently
  const int64_t lane_bytes = (workspace_bytes / (2 * kSlots)) & ~int64_t(255);
  const int64_t chunks = batch < s->chunks ? (batch > 0 ? batch : 1) : s->chunks;
  const int64_t per = (((batch + chunks - 1) / chunks + 15) / 16) * 16;
  const int64_t need = trs_xdeepfm_workspace_bytes(per, fields, embed, cin_layer_sizes, cin_layers, cin_is_direct);
  TRS_REQUIRE(need >= 0 && lane_bytes >= need,
              "trs_session_submit_xdeepfm: workspace too small: %lld bytes per lane x %d lanes needed, %lld given",
              (long long)need, 2 * kSlots, (long long)workspace_bytes);
  XdfmCall c{offsets, fields, w_feat, w_emb, rows, embed, cin_w, cin_scale, cin_shift, cin_layer_sizes, cin_layers,
             cin_is_direct, cin_activation, cin_fc_w, cin_fc_b, mlp_dims, mlp_layers, mlp_w, mlp_b, mlp_activation, bias,
             static_cast<char*>(workspace), lane_bytes};
  return session_submit(s, idx_host, idx_bits, batch, fields, xdeepfm_call, &c, rows, logits_host, ticket);
}

extern "C" int trs_session_submit_ffm(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets,
                                      int64_t batch, int fields, const float* w_feat, const float* const* tables,
                                      const float* packed, int64_t rows, int embed, const float* bias,
                                      float* logits_host, int64_t* ticket) {
  TRS_REQUIRE(packed || (w_feat && tables), "trs_session_submit_ffm: need the interleaved shadow or the tables");
  FfmCall c{offsets, fields, w_feat, tables, packed, rows, embed, bias};
  return session_submit(s, idx_host, idx_bits, batch, fields, ffm_call, &c, rows, logits_host, ticket);
}

extern "C" int trs_session_submit_fn(trs_session* s, const void* idx_host, int idx_bits, int64_t batch, int fields,
                                     trs_forward_fn fn, void* ctx, int64_t rows, float* logits_host, int64_t* ticket) {
  return session_submit(s, idx_host, idx_bits, batch, fields, fn, ctx, rows, logits_host, ticket);
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
  DeepFmCall c{offsets, fields, w_feat, w_emb, nullptr, nullptr, rows, embed, 0, mlp_dims, mlp_layers, mlp_w, mlp_b, activation};
  return session_submit(s, idx_host, idx_bits, batch, fields, deepfm_call, &c, rows, logits_host, ticket);
}

extern "C" int trs_session_submit_deepfm_packed(trs_session* s, const void* idx_host, int idx_bits,
                                                const int64_t* offsets, int64_t batch, int fields,
                                                const float* packed, int64_t rows, const int* mlp_dims,
                                                int mlp_layers, const float* const* mlp_w,
                                                const float* const* mlp_b, int activation, float* logits_host,
                                                int64_t* ticket) {
  TRS_REQUIRE(packed, "trs_session_submit_deepfm_packed: null packed table");
  DeepFmCall c{offsets, fields, nullptr, nullptr, packed, nullptr, rows, 16, 0, mlp_dims, mlp_layers, mlp_w, mlp_b, activation};
  return session_submit(s, idx_host, idx_bits, batch, fields, deepfm_call, &c, rows, logits_host, ticket);
}

extern "C" int trs_session_submit_deepfm_tc(trs_session* s, const void* idx_host, int idx_bits, const int64_t* offsets,
                                            int64_t batch, int fields, const float* packed, int64_t rows,
                                            const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                            const float* const* mlp_b, int activation, const float* workspace,
                                            int variant, float* logits_host, int64_t* ticket) {
  TRS_REQUIRE(packed && workspace, "trs_session_submit_deepfm_tc: null packed table / workspace");
  DeepFmCall c{offsets, fields, nullptr, nullptr, packed, workspace, rows, 16, variant, mlp_dims, mlp_layers, mlp_w, mlp_b, activation};
  return session_submit(s, idx_host, idx_bits, batch, fields, deepfm_call, &c, rows, logits_host, ticket);
}

extern "C" int trs_session_deepfm_forward_host(trs_session* s, const void* idx_host, int idx_bits,
                                               const int64_t* offsets, int64_t batch, int fields,
                                               const float* w_feat, const float* w_emb, int64_t rows, int embed,
                                               const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                               const float* const* mlp_b, int activation, float* logits_host,
                                               int64_t* oob_count) {
  int64_t ticket = 0;
  int rc = trs_session_submit_deepfm(s, idx_host, idx_bits, offsets, batch, fields, w_feat, w_emb, rows, embed, mlp_dims,
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
  int64_t ticket = 0;
  int rc = trs_session_submit_deepfm_packed(s, idx_host, idx_bits, offsets, batch, fields, packed, rows, mlp_dims,
                                            mlp_layers, mlp_w, mlp_b, activation, logits_host, &ticket);
  if (rc != TRS_OK) return rc;
  return trs_session_wait(s, ticket, oob_count);
}
