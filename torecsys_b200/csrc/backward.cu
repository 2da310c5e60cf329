// First dedicated BACKWARD kernels of the hot path: the two ops every a12 model's training step goes
// through -- the embedding lookup and the FM layer.  (Every other layer still back-propagates by recomputing its
// formula with differentiable torch CUDA ops, torecsys_b200/autograd.py.)
//
//   trs_embedding_grad   d weight of MultiIndicesEmbedding / SingleIndexEmbedding (nn.Embedding's dense gradient,
//                        torecsys/inputs/base/multi_indices_emb.py:48, sparse=False): grad_weight[idx + off] += grad_out
//   trs_fm_backward      d x of FactorizationMachineLayer (factorization_machine.py:46-81):
//                        grad_x[b,n,e] = grad_out[b,e] * (sum_m x[b,m,e] - x[b,n,e])
#include "common.cuh"

namespace trs {
namespace {

// One work item = one 16-byte chunk of one looked-up row: grad_weight[row][chunk] += grad_out[pos][chunk] with ONE
// vector reduction (red.global.add.v4.f32, sm_90+): the rows of a batch are random, collisions are rare and resolved
// by the memory system; the sum over colliding lookups is order-dependent in the last bit, like torch's index_add_.
template <int IdxBits>
__global__ void __launch_bounds__(256) embedding_grad_vec_kernel(const float4* __restrict__ grad_out,
                                                                 const void* __restrict__ idx,
                                                                 const int64_t* __restrict__ offsets, int64_t rows,
                                                                 uint32_t chunks, uint32_t fields, int64_t items,
                                                                 int64_t padding_row, float* __restrict__ grad_weight) {
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = item / chunks;
    const uint32_t chunk = static_cast<uint32_t>(item - pos * chunks);
    int64_t r = load_index<IdxBits>(idx, pos);
    if (offsets != nullptr) r += __ldg(offsets + pos % fields);
    if (r < 0 || r >= rows || r == padding_row) continue;   // out-of-range lookups were reported by the forward
    const float4 g = ldg_stream_f4(grad_out + item);
    float* dst = grad_weight + (r * chunks + chunk) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                 : "memory");
  }
}

template <int IdxBits>
__global__ void __launch_bounds__(256) embedding_grad_scalar_kernel(const float* __restrict__ grad_out,
                                                                    const void* __restrict__ idx,
                                                                    const int64_t* __restrict__ offsets, int64_t rows,
                                                                    uint32_t embed, uint32_t fields, int64_t items,
                                                                    int64_t padding_row,
                                                                    float* __restrict__ grad_weight) {
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = item / embed;
    const uint32_t e = static_cast<uint32_t>(item - pos * embed);
    int64_t r = load_index<IdxBits>(idx, pos);
    if (offsets != nullptr) r += __ldg(offsets + pos % fields);
    if (r < 0 || r >= rows || r == padding_row) continue;
    atomicAdd(grad_weight + r * embed + e, ldg_stream_f1(grad_out + item));
  }
}

// FM backward: one warp per sample; pass 1 sums the fields per embedding component, pass 2 writes the gradient.
__global__ void __launch_bounds__(256) fm_backward_kernel(const float* __restrict__ x, const float* __restrict__ grad_out,
                                                          int64_t batch, int fields, int embed,
                                                          float* __restrict__ grad_x) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int tile = fields * embed;
  for (int64_t b = warp0; b < batch; b += warps) {
    const float* xb = x + b * tile;
    float* gb = grad_x + b * tile;
    for (int e0 = 0; e0 < embed; e0 += 32) {   // 32 embedding components per pass (one pass for E <= 32)
      const int e = e0 + lane;
      float s = 0.f;
      if (e < embed)
        for (int n = 0; n < fields; ++n) s += __ldg(xb + n * embed + e);
      if (e < embed) {
        const float g = __ldg(grad_out + b * embed + e);
        for (int n = 0; n < fields; ++n) gb[n * embed + e] = g * (s - __ldg(xb + n * embed + e));
      }
    }
  }
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_embedding_grad(const float* grad_out, const void* idx, int idx_bits, const int64_t* offsets,
                                  int64_t batch, int fields, int64_t rows, int embed, int64_t padding_row,
                                  float* grad_weight, void* stream) {
  TRS_REQUIRE(grad_out && idx && grad_weight, "trs_embedding_grad: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_embedding_grad: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && rows > 0 && embed > 0, "trs_embedding_grad: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t lookups = batch * fields;
  if ((embed & 3) == 0 && aligned16(grad_out) && aligned16(grad_weight)) {
    const uint32_t chunks = embed / 4;
    const int64_t items = lookups * chunks;
    const int grid = grid_for(items, 256, 16);
    if (idx_bits == 64)
      embedding_grad_vec_kernel<64><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(grad_out), idx, offsets, rows,
                                                         chunks, fields, items, padding_row, grad_weight);
    else
      embedding_grad_vec_kernel<32><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(grad_out), idx, offsets, rows,
                                                         chunks, fields, items, padding_row, grad_weight);
    return check_launch("embedding_grad_vec_kernel");
  }
  const int64_t items = lookups * embed;
  const int grid = grid_for(items, 256, 16);
  if (idx_bits == 64)
    embedding_grad_scalar_kernel<64><<<grid, 256, 0, s>>>(grad_out, idx, offsets, rows, embed, fields, items, padding_row,
                                                          grad_weight);
  else
    embedding_grad_scalar_kernel<32><<<grid, 256, 0, s>>>(grad_out, idx, offsets, rows, embed, fields, items, padding_row,
                                                          grad_weight);
  return check_launch("embedding_grad_scalar_kernel");
}

extern "C" int trs_fm_backward(const float* x, const float* grad_out, int64_t batch, int fields, int embed,
                               float* grad_x, void* stream) {
  TRS_REQUIRE(x && grad_out && grad_x, "trs_fm_backward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 0 && embed > 0, "trs_fm_backward: bad sizes");
  if (batch == 0) return TRS_OK;
  fm_backward_kernel<<<grid_for(batch * 32, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, grad_out, batch,
                                                                                                  fields, embed, grad_x);
  return check_launch("fm_backward_kernel");
}
