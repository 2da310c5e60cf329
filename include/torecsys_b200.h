/*
 * torecsys_b200 -- C ABI of the B200-native CTR forward hot path.
 *
 * The reference (p768lwy3/torecsys) is pure Python on torch ATen and has no FFI of its own; the boundary a
 * maintainer would bind is therefore "one C entry point per reference forward() on the hot path"
 * (SURVEY.md section 8a rows a1..a12).  Every entry point below cites the reference forward it replaces.
 * INTEGRATION.md shows the ctypes stub that goes into the reference module for each of them.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer on the current CUDA device unless the
 *     parameter name ends in `_host` (then it is a host pointer, pinned or pageable);
 *   - float tensors are contiguous row-major fp32, indices are contiguous row-major int64 (`idx_bits` = 64) or
 *     int32 (`idx_bits` = 32; the reference accepts both, multi_indices_emb.py:104 promotes);
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream); calls are asynchronous and
 *     CUDA-graph capturable (no host synchronisation, no allocation) unless stated otherwise;
 *   - `status` (may be NULL) points to `TRS_STATUS_WORDS` int32 words in device memory that kernels update when
 *     they meet an out-of-range row id: word0 += 1 per offending lookup, word1 = flat position (b*N+n) of one
 *     offender.  An offending lookup reads as a row of zeros, never out of bounds.  The Python layer turns a
 *     non-zero word0 into IndexError (the reference raises IndexError from nn.Embedding on CPU);
 *   - return value: TRS_OK (0) or a negative TRS_ERR_* code; `trs_last_error()` gives the message (thread local).
 *     Argument errors are reported before anything is launched.
 */
#ifndef TORECSYS_B200_H_
#define TORECSYS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRS_OK 0
#define TRS_ERR_INVALID_ARGUMENT (-1)
#define TRS_ERR_UNSUPPORTED (-2)
#define TRS_ERR_CUDA (-3)

#define TRS_STATUS_WORDS 2

/* activation ids (nn.ReLU / None / nn.Sigmoid / nn.Tanh instances of the reference constructors) */
#define TRS_ACT_NONE 0
#define TRS_ACT_RELU 1
#define TRS_ACT_SIGMOID 2
#define TRS_ACT_TANH 3

/* ---- library ---------------------------------------------------------------------------------------------- */
const char* trs_version(void);
const char* trs_last_error(void);
/* compute capability of the current device as major*10+minor (100 on B200); negative on error */
int trs_device_arch(void);

/* ---- a1/a2: embedding row gather ----------------------------------------------------------------------------
 * Replaces SingleIndexEmbedding.forward (torecsys/inputs/base/single_index_emb.py:45-59; offsets = NULL, N = 1)
 * and MultiIndicesEmbedding.forward (torecsys/inputs/base/multi_indices_emb.py:92-112):
 *     out[b, n, :] = weight[idx[b, n] + offsets[n], :]            bit-exact copy of fp32 rows
 * weight (rows, embed) ; idx (batch, fields) ; offsets (fields) int64 or NULL ; out (batch, fields, embed).
 * `flatten` of the reference is a view of the same bytes and is done by the caller. */
int trs_embedding_gather(const float* weight, int64_t rows, int embed,
                         const void* idx, int idx_bits, const int64_t* offsets,
                         int64_t batch, int fields, float* out, int32_t* status, void* stream);

/* ---- a4: index-column concatenation --------------------------------------------------------------------------------
 * Replaces the `torch.cat(inputs, dim=1)` of Inputs.forward (torecsys/inputs/inputs.py:76-81): the batch dict holds one
 * (batch,) or (batch, w) index tensor per feature; the embedding's (batch, N) index matrix is their concatenation.
 *     out[b, off_c + j] = columns[c][b * widths[c] + j]            off_c = widths[0] + ... + widths[c-1]
 * columns / widths: HOST arrays of length ncols (DEVICE pointers / elements per row); all columns int64 or all int32. */
int trs_index_concat(const void* const* columns, const int* widths, int ncols, int idx_bits, int64_t batch,
                     void* out, void* stream);

/* ---- a3: field-aware gather -----------------------------------------------------------------------------------
 * Replaces MultiIndicesFieldAwareEmbedding.forward (torecsys/inputs/base/multi_indices_field_aware_emb.py:90-111):
 *     out[b, t*N + f, :] = tables[t][idx[b, f] + offsets[f], :]   for t, f in [0, N)
 * tables: device array of N device pointers, each (rows, embed); out (batch, N*N, embed). */
int trs_embedding_gather_field_aware(const float* const* tables, int64_t rows, int embed,
                                     const void* idx, int idx_bits, const int64_t* offsets,
                                     int64_t batch, int fields, float* out, int32_t* status, void* stream);

/* ---- a5: FM second order ----------------------------------------------------------------------------------------
 * Replaces FactorizationMachineLayer.forward (torecsys/layers/ctr/factorization_machine.py:46-81), eval mode:
 *     out[b, e] = 0.5 * ((sum_n x[b,n,e])^2 - sum_n x[b,n,e]^2)      x (batch, fields, embed) -> out (batch, embed) */
int trs_fm_forward(const float* x, int64_t batch, int fields, int embed, float* out, void* stream);

/* ---- 8f-2: backward kernels of the ops the a12 models' training steps go through ----------------------------------------
 * trs_embedding_grad: the dense weight gradient of nn.Embedding behind MultiIndicesEmbedding / SingleIndexEmbedding
 * (multi_indices_emb.py:48, single_index_emb.py:41; sparse=False):
 *     grad_weight[idx[b,n] + offsets[n], :] += grad_out[b, n, :]       (grad_weight is NOT zeroed here)
 * padding_row (-1 = none) is skipped like nn.Embedding's padding_idx; out-of-range lookups are skipped (the forward
 * reported them).  Vector reductions (red.global.add.v4.f32): colliding rows add in an unspecified order.
 * trs_fm_backward: grad_x[b,n,e] = grad_out[b,e] * (sum_m x[b,m,e] - x[b,n,e])  (factorization_machine.py:46-81). */
int trs_embedding_grad(const float* grad_out, const void* idx, int idx_bits, const int64_t* offsets,
                       int64_t batch, int fields, int64_t rows, int embed, int64_t padding_row,
                       float* grad_weight, void* stream);
int trs_fm_backward(const float* x, const float* grad_out, int64_t batch, int fields, int embed,
                    float* grad_x, void* stream);

/* Sparse form of the embedding gradient: what nn.Embedding(sparse=True) hands autograd (multi_indices_emb.py:48 forwards
 * **kwargs to nn.Embedding, so MultiIndicesEmbedding(..., sparse=True) is part of the reference's interface).
 * trs_embedding_rows: out_rows[b*N + n] = idx[b,n] + offsets[n] (int64): the COO indices, in lookup order; the COO values
 * are grad_out viewed (B*N, E).  trs_embedding_grad_segments coalesces that tensor once its row ids have been sorted:
 *     out_values[s, :] = sum over m in [starts[s], starts[s+1]) of grad_out[perm[m], :]      (summed in sorted order)
 * perm (lookups) = permutation of the sort, starts (segments + 1) = first sorted position of each distinct row, closed by
 * the number of lookups.  Deterministic, unlike trs_embedding_grad's reductions. */
int trs_embedding_rows(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                       int64_t* out_rows, void* stream);
int trs_embedding_grad_segments(const float* grad_out, const int64_t* perm, const int64_t* starts, int64_t segments,
                                int embed, float* out_values, void* stream);

/* trs_ffm_backward: gradient of FieldAwareFactorizationMachineLayer.forward (field_aware_factorization_machine.py:50-94)
 *     grad_v[b, a*N + c, :] = grad_out[b, p(min(a,c), max(a,c)), :] * v[b, c*N + a, :]   (a != c), 0 on the diagonal
 * trs_ipn_backward: gradient of InnerProductNetworkLayer.forward (inner_product_network.py:54-79)
 *     grad_x[b, i, :] = sum_{j != i} grad_out[b, p(i,j)] * x[b, j, :]
 * trs_cross_backward: gradients of CrossNetworkLayer.forward (cross_network.py:52-87) for x, every W_l and b_l.  The
 * chain starts from h_0 = x.detach() as upstream does (:65): x receives gradient through the `x * (...) + x` terms
 * only.  grad_weights (layers, E, E) and grad_biases (layers, E) are OVERWRITTEN (zeroed on the stream, then
 * accumulated with float atomics, one add per CTA).  embed 8, 16, 32 or 64 (TRS_ERR_UNSUPPORTED otherwise). */
int trs_ffm_backward(const float* v, const float* grad_out, int64_t batch, int fields, int embed,
                     float* grad_v, void* stream);
int trs_ipn_backward(const float* x, const float* grad_out, int64_t batch, int fields, int embed,
                     float* grad_x, void* stream);
int trs_cross_backward(const float* x, const float* weights, const float* biases, const float* grad_out,
                       int layers, int64_t rows, int embed, float* grad_x, float* grad_weights,
                       float* grad_biases, void* stream);

/* ---- a6: field-aware FM ------------------------------------------------------------------------------------------
 * Replaces FieldAwareFactorizationMachineLayer.forward
 * (torecsys/layers/ctr/field_aware_factorization_machine.py:50-94), eval mode:
 *     out[b, p, :] = v[b, i*N + j, :] * v[b, j*N + i, :]   for pairs p = (i<j) in lexicographic order
 * v (batch, N*N, embed) -> out (batch, N(N-1)/2, embed). */
int trs_ffm_forward(const float* v, int64_t batch, int fields, int embed, float* out, void* stream);

/* ---- a9: inner product network -----------------------------------------------------------------------------------
 * Replaces InnerProductNetworkLayer.forward (torecsys/layers/ctr/inner_product_network.py:54-79):
 *     out[b, p] = sum_e x[b,i,e] * x[b,j,e]        x (batch, fields, embed) -> out (batch, N(N-1)/2) */
int trs_ipn_forward(const float* x, int64_t batch, int fields, int embed, float* out, void* stream);

/* ---- a10: bilinear interaction -----------------------------------------------------------------------------------
 * Replaces BilinearInteractionLayer.forward (torecsys/layers/ctr/bilinear_interaction.py:230-255) with
 * FieldAllTypeBilinear.forward (:72-76; each_type = 0, weight (E,E), bias (E)) or
 * FieldEachTypeBilinear.forward (:144-149; each_type = 1, weight (P,E,E), bias (P,E)):
 *     out[b, p, :] = (x[b,i,:] @ W_(p)) * x[b,j,:] + bias_(p)       bias may be NULL
 * x (batch, fields, embed) -> out (batch, P, embed). */
int trs_bilinear_forward(const float* x, const float* weight, const float* bias, int each_type,
                         int64_t batch, int fields, int embed, float* out, void* stream);
/* The same forward writing sample b's (P, embed) block at out + b * out_stride (floats): lets a caller that concatenates
 * several interaction outputs per sample (FiBiNET: torch.cat([emb_interaction, senet_interaction], dim='N'),
 * feature_importance_and_bilinear_feature_interaction_network.py forward) have them written in place.  Tensor-core
 * kernel only: embed 8 / 16 / 32, 16-byte aligned x / out, out_stride a multiple of 4 (TRS_ERR_UNSUPPORTED otherwise). */
int trs_bilinear_forward_strided(const float* x, const float* weight, const float* bias, int each_type,
                                 int64_t batch, int fields, int embed, int64_t out_stride, float* out,
                                 void* stream);

/* trs_bilinear_backward: gradients of BilinearInteractionLayer.forward (bilinear_interaction.py:230-255) for x, the
 * weight and the bias, given grad_out (batch, P, embed).  With t[b,p,:] = grad_out[b,p,:] * x[b,j,:]:
 *     grad_x[b,i,:] += W_(p) t[b,p,:]        grad_x[b,j,:] += grad_out[b,p,:] * (x[b,i,:] @ W_(p))
 *     grad_weight_(p) += sum_b x[b,i,:]^T t[b,p,:]        grad_bias_(p) += sum_b grad_out[b,p,:]
 * grad_x (batch, fields, embed), grad_weight ((P,)E,E) and grad_bias ((P,)E; may be NULL) are OVERWRITTEN (the
 * parameter gradients are zeroed on the stream, then accumulated with float atomics, one add per CTA).
 * embed 8, 16 or 32 and 2*16*fields*embed*4 bytes of shared memory (TRS_ERR_UNSUPPORTED otherwise). */
int trs_bilinear_backward(const float* x, const float* weight, const float* grad_out, int each_type,
                          int64_t batch, int fields, int embed, float* grad_x, float* grad_weight,
                          float* grad_bias, void* stream);

/* ---- a11: attentional FM -----------------------------------------------------------------------------------------
 * Replaces AttentionalFactorizationMachineLayer.forward
 * (torecsys/layers/ctr/attentional_factorization_machine.py:86-120), eval mode (dropouts are identity):
 *     prod_p = x_i * x_j ; s = softmax_p( w2 . relu(W1 prod_p + b1) + b2 ) ; out[b,:] = sum_p s_p prod_p
 * W1 (attn, embed), b1 (attn), w2 (attn) [= OutProj.weight (1, attn)], b2 (1).
 * out (batch, embed), scores (batch, P) [= the reference's (B, P, 1)]. */
int trs_afm_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                    int64_t batch, int fields, int embed, int attn, float* out, float* scores, void* stream);

/* trs_afm_backward: gradients of AttentionalFactorizationMachineLayer.forward (:86-120, eval mode) for x, W1, b1, w2
 * and b2, given the attention `scores` (batch, P) the forward returned, grad_out (batch, embed) and, optionally,
 * grad_scores (batch, P; NULL = zero).  With da_p = <grad_out, prod_p> + grad_scores_p, c = sum_q s_q da_q,
 * ds_p = s_p (da_p - c), dh_p = w2 * ds_p * [W1 prod_p + b1 > 0], dprod_p = s_p grad_out + W1^T dh_p:
 *     grad_x[b,i,:] += dprod_p * x[b,j,:]   grad_x[b,j,:] += dprod_p * x[b,i,:]
 *     grad_w1 += dh_p (x) prod_p   grad_b1 += dh_p   grad_w2 += ds_p relu(W1 prod_p + b1)   grad_b2 += ds_p
 * Every output is OVERWRITTEN (parameter gradients zeroed on the stream, then one float atomic per word and CTA).
 * (embed, attn) in {(8, 8|16|32), (16, 8|16), (32, 8)} -- trs_afm_backward_supported() -- and
 * 2*16*fields*embed*4 bytes of shared memory; TRS_ERR_UNSUPPORTED otherwise. */
int trs_afm_backward_supported(int embed, int attn);
int trs_afm_backward(const float* x, const float* w1, const float* b1, const float* w2, const float* scores,
                     const float* grad_out, const float* grad_scores, int64_t batch, int fields, int embed,
                     int attn, float* grad_x, float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2,
                     void* stream);

/* ---- 8f-3: outer product network (PNN) ------------------------------------------------------------------------------
 * Replaces OuterProductNetworkLayer.forward (torecsys/layers/ctr/outer_product_network.py:80-131), pairs p = (i<j)
 * in lexicographic order:
 *     TRS_OPN_MAT  kernel (E, P, E):  out[b,p] = sum_h x[b,j,h] * ( sum_e kernel[h,p,e] * x[b,i,e] )
 *     TRS_OPN_VEC  kernel (1, P, E):  out[b,p] = sum_e x[b,i,e] * x[b,j,e] * kernel[p,e]
 *     TRS_OPN_NUM  kernel (1, P, 1):  out[b,p] = sum_e x[b,i,e] * x[b,j,e] * kernel[p]
 * x (batch, fields, embed) -> out (batch, P). */
#define TRS_OPN_MAT 0
#define TRS_OPN_VEC 1
#define TRS_OPN_NUM 2
int trs_opn_forward(const float* x, const float* kernel, int kernel_type, int64_t batch, int fields, int embed,
                    float* out, void* stream);

/* ---- 8f-3: squeeze-and-excitation / compose-excitation network (FiBiNET, FAT-DeepFFM) -----------------------------------
 * Replaces ComposeExcitationNetworkLayer.forward (torecsys/layers/ctr/compose_excitation_network.py:72-109):
 *     pooled[b,m] = mean_e x[b,m,e]                                   (nn.AdaptiveAvgPool1d(1))
 *     a[b,:]      = act( w2 . act( w1 . pooled[b,:] + b1 ) + b2 )     ReductionLinear w1 (R, M), AdditionLinear w2 (M, R)
 *     out[b,m,:]  = x[b,m,:] * a[b,m]
 * x (batch, rows_per_sample = M, embed): M = num_fields, or num_fields^2 when the layer is `squared`.
 * workspace: device scratch of at least trs_senet_workspace_bytes(batch, M) bytes. */
int64_t trs_senet_workspace_bytes(int64_t batch, int rows_per_sample);
int trs_senet_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                      int activation, int64_t batch, int rows_per_sample, int embed, int reduced,
                      float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- a7: cross network -------------------------------------------------------------------------------------------
 * Replaces CrossNetworkLayer.forward (torecsys/layers/ctr/cross_network.py:52-87):
 *     h_0 = x ; h_{l+1} = x * (h_l @ W_l^T + b_l) + x    per (b, n) row, W_l (E, E), b_l (E)
 * weights (layers, E, E), biases (layers, E) packed contiguously; x (rows, embed) with rows = batch*fields. */
int trs_cross_forward(const float* x, const float* weights, const float* biases, int layers,
                      int64_t rows, int embed, float* out, void* stream);

/* Experimental: the same cross network with the whole layer chain in TENSOR MEMORY (tcgen05.mma, A operand in TMEM,
 * 3xTF32).  Same results; measured slower than trs_cross_forward's mma.sync chain at embed = 32 (the per-instruction
 * cost of a narrow tcgen05.mma), so it is not on the default path.  embed 16 or 32, rows >= 640
 * (TRS_ERR_UNSUPPORTED otherwise). */
int trs_cross_forward_tc5(const float* x, const float* weights, const float* biases, int layers,
                          int64_t rows, int embed, float* out, void* stream);

/* ---- a8: compress interaction network ------------------------------------------------------------------------------
 * Replaces CompressInteractionNetworkLayer.forward (torecsys/layers/ctr/compress_interaction_network.py:85-184), eval:
 *     per layer l:  z[b, xf*H + y, e] = x[b,xf,e] * h[b,y,e]
 *                   o = act( scale_l * (conv_w_l z) + shift_l )      (Conv1d bias and eval-BatchNorm folded by the
 *                                                                     caller into a per-channel scale/shift)
 *                   direct: d = h = o ; else: d = o[:, :H_l], h = o[:, H_l:]   (every layer)
 *     out = fc_w . ( sum_e cat_l d ) + fc_b
 * conv_w[l] (C_l, N*H_{l-1}), scale[l]/shift[l] (C_l) are HOST arrays of DEVICE pointers (length `layers`);
 * layer_sizes (HOST) = H_1..H_L; C_l = H_l if is_direct else 2*H_l.
 * workspace: device scratch of at least trs_cin_workspace_bytes(...) bytes.
 * x (batch, fields, embed) -> out (batch, out_features). */
int64_t trs_cin_workspace_bytes(int64_t batch, int fields, int embed, const int* layer_sizes, int layers,
                                int is_direct);
int trs_cin_forward(const float* x, const float* const* conv_w, const float* const* scale,
                    const float* const* shift, const int* layer_sizes, int layers, int is_direct, int activation,
                    const float* fc_w, const float* fc_b, int out_features,
                    int64_t batch, int fields, int embed, float* out, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* ---- MLP (adjacent; used inside the fused model entry points and by DNNLayer) ---------------------------------------
 * Replaces MultilayerPerceptionLayer.forward (torecsys/layers/ctr/multilayer_perceptron.py:63-84), eval mode:
 *     h = act(h @ W_i^T + b_i) for the hidden Linears, then LinearOutput without activation, on the last dim.
 * The parameter pack used by every entry point that embeds an MLP:
 *   dims (HOST, int[layers+1]) = in, hidden..., out ; weights[i] (dims[i+1], dims[i]), biases[i] (dims[i+1]) are HOST
 *   arrays of DEVICE pointers of length `layers` (the last one is LinearOutput). */
int trs_mlp_forward(const float* x, int64_t rows, const int* dims, int layers,
                    const float* const* weights, const float* const* biases, int activation,
                    float* out, void* stream);

/* Backward of trs_mlp_forward for the NARROW MLPs of the CTR models on this path (DeepFM / xDeepFM / NFM / FNN deep
 * branches, DCN's per-field MLP): dims[0] % 4 == 0 and every other width <= 32 (trs_mlp_backward_supported; the input
 * tile and grad_W_1 must fit shared memory).  One kernel (csrc/mlp_bwd.cu): forward recomputed per 32-row tile,
 * layers walked backwards in FP32, parameter gradients accumulated in shared memory and added to global memory once per
 * CTA.  grad_weights[l] (dims[l+1], dims[l]) and grad_biases[l] (dims[l+1]) are OVERWRITTEN (zeroed on the stream first);
 * grad_x (rows, dims[0]) may be NULL; grad_biases may be NULL or hold NULL entries.  HOST arrays of DEVICE pointers. */
int trs_mlp_backward_supported(const int* dims, int layers);
int trs_mlp_backward(const float* x, int64_t rows, const int* dims, int layers,
                     const float* const* weights, const float* const* biases, int activation,
                     const float* grad_out, float* grad_x, float* const* grad_weights, float* const* grad_biases,
                     void* stream);

/* Backward of trs_senet_forward (compose_excitation_network.py:72-109) for x, w1, b1, w2, b2: one warp per sample,
 * rows_per_sample <= 64 and reduced <= 32 (trs_senet_backward_supported: FiBiNET's SENET; the N^2-row CEN of FAT-DeepFFM
 * is not covered).  The four parameter gradients are OVERWRITTEN (zeroed on the stream, accumulated per CTA in shared
 * memory, one float atomic per element and CTA). */
int trs_senet_backward_supported(int rows_per_sample, int reduced);
int trs_senet_backward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                       int activation, const float* grad_out, int64_t batch, int rows_per_sample, int embed, int reduced,
                       float* grad_x, float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2, void* stream);

/* ---- a12: fused model forwards, indices -> logits (the L2 boundary: Sequential.forward,
 *      torecsys/models/sequential.py:31-44 = Inputs.forward (torecsys/inputs/inputs.py:56-89) + model.forward) -------
 * Common arguments: idx (batch, fields); offsets (fields) int64; w_feat (rows, 1) = the first-order
 * MultiIndicesEmbedding(embed_size=1) table; w_emb (rows, embed); logits (batch, 1).
 */

/* FactorizationMachineModel.forward (torecsys/models/ctr/factorization_machine.py:42-71):
 *     logit = sum_n w_feat[r_n] + sum_e FM(emb)[e] (+ bias[0] when bias != NULL) */
int trs_fm_model_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                         const float* w_feat, const float* w_emb, int64_t rows, int embed,
                         const float* bias, float* logits, int32_t* status, void* stream);

/* DeepFactorizationMachineModel.forward (torecsys/models/ctr/deep_fm.py:55-110):
 *     logit = MLP(flatten(emb)) + sum_e FM(emb)[e] + sum_n w_feat[r_n]           (MLP in = fields*embed, out = 1) */
int trs_deepfm_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                       const float* w_feat, const float* w_emb, int64_t rows, int embed,
                       const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                       int activation, float* logits, int32_t* status, void* stream);

/* Packed-table variant of the DeepFM forward (embed_size = 16 only): the B200-first memory layout of this path.
 * A random row read costs one 128-byte DRAM transaction whatever its size (measured), so the two
 * lookups per field of the reference (MultiIndicesEmbedding(16) and MultiIndicesEmbedding(1) on the same row id,
 * torecsys/inputs/inputs.py:69-87) are served from ONE 128-byte aligned shadow row:
 *     packed[r] = [ w_emb[r][0..15] | w_feat[r] | 15 x 0 ]        packed (rows, 32) fp32, 128-byte aligned
 * trs_fm_pack_table builds it from the two registered tables (run again whenever they change);
 * trs_deepfm_forward_packed computes exactly what trs_deepfm_forward computes.
 * Restrictions: hidden widths 16, ReLU, fields <= 40, rows < 2^31 (TRS_ERR_UNSUPPORTED otherwise). */
int trs_fm_pack_table(const float* w_emb, const float* w_feat, int64_t rows, int embed, float* packed, void* stream);
/* FactorizationMachineModel.forward on the packed table (same result as trs_fm_model_forward; bias may be NULL) */
int trs_fm_model_forward_packed(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                const float* packed, int64_t rows, const float* bias, float* logits,
                                int32_t* status, void* stream);
int trs_deepfm_forward_packed(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                              const float* packed, int64_t rows,
                              const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                              const float* const* mlp_b, int activation, float* logits, int32_t* status,
                              void* stream);
/* Same computation with launch flags.  TRS_LAUNCH_OVERLAP_PREVIOUS launches the kernel with programmatic stream
 * serialisation (Hopper/Blackwell "programmatic dependent launch"): its CTAs start on every SM the previous kernel
 * of the stream has already left and READ their inputs at once, but write nothing (logits, status) before that
 * previous kernel has completed and its writes are visible.  The drain of batch k then overlaps the start of batch
 * k+1 -- back-to-back Sequential.forward calls (torecsys/models/sequential.py:31-44) on batches that are already
 * resident.  Contract of the flag: no kernel still running on `stream` writes idx, offsets, packed or the MLP
 * parameters of this call (work enqueued as copies/memsets is ordered as usual).  flags = 0 is exactly
 * trs_deepfm_forward_packed. */
#define TRS_LAUNCH_OVERLAP_PREVIOUS 1u
/* 1 when trs_deepfm_forward_packed[_ex] runs this deep branch on the packed table through the gathering tcgen05 layer
 * (deep_fm.py:55-110 with paper-size `deep_layer_sizes`: some layer >= 64 x 64, batch >= 1 024): the row and its
 * first-order value come out of the same 128-byte line. */
int trs_deepfm_packed_wide_supported(int fields, const int* mlp_dims, int mlp_layers, int64_t batch);
int trs_deepfm_forward_packed_ex(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                 const float* packed, int64_t rows,
                                 const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                 const float* const* mlp_b, int activation, float* logits, int32_t* status,
                                 unsigned flags, void* stream);

/* DeepFM forward on the packed table with layer 1 on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
 * accumulators in tensor memory, 3xTF32 = fp32-accurate): the same computation as trs_deepfm_forward_packed
 * (torecsys/models/ctr/deep_fm.py:55-110 behind torecsys/models/sequential.py:31-44) -- csrc/deepfm_tc5.cu.
 * W1 = mlp_w[0] is pre-split ONCE per model into the tensor-core operand layout by trs_deepfm_tc_prepare (run it
 * again whenever W1 changes); `workspace` holds trs_deepfm_tc_workspace_bytes(fields, variant) bytes of device
 * memory, 16-byte aligned.  variant 0: one CTA of 19 warps per SM (4 stages of 128 samples x 3 fields);
 * variant 1: two CTAs of 13 warps per SM (2 stages of 128 x 2), so that back-to-back launches with
 * TRS_LAUNCH_OVERLAP_PREVIOUS overlap on every SM.  Restrictions as for the packed kernel (hidden widths 16, ReLU,
 * rows < 2^31); trs_deepfm_tc_supported() answers without launching. */
int64_t trs_deepfm_tc_workspace_bytes(int fields, int variant);
int trs_deepfm_tc_supported(int fields, int embed, const int* mlp_dims, int mlp_layers, int activation, int64_t rows,
                            int variant);
int trs_deepfm_tc_prepare(int fields, const float* w1, int variant, float* workspace, void* stream);
/* The same forward on a ROW-SHARDED packed table (BASELINE.json north_star: "row-sharding the large embedding tables
 * ... only when a table exceeds one GPU's HBM"; the reference keeps ONE shared table nn.Embedding(sum(field_sizes), E),
 * torecsys/inputs/base/multi_indices_emb.py:45-57).  Row g of the packed table lives on rank g % world, at local row
 * g / world of shards[g % world]; `shards` is a HOST array of `world` device addresses as mapped in THIS process: the
 * rank's own shard in its HBM, the others peer-mapped over NVLink (CUDA IPC / symmetric memory).  The exchange of
 * looked-up vectors happens inside the kernel: the row copies (cp.async, one 80-byte request per row) read remote
 * shards directly while the tensor core works on the rows that have landed -- no staging buffer, no second kernel.
 * Bit-identical to trs_deepfm_forward_tc on the unsharded table.  world <= 8, rows / world < 2^28. */
int trs_deepfm_forward_tc_sharded(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                  const float* const* shards, int world, int64_t rows,
                                  const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                  const float* const* mlp_b, int activation, const float* workspace, int variant,
                                  float* logits, int32_t* status, unsigned flags, void* stream);
/* debug: event clocks of CTA 0's warp roles into a device buffer of 7 x 512 x 4 int64 (NULL switches it off) */
int trs_debug_tc5_trace(long long* device_buf);
int trs_deepfm_forward_tc(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                          const float* packed, int64_t rows,
                          const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                          const float* const* mlp_b, int activation, const float* workspace, int variant,
                          float* logits, int32_t* status, unsigned flags, void* stream);

/* DeepAndCrossNetworkModel.forward (torecsys/models/ctr/deep_and_cross_network.py:58-98):
 *     logit = fc( flatten( cat[ Cross(emb) (B,N,E), MLP_per_field(emb) (B,N,Od) ], dim=-1 ) )
 * cross_w (cross_layers, E, E), cross_b (cross_layers, E); MLP in = embed, out = Od; fc_w (1, N*(E+Od)), fc_b (1). */
int trs_dcn_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                    const float* w_emb, int64_t rows, int embed,
                    const float* cross_w, const float* cross_b, int cross_layers,
                    const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                    int activation, const float* fc_w, const float* fc_b,
                    float* logits, int32_t* status, void* stream);

/* XDeepFactorizationMachineModel.forward (torecsys/models/ctr/xdeep_fm.py:82-124):
 *     logit = sum_n w_feat[r_n] + CIN(emb)[0] + MLP(flatten(emb)) + bias[0]
 * CIN arguments as in trs_cin_forward with out_features = 1. */
int64_t trs_xdeepfm_workspace_bytes(int64_t batch, int fields, int embed, const int* cin_layer_sizes, int cin_layers,
                                    int cin_is_direct);
int trs_xdeepfm_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                        const float* w_feat, const float* w_emb, int64_t rows, int embed,
                        const float* const* cin_w, const float* const* cin_scale, const float* const* cin_shift,
                        const int* cin_layer_sizes, int cin_layers, int cin_is_direct, int cin_activation,
                        const float* cin_fc_w, const float* cin_fc_b,
                        const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                        int mlp_activation, const float* bias,
                        float* logits, void* workspace, int64_t workspace_bytes, int32_t* status, void* stream);

/* FieldAwareFactorizationMachineModel.forward (torecsys/models/ctr/field_aware_factorization_machine.py:39-81):
 *     logit = sum_{i<j} < tables[i][r_j], tables[j][r_i] > + sum_n w_feat[r_n] + bias[0]
 * (= sum over p,e of FFM(field_emb)[b,p,e] with field_emb[b, t*N+f] = tables[t][r_f]; the 39 diagonal rows the
 * reference gathers and never uses are not read).  tables: device array of N device pointers. */
int trs_ffm_model_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                          const float* w_feat, const float* const* tables, int64_t rows, int embed,
                          const float* bias, float* logits, int32_t* status, void* stream);

/* Interleaved-table variant of the FFM model forward: the B200-first memory layout of configs[4].  The N rows
 * T_0[r] .. T_{N-1}[r] of one row id are always read together (field_emb[b, t*N+f] = T_t[r_f] for every t,
 * multi_indices_field_aware_emb.py:100-105), so the shadow stores them contiguously, with the first-order weight:
 *     packed[r] = [ T_0[r] | T_1[r] | ... | T_{N-1}[r] | w_feat[r] | 0.. ]   pitch = trs_ffm_interleaved_pitch() floats
 * (a multiple of 32 floats = 128 bytes).  A sample then reads N contiguous chunks (one TMA bulk copy each) instead
 * of N(N-1) scattered 64-byte rows.  trs_ffm_pack_tables builds the shadow from the registered tables (w_feat may be
 * NULL: zeros; run again whenever they change; costs rows x pitch x 4 bytes);
 * trs_ffm_model_forward_interleaved computes exactly what trs_ffm_model_forward computes.
 * Restrictions: embed a power of two in [4, 128], fields <= 64, two samples' chunks within 227 KB of shared memory
 * (TRS_ERR_UNSUPPORTED otherwise). */
int64_t trs_ffm_interleaved_pitch(int fields, int embed);
int trs_ffm_pack_tables(const float* const* tables, const float* w_feat, int64_t rows, int fields, int embed,
                        float* packed, void* stream);
int trs_ffm_model_forward_interleaved(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                      const float* packed, int64_t rows, int embed,
                                      const float* bias, float* logits, int32_t* status, void* stream);

/* The same forward restricted to a LIST of pairs -- one rank's share when the tables are sharded over GPUs
 * (SURVEY.md 8e, configs[4]).  pair_list: DEVICE array of n_pairs entries (i << 16) | j, i < j (NULL = every pair);
 * the first-order term and the bias are added only for samples b in [first_begin, first_end), so that summing the
 * outputs of all ranks (reduce-scatter) counts them once.  `tables` may hold peer-mapped addresses of other GPUs. */
int trs_ffm_model_forward_pairs(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                const float* w_feat, const float* const* tables, int64_t rows, int embed,
                                const float* bias, const int* pair_list, int n_pairs,
                                int64_t first_begin, int64_t first_end,
                                float* logits, int32_t* status, void* stream);

/* ---- configs[4]: the field-aware tables SHARDED over the GPUs of one NVSwitch box, exchange at chunk granularity ------
 * (BASELINE.json north_star: "row-sharding the large embedding tables with an ... exchange of looked-up vectors over
 * NVLink only when a table exceeds one GPU's HBM"; the reference holds N full tables on one device,
 * torecsys/inputs/base/multi_indices_field_aware_emb.py:49-54, and reads T_t[r_f] for every (t, f), :90-111.)
 * Rank m owns the tables {t : t % world == m} and stores them interleaved per row id,
 *     shard_m[r] = [ T_m[r] | T_{m+world}[r] | ... ]   pitch = ceil(fields / world) * embed floats,
 * so what one rank holds of one row id is ONE contiguous chunk (320 B at 39 fields / 8 ranks / embed 16), fetched by
 * one bulk copy -- from local HBM or, through a peer mapping, over NVLink.  The pairs between the tables of ranks k and
 * m ("block (k, m)") are reduced on k for the samples of one parity and on m for the others: every dot product moves
 * exactly one of its two vectors, all bytes of every fetched chunk are used (csrc/ffm_blocks.cu).
 *   trs_ffm_shard_plan     HOST: the copy list and the dot-product item list of `rank`, for sample parity 0 and 1.
 *                          copy_tab int32 [2][copy_capacity][2] = {src rank | field << 8 | (bytes/16) << 16, dst byte
 *                          offset in the sample's stage}; item_tab uint32 [2][item_capacity] = 16-byte piece index of
 *                          one operand | the other << 16.  Pass NULL tables to query n_copies[2], n_items[2],
 *                          tx_bytes[2] (bytes fetched per sample) and *stage_bytes only.  No GPU needed.
 *   trs_ffm_shard_pack     shard[r][slot][:] = tables[slot][r][:] for the `slots` owned tables (DEVICE array of DEVICE
 *                          pointers, like trs_ffm_pack_tables), slots_pitch = ceil(fields / world); unused slots are zero-filled.
 *   trs_ffm_shard_resolve  rows_out[b][f] = idx[b][f] + offsets[f] as int32 (bounds-checked, status as everywhere) and
 *                          first_out[b] = bias[0] + sum_f w_feat[rows_out[b][f]] (either may be NULL): the part of
 *                          FieldAwareFactorizationMachineModel.forward (field_aware_factorization_machine.py:62-81)
 *                          that needs no exchange.
 *   trs_ffm_shard_blocks   partial[s] = the blocks `rank` reduces for sample s, for ALL batch_all samples (rows_all =
 *                          the all-gathered row ids), + first[s - own_lo] for own_lo <= s < own_hi.  shards = HOST array
 *                          of `world` device addresses as mapped in THIS process.  copy_tab / item_tab: DEVICE copies
 *                          of the plan tables.  Summing `partial` over the ranks (reduce-scatter) gives the logits. */
int trs_ffm_shard_plan(int fields, int world, int rank, int embed, int32_t* copy_tab, int copy_capacity,
                       uint32_t* item_tab, int item_capacity, int* n_copies, int* n_items, int* tx_bytes,
                       int* stage_bytes);
int trs_ffm_shard_pack(const float* const* tables, int slots, int slots_pitch, int64_t rows, int embed, float* shard,
                       void* stream);
int trs_ffm_shard_resolve(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                          int64_t rows, const float* w_feat, const float* bias, int32_t* rows_out, float* first_out,
                          int32_t* status, void* stream);
int trs_ffm_shard_blocks(const int32_t* rows_all, int64_t batch_all, int fields, int embed,
                         const float* const* shards, int world, int rank, const int32_t* copy_tab, int copy_capacity,
                         const uint32_t* item_tab, int item_capacity, const float* first, int64_t own_lo,
                         int64_t own_hi, float* partial, void* stream);

/* ---- 8f-3: fused indices -> logits forwards of three more models (Sequential.forward = Inputs.forward + model) --------
 * Arguments as in trs_deepfm_forward; the MLP (HOST arrays of DEVICE pointers) must end in ONE output.
 *   trs_nfm_forward        NeuralFactorizationMachineModel.forward (neural_factorization_machine.py:66-96):
 *                          logit = MLP(FM(emb)) + sum_n w_feat[r_n] (+ bias[0]);              MLP in = embed
 *   trs_fnn_forward        FactorizationMachineSupportedNeuralNetworkModel.forward
 *                          (factorization_machine_supported_neural_network.py:61-101):
 *                          logit = MLP(cat[w_feat[r_0..r_{N-1}], FM(emb)]);   `bias` ignored;   MLP in = fields + embed
 *   trs_pnn_inner_forward  ProductNeuralNetworkModel.forward with prod_method='inner' (product_neural_network.py:81-115):
 *                          logit = MLP(cat[IPN(emb), w_feat[r_0..r_{N-1}], bias[0]]);         MLP in = NC2 + fields + 1
 *                          (bias == NULL feeds 0 in that slot) */
int trs_nfm_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                    const float* w_feat, const float* w_emb, int64_t rows, int embed,
                    const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                    int activation, const float* bias, float* logits, int32_t* status, void* stream);
int trs_fnn_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                    const float* w_feat, const float* w_emb, int64_t rows, int embed,
                    const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                    int activation, const float* bias, float* logits, int32_t* status, void* stream);
int trs_pnn_inner_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                          const float* w_feat, const float* w_emb, int64_t rows, int embed,
                          const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                          int activation, const float* bias, float* logits, int32_t* status, void* stream);

/* ---- host-buffer entry points (the e2e path: pageable/pinned host indices in, host logits out) ----------------------
 * This is the step BEFORE the hot path in the reference: the DataLoader's collate function hands Sequential.forward a
 * host (B, N) int64 tensor (torecsys/data/dataloader/collate_fn.py:82, torecsys/models/sequential.py:31-44) and the
 * caller reads the (B, 1) prediction back.  A session owns trs_session_depth() independent slots (pinned staging,
 * device buffers, two streams each).  A submitted batch is split into `chunks` slices whose H2D(idx) / kernel /
 * D2H(logits) overlap; batches in different slots overlap each other, so a caller that keeps two or more batches in
 * flight keeps the host->device link busy all the time.
 *   trs_session_submit_*   enqueue one batch and return a ticket; no host synchronisation.  idx_host and logits_host
 *                          must stay valid until the ticket has been waited for.  TRS_ERR_INVALID_ARGUMENT when
 *                          every slot is in flight.
 *   trs_session_wait       block until the batch of `ticket` is complete: its logits are in logits_host and
 *                          *oob_count (when not NULL) = number of out-of-range lookups (status word0).
 *   trs_session_deepfm_forward_host[_packed] = submit + wait. */
typedef struct trs_session trs_session;
int trs_session_create(int64_t max_batch, int fields, int chunks, trs_session** out_session);
int trs_session_destroy(trs_session* session);
int trs_session_depth(void);
/* Host-side index narrowing (off by default).  With `threads` > 0 the session converts int64 host indices to int32 in
 * its pinned staging memory with that many host threads (the submitting thread included) before they cross the
 * link, halving the bytes of the link-bound path; the result is identical, values that do not fit int32 are
 * reported as out-of-range lookups exactly as in the int64 path.  threads = 0 turns it off, -1 picks
 * min(8, usable CPUs / 2).  Returns the thread count in use (>= 0) or a negative TRS_ERR_* code. */
int trs_session_set_index_narrowing(trs_session* session, int threads);
/* The conversion those threads run, on plain host arrays (no device involved; the CPU test suite checks every form
 * against the scalar one): dst[i] = src[i] if it fits int32, else INT32_MIN.  which: 0 = the form the sessions use on
 * this CPU, 1 = scalar, 2 = AVX2, 3 = AVX-512 (TRS_ERR_UNSUPPORTED when the CPU lacks the instruction set). */
int trs_host_narrow_indices(const int64_t* src, int32_t* dst, int64_t n, int which);
/* The narrowing pool of the sessions on plain host arrays: `threads` host threads (the caller included) convert
 * src[0, n) -> dst `reps` times; returns the best time of one pass in nanoseconds (how the conversion scales over the
 * host's cores: tools/r2_narrow_scaling.py), or a negative TRS_ERR_* code. */
int64_t trs_host_narrow_pool_ns(const int64_t* src, int32_t* dst, int64_t n, int threads, int reps);
/* Ordering against the caller's own work: the slots run on private streams, so device work the caller enqueued on
 * `stream` before a submit (packing the shadow table, uploading offsets, an optimizer step, load_state_dict) is not
 * ordered before the batch unless the session is told which stream that is.  With enabled != 0 every later submit
 * records an event on `stream` and makes its slot wait for it on the device (no host synchronisation). */
int trs_session_set_producer_stream(trs_session* session, void* stream, int enabled);
/* Generic form: the batch is fed to ANY fused indices -> logits entry point of this library (the five callers of
 * torecsys/models/sequential.py:31-44: trs_fm_model_forward, trs_deepfm_forward*, trs_dcn_forward,
 * trs_xdeepfm_forward, trs_ffm_model_forward*, and the 8f-3 models).  `fn` is called once per slice, synchronously
 * inside the submit call, and must enqueue exactly one forward over `batch` samples starting at `idx_dev` on `stream`:
 *     fn(ctx, lane, idx_dev, idx_bits, batch, logits_dev, status_dev, stream) -> TRS_OK or an error code.
 * lane < trs_session_lanes() identifies the (slot, stream) pair: slices with different lanes may run concurrently (give
 * each its own scratch), slices of one lane are stream-ordered.
 * `rows` (rows of the largest table, 0 = unknown) only gates the optional host-side index narrowing. */
typedef int (*trs_forward_fn)(void* ctx, int lane, const void* idx_dev, int idx_bits, int64_t batch, float* logits_dev,
                              int32_t* status_dev, void* stream);
int trs_session_submit_fn(trs_session* session, const void* idx_host, int idx_bits, int64_t batch, int fields,
                          trs_forward_fn fn, void* ctx, int64_t rows, float* logits_host, int64_t* ticket);
int trs_session_lanes(void);
/* The other four callers of torecsys/models/sequential.py:31-44 with host buffers; arguments as in the device entry
 * points (trs_fm_model_forward[_packed], trs_dcn_forward, trs_xdeepfm_forward, trs_ffm_model_forward[_interleaved]).
 * fm / ffm: `packed` != NULL selects the packed / interleaved shadow table, else the registered tables are used.
 * xdeepfm: `workspace` is cut into trs_session_lanes() equal parts, each of which must hold
 * trs_xdeepfm_workspace_bytes(slice batch = ceil(batch / chunks) rounded up to 16, ...). */
int trs_session_submit_fm(trs_session* session, const void* idx_host, int idx_bits, const int64_t* offsets,
                          int64_t batch, int fields, const float* w_feat, const float* w_emb, const float* packed,
                          int64_t rows, int embed, const float* bias, float* logits_host, int64_t* ticket);
int trs_session_submit_dcn(trs_session* session, const void* idx_host, int idx_bits, const int64_t* offsets,
                           int64_t batch, int fields, const float* w_emb, int64_t rows, int embed,
                           const float* cross_w, const float* cross_b, int cross_layers,
                           const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                           int activation, const float* fc_w, const float* fc_b, float* logits_host, int64_t* ticket);
int trs_session_submit_xdeepfm(trs_session* session, const void* idx_host, int idx_bits, const int64_t* offsets,
                               int64_t batch, int fields, const float* w_feat, const float* w_emb, int64_t rows,
                               int embed, const float* const* cin_w, const float* const* cin_scale,
                               const float* const* cin_shift, const int* cin_layer_sizes, int cin_layers,
                               int cin_is_direct, int cin_activation, const float* cin_fc_w, const float* cin_fc_b,
                               const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                               const float* const* mlp_b, int mlp_activation, const float* bias, void* workspace,
                               int64_t workspace_bytes, float* logits_host, int64_t* ticket);
int trs_session_submit_ffm(trs_session* session, const void* idx_host, int idx_bits, const int64_t* offsets,
                           int64_t batch, int fields, const float* w_feat, const float* const* tables,
                           const float* packed, int64_t rows, int embed, const float* bias, float* logits_host,
                           int64_t* ticket);
/* DeepFM on the packed table through the tcgen05 kernel (trs_deepfm_forward_tc; workspace from trs_deepfm_tc_prepare) */
int trs_session_submit_deepfm_tc(trs_session* session, const void* idx_host, int idx_bits,
                                 const int64_t* offsets, int64_t batch, int fields,
                                 const float* packed, int64_t rows,
                                 const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                 const float* const* mlp_b, int activation, const float* workspace, int variant,
                                 float* logits_host, int64_t* ticket);
int trs_session_submit_deepfm(trs_session* session, const void* idx_host, int idx_bits,
                              const int64_t* offsets, int64_t batch, int fields,
                              const float* w_feat, const float* w_emb, int64_t rows, int embed,
                              const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                              const float* const* mlp_b, int activation,
                              float* logits_host, int64_t* ticket);
int trs_session_submit_deepfm_packed(trs_session* session, const void* idx_host, int idx_bits,
                                     const int64_t* offsets, int64_t batch, int fields,
                                     const float* packed, int64_t rows,
                                     const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                     const float* const* mlp_b, int activation,
                                     float* logits_host, int64_t* ticket);
int trs_session_wait(trs_session* session, int64_t ticket, int64_t* oob_count);
int trs_session_deepfm_forward_host(trs_session* session, const void* idx_host, int idx_bits,
                                    const int64_t* offsets, int64_t batch, int fields,
                                    const float* w_feat, const float* w_emb, int64_t rows, int embed,
                                    const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                    const float* const* mlp_b, int activation,
                                    float* logits_host, int64_t* oob_count);
int trs_session_deepfm_forward_host_packed(trs_session* session, const void* idx_host, int idx_bits,
                                           const int64_t* offsets, int64_t batch, int fields,
                                           const float* packed, int64_t rows,
                                           const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                           const float* const* mlp_b, int activation,
                                           float* logits_host, int64_t* oob_count);

#ifdef __cplusplus
}
#endif
#endif /* TORECSYS_B200_H_ */
