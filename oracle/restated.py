"""CPU restatement of the reference's CTR forward hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference file:line (relative to /root/reference) whose eval-mode
arithmetic it restates.  Plain tensors in, plain tensors out (no named tensors), any float dtype
(fp32 = the reference's arithmetic; fp64 = error budgeting).  Dropout is identity (eval) and
BatchNorm uses running statistics (eval): SURVEY.md section 8a quirk 9.

Pinned by tests/test_oracle_golden.py against tests/golden/*.npz, which were produced by the
reference itself (oracle/make_golden.py).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- embeddings
def field_offsets(field_sizes: Sequence[int]) -> torch.Tensor:
    """torecsys/inputs/base/multi_indices_emb.py:54 and multi_indices_field_aware_emb.py:56.

    The reference builds the per-field row offsets through a float32 tensor
    (`torch.Tensor((0, *cumsum[:-1])).long()`), so cumulative sizes above 2**24 are rounded.
    Restated with the same rounding: exclusive prefix sum -> fp32 -> int64.
    """
    csum = np.cumsum(np.asarray(field_sizes, dtype=np.int64))
    excl = np.concatenate([[0], csum[:-1]]).astype(np.float32)
    return torch.from_numpy(excl.astype(np.int64))


def single_index_embedding(weight: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """torecsys/inputs/base/single_index_emb.py:56-59: (B,1) -> (B,1,E) = W[idx]."""
    return weight[idx.long()]


def multi_indices_embedding(weight: torch.Tensor, idx: torch.Tensor, offsets: torch.Tensor,
                            flatten: bool = False) -> torch.Tensor:
    """torecsys/inputs/base/multi_indices_emb.py:103-112: out[b,n,:] = W[idx[b,n] + off[n]]."""
    rows = idx.long() + offsets.view(1, -1)
    out = weight[rows]
    if flatten:
        out = out.reshape(out.shape[0], 1, -1)
    return out


def multi_indices_field_aware_embedding(weights: Sequence[torch.Tensor], idx: torch.Tensor,
                                        offsets: torch.Tensor) -> torch.Tensor:
    """torecsys/inputs/base/multi_indices_field_aware_emb.py:102-111.

    out[b, t*N + f, :] = W_t[idx[b,f] + off[f]]  (table-major concatenation on dim 1).
    """
    rows = idx.long() + offsets.view(1, -1)
    return torch.cat([w[rows] for w in weights], dim=1)


# ----------------------------------------------------------------------------- pair helpers
def pair_indices(num_fields: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lexicographic (i<j) pair list: inner_product_network.py:45-52, bilinear_interaction.py:205-211,
    attentional_factorization_machine.py:66-72, field_aware_factorization_machine.py:75-76."""
    rows, cols = [], []
    for i in range(num_fields - 1):
        for j in range(i + 1, num_fields):
            rows.append(i)
            cols.append(j)
    return torch.tensor(rows, dtype=torch.long), torch.tensor(cols, dtype=torch.long)


# ----------------------------------------------------------------------------- layers
def fm_layer(x: torch.Tensor) -> torch.Tensor:
    """torecsys/layers/ctr/factorization_machine.py:62-73: 0.5*((sum_n x)^2 - sum_n x^2), (B,N,E)->(B,E)."""
    s = x.sum(dim=1)
    q = (x ** 2).sum(dim=1)
    return 0.5 * (s ** 2 - q)


def ffm_layer(v: torch.Tensor, num_fields: int) -> torch.Tensor:
    """torecsys/layers/ctr/field_aware_factorization_machine.py:68-87.

    v (B, N*N, E) viewed (B,N,N,E); out[b,p,:] = v[b,i,j,:] * v[b,j,i,:] for i<j lexicographic.
    Written as the reference does it (one product per pair, then one concatenation) so the CPU
    baseline timing follows the reference's operation sequence.
    """
    b, _, e = v.shape
    v4 = v.reshape(b, num_fields, num_fields, e)
    outs = []
    for i in range(num_fields - 1):
        for j in range(i + 1, num_fields):
            outs.append((v4[:, i, j] * v4[:, j, i]).unsqueeze(1))
    return torch.cat(outs, dim=1)


def cross_layer(x: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                cut_gradient_through_h0: bool = False) -> torch.Tensor:
    """torecsys/layers/ctr/cross_network.py:65-79: h <- x * Linear_l(h) + x (Linear(E,E) on the last dim,
    residual is x0, not h).  (B,N,E)->(B,N,E).  The reference starts the chain from emb_inputs.detach() (:65): same
    forward values, but no gradient reaches x through h_0 -- cut_gradient_through_h0 restates that for gradient tests."""
    h = x.detach() if cut_gradient_through_h0 else x
    for w, b in zip(weights, biases):
        h = F.linear(h, w, b)
        h = x * h
        h = h + x
    return h


def _activation(name: Optional[str]):
    if name is None or name == 'none':
        return lambda t: t
    return {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[name]


def cin_layer(x: torch.Tensor,
              conv_w: Sequence[torch.Tensor], conv_b: Sequence[Optional[torch.Tensor]],
              bn: Sequence[Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, float]]],
              fc_w: torch.Tensor, fc_b: torch.Tensor,
              is_direct: bool = False, activation: Optional[str] = 'relu') -> torch.Tensor:
    """torecsys/layers/ctr/compress_interaction_network.py:105-182 (eval mode).

    conv_w[l]: (C_l, N*H_{l-1}) (the Conv1d weight with its trailing kernel dim dropped),
    bn[l] = (weight, bias, running_mean, running_var, eps) or None.
    z[b, xf*H + y, e] = x[b,xf,e] * h[b,y,e] (x-major channel order, :125-132);
    o = act(BN(conv(z))) (:137); not direct: first half of the channels is the layer's output and the
    second half the next hidden state, for EVERY layer (:151-156, the `i != len-1` guard never fails);
    out = fc(sum_e cat(directs)) (:176-179).
    """
    act = _activation(activation)
    b, n, e = x.shape
    h = x
    directs = []
    for l in range(len(conv_w)):
        hh = h.shape[1]
        z = (x.unsqueeze(2) * h.unsqueeze(1)).reshape(b, n * hh, e)
        o = torch.einsum('oc,bce->boe', conv_w[l], z)
        if conv_b[l] is not None:
            o = o + conv_b[l].view(1, -1, 1)
        if bn[l] is not None:
            g, beta, mean, var, eps = bn[l]
            o = (o - mean.view(1, -1, 1)) / torch.sqrt(var.view(1, -1, 1) + eps) * g.view(1, -1, 1) \
                + beta.view(1, -1, 1)
        o = act(o)
        if is_direct:
            d, h = o, o
        else:
            half = o.shape[1] // 2
            d, h = o[:, :half], o[:, half:]
        directs.append(d)
    pooled = torch.cat(directs, dim=1).sum(dim=-1)
    return F.linear(pooled, fc_w, fc_b)


def ipn_layer(x: torch.Tensor) -> torch.Tensor:
    """torecsys/layers/ctr/inner_product_network.py:67-77: out[b,p] = <x_i, x_j>, (B,N,E)->(B,P)."""
    r, c = pair_indices(x.shape[1])
    return (x[:, r] * x[:, c]).sum(dim=-1)


def bilinear_layer(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                   bilinear_type: str = 'all') -> torch.Tensor:
    """torecsys/layers/ctr/bilinear_interaction.py:244-255 with FieldAllTypeBilinear.forward (:72-76,
    weight (E,E), bias (E)) or FieldEachTypeBilinear.forward (:144-149, weight (P,E,E), bias (P,E)).
    out[b,p,:] = (x_i @ W_(p)) * x_j + b_(p)."""
    r, c = pair_indices(x.shape[1])
    p, q = x[:, r], x[:, c]
    if bilinear_type == 'all':
        out = torch.matmul(p, weight) * q
    elif bilinear_type == 'each':
        out = torch.matmul(p.unsqueeze(-2), weight).squeeze(-2) * q
    else:
        raise ValueError(bilinear_type)
    if bias is not None:
        out = out + bias
    return out


def afm_layer(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """torecsys/layers/ctr/attentional_factorization_machine.py:99-120 (eval: both dropouts identity).

    products = x_i * x_j (B,P,E); scores = softmax_p(OutProj(relu(Linear(products)))) (B,P,1);
    out = sum_p scores * products (B,E).  Returns (out, scores)."""
    r, c = pair_indices(x.shape[1])
    prod = x[:, r] * x[:, c]
    s = F.linear(torch.relu(F.linear(prod, w1, b1)), w2, b2)
    s = torch.softmax(s, dim=1)
    return (prod * s).sum(dim=1), s


def mlp_layer(x: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
              activation: Optional[str] = 'relu') -> torch.Tensor:
    """torecsys/layers/ctr/multilayer_perceptron.py:53-84 (eval: dropout identity).  weights[-1] is
    `LinearOutput` (no activation after it); every other Linear is followed by the activation."""
    act = _activation(activation)
    h = x
    for w, b in zip(weights[:-1], biases[:-1]):
        h = act(F.linear(h, w, b))
    return F.linear(h, weights[-1], biases[-1])


def opn_layer(x: torch.Tensor, kernel: torch.Tensor, kernel_type: str = 'mat') -> torch.Tensor:
    """torecsys/layers/ctr/outer_product_network.py:80-131.  p = x[:, i_p], q = x[:, j_p] for pairs i<j;
    'mat' (kernel (E,P,E)): kp[b,h,p] = sum_e p[b,p,e] * kernel[h,p,e] (:105-110), out[b,p] = sum_h kp[b,h,p] * q[b,p,h]
    (:116); 'vec' (1,P,E) / 'num' (1,P,1): out[b,p] = sum_e p * q * kernel (:124)."""
    i, j = pair_indices(x.shape[1])
    p, q = x[:, i], x[:, j]
    if kernel_type == 'mat':
        kp = (p.unsqueeze(1) * kernel).sum(dim=-1)          # (B, E=h, P)
        return (kp.permute(0, 2, 1) * q).sum(dim=-1)
    if kernel_type in ('vec', 'num'):
        return (p * q * kernel).sum(dim=-1)
    raise ValueError('kernel_type only allows: ["mat", "num", "vec"].')


def senet_layer(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                activation: Optional[str] = 'relu') -> torch.Tensor:
    """torecsys/layers/ctr/compose_excitation_network.py:72-109: AdaptiveAvgPool1d(1) over the embedding (:82),
    fc = ReductionLinear -> act -> AdditionLinear -> act (:66-70, the SAME activation instance twice),
    out[b,m,:] = x[b,m,:] * a[b,m] (the einsum 'ijk,ijh->ijk' with h of size 1, :104).  x is (B, M, E) with
    M = num_fields or num_fields^2 (`squared`)."""
    act = _activation(activation)
    pooled = F.adaptive_avg_pool1d(x, 1).squeeze(-1)
    a = act(F.linear(act(F.linear(pooled, w1, b1)), w2, b2))
    return x * a.unsqueeze(-1)


# ----------------------------------------------------------------------------- model glue (a12)
def fm_model(feat: torch.Tensor, emb: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """torecsys/models/ctr/factorization_machine.py:55-71: sum_n feat + sum_e FM(emb) (+ bias) -> (B,1)."""
    out = fm_layer(emb).sum(dim=1, keepdim=True) + feat.sum(dim=1)
    if bias is not None:
        out = out + bias.reshape(1, 1)
    return out


def deepfm_model(feat: torch.Tensor, emb: torch.Tensor, mlp_w, mlp_b, activation='relu') -> torch.Tensor:
    """torecsys/models/ctr/deep_fm.py:67-108: MLP(flat emb) + sum(cat[FM(emb), feat]) -> (B,1), no bias."""
    b = emb.shape[0]
    fm_out = torch.cat([fm_layer(emb), feat.reshape(b, -1)], dim=1).sum(dim=1, keepdim=True)
    deep_out = mlp_layer(emb.reshape(b, -1), mlp_w, mlp_b, activation)
    return deep_out + fm_out


def dcn_model(emb: torch.Tensor, cross_w, cross_b, mlp_w, mlp_b, fc_w, fc_b, activation='relu') -> torch.Tensor:
    """torecsys/models/ctr/deep_and_cross_network.py:76-98: fc(flatten(cat[Cross(x), MLP_per_field(x)], -1)).
    The cross branch starts from emb.detach() as the reference layer does (cross_network.py:65): identical forward
    values, and differentiating this formula gives the reference's embedding gradient (quirk 3 of SURVEY.md 8a)."""
    b = emb.shape[0]
    cat = torch.cat([cross_layer(emb, cross_w, cross_b, cut_gradient_through_h0=True),
                     mlp_layer(emb, mlp_w, mlp_b, activation)], dim=-1)
    return F.linear(cat.reshape(b, -1), fc_w, fc_b)


def xdeepfm_model(feat, emb, cin_args: dict, mlp_w, mlp_b, bias, activation='relu') -> torch.Tensor:
    """torecsys/models/ctr/xdeep_fm.py:95-124: sum_n feat + CIN(emb) + MLP(flat emb) + bias -> (B,1)."""
    b = emb.shape[0]
    return feat.sum(dim=1) + cin_layer(emb, **cin_args) + mlp_layer(emb.reshape(b, -1), mlp_w, mlp_b, activation) \
        + bias.reshape(1, 1)


def ffm_model(feat: torch.Tensor, field_emb: torch.Tensor, num_fields: int, bias: torch.Tensor) -> torch.Tensor:
    """torecsys/models/ctr/field_aware_factorization_machine.py:55-81: sum_{p,e} FFM + sum_n feat + bias."""
    second = ffm_layer(field_emb, num_fields).sum(dim=(1, 2)).unsqueeze(1)
    return second + feat.sum(dim=1) + bias.reshape(1, 1)


# ----------------------------------------------------------------------------- model glue (8f-3: the other consumers)
def pnn_model(feat, emb, second: torch.Tensor, mlp_w, mlp_b, bias: Optional[torch.Tensor], activation='relu'):
    """torecsys/models/ctr/product_neural_network.py:81-115: MLP(cat[pnn(emb) (B,P), feat (B,N), bias (B,1)]);
    `second` = the inner/outer product layer's output."""
    b = emb.shape[0]
    parts = [second, feat.reshape(b, -1)]
    if bias is not None:
        parts.append(bias.reshape(1, 1).repeat(b, 1))
    return mlp_layer(torch.cat(parts, dim=1), mlp_w, mlp_b, activation)


def fibinet_model(emb, senet_args, bil_emb, bil_senet, bilinear_type, mlp_w, mlp_b, activation='relu'):
    """torecsys/models/ctr/feature_importance_and_bilinear_feature_interaction_network.py:74-109:
    MLP(flatten(cat[Bilinear_a(emb), Bilinear_b(SENET(emb))], dim=N))."""
    b = emb.shape[0]
    inter = bilinear_layer(emb, bil_emb[0], bil_emb[1], bilinear_type)
    s_inter = bilinear_layer(senet_layer(emb, *senet_args), bil_senet[0], bil_senet[1], bilinear_type)
    return mlp_layer(torch.cat([inter, s_inter], dim=1).reshape(b, -1), mlp_w, mlp_b, activation)


def afm_model(feat, emb, afm_args, bias: Optional[torch.Tensor]):
    """torecsys/models/ctr/attentional_factorization_machine.py:53-84: sum_e AFM(emb) + sum_n feat (+ bias)."""
    out = afm_layer(emb, *afm_args)[0].sum(dim=1, keepdim=True) + feat.sum(dim=1)
    return out if bias is None else out + bias.reshape(1, 1)


def nfm_model(feat, emb, mlp_w, mlp_b, bias: Optional[torch.Tensor], activation='relu'):
    """torecsys/models/ctr/neural_factorization_machine.py:66-96: MLP(FM(emb)) + sum_n feat (+ bias)."""
    out = mlp_layer(fm_layer(emb), mlp_w, mlp_b, activation) + feat.sum(dim=1)
    return out if bias is None else out + bias.reshape(1, 1)


def fnn_model(feat, emb, mlp_w, mlp_b, activation='relu'):
    """torecsys/models/ctr/factorization_machine_supported_neural_network.py:61-101: MLP(cat[feat (B,N), FM(emb)])."""
    b = emb.shape[0]
    return mlp_layer(torch.cat([feat.reshape(b, -1), fm_layer(emb)], dim=1), mlp_w, mlp_b, activation)


def deep_ffm_model(field_emb, num_fields, mlp_w, mlp_b, activation='relu'):
    """torecsys/models/ctr/deep_ffm.py:63-104: sum_O MLP(flatten FFM(v)) + sum_{n,e} v."""
    b = field_emb.shape[0]
    second = mlp_layer(ffm_layer(field_emb, num_fields).reshape(b, -1), mlp_w, mlp_b, activation)
    return second.sum(dim=1, keepdim=True) + field_emb.sum(dim=(1, 2)).unsqueeze(1)


def fat_deep_ffm_model(field_emb, num_fields, cen_args, mlp_w, mlp_b, activation='relu'):
    """torecsys/models/ctr/fat_deep_ffm.py:69-112: aem = CEN(v); sum_{n,e} aem + MLP(flatten FFM(aem))."""
    b = field_emb.shape[0]
    aem = senet_layer(field_emb, *cen_args)
    return aem.sum(dim=(1, 2)).unsqueeze(1) + mlp_layer(ffm_layer(aem, num_fields).reshape(b, -1), mlp_w, mlp_b,
                                                        activation)


# ----------------------------------------------------------------------------- end-to-end (indices -> logits)
def deepfm_from_indices(idx, offsets, w_feat, w_emb, mlp_w, mlp_b, activation='relu'):
    """Inputs.forward + DeepFM.forward: torecsys/inputs/inputs.py:69-87 then deep_fm.py:55-110."""
    return deepfm_model(multi_indices_embedding(w_feat, idx, offsets),
                        multi_indices_embedding(w_emb, idx, offsets), mlp_w, mlp_b, activation)


def fm_from_indices(idx, offsets, w_feat, w_emb, bias):
    return fm_model(multi_indices_embedding(w_feat, idx, offsets), multi_indices_embedding(w_emb, idx, offsets), bias)


def dcn_from_indices(idx, offsets, w_emb, cross_w, cross_b, mlp_w, mlp_b, fc_w, fc_b, activation='relu'):
    return dcn_model(multi_indices_embedding(w_emb, idx, offsets), cross_w, cross_b, mlp_w, mlp_b, fc_w, fc_b,
                     activation)


def xdeepfm_from_indices(idx, offsets, w_feat, w_emb, cin_args, mlp_w, mlp_b, bias, activation='relu'):
    return xdeepfm_model(multi_indices_embedding(w_feat, idx, offsets), multi_indices_embedding(w_emb, idx, offsets),
                         cin_args, mlp_w, mlp_b, bias, activation)


def ffm_from_indices(idx, offsets, w_feat, tables, bias):
    n = idx.shape[1]
    return ffm_model(multi_indices_embedding(w_feat, idx, offsets),
                     multi_indices_field_aware_embedding(tables, idx, offsets), n, bias)
