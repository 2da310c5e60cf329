"""torch.ops.torecsys_b200.* (torecsys_b200/dispatch.py): the hot-path ops are registered with the
dispatcher -- fake implementations propagate shapes without a GPU, there is no CPU kernel, and on the GPU the
registered ops give bit-for-bit what the ctypes wrappers give, differentiate through the library's backward kernels
and survive a full-graph torch.compile capture (backend aot_eager: graph capture + functionalisation, no codegen)."""
import pytest
import torch

OPS = ['embedding_gather', 'fm', 'ffm', 'ipn', 'cross', 'bilinear', 'afm']


def test_ops_are_registered_with_schemas():
    import torecsys_b200  # noqa: F401
    for name in OPS:
        op = getattr(torch.ops.torecsys_b200, name)
        assert name in str(op.default._schema)
    assert 'Tensor? offsets' in str(torch.ops.torecsys_b200.embedding_gather.default._schema)
    assert '(Tensor, Tensor)' in str(torch.ops.torecsys_b200.afm.default._schema)


def test_fake_implementations_propagate_shapes_without_a_gpu():
    import torecsys_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    t = torch.ops.torecsys_b200
    with FakeTensorMode():
        x = torch.empty(7, 39, 16, device='cuda')
        assert t.fm(x).shape == (7, 16) and t.ipn(x).shape == (7, 741)
        assert t.ffm(torch.empty(7, 39 * 39, 16, device='cuda'), 39).shape == (7, 741, 16)
        assert t.cross(x, torch.empty(3, 16, 16, device='cuda'), torch.empty(3, 16, device='cuda')).shape == x.shape
        assert t.bilinear(x, torch.empty(16, 16, device='cuda'), None, False).shape == (7, 741, 16)
        out, scores = t.afm(x, torch.empty(8, 16, device='cuda'), torch.empty(8, device='cuda'),
                            torch.empty(1, 8, device='cuda'), torch.empty(1, device='cuda'))
        assert out.shape == (7, 16) and scores.shape == (7, 741, 1)
        w = torch.empty(1000, 16, device='cuda')
        idx = torch.empty(7, 39, dtype=torch.int64, device='cuda')
        assert t.embedding_gather(w, idx, torch.empty(39, dtype=torch.int64, device='cuda')).shape == (7, 39, 16)
        assert t.embedding_gather(w, idx, None).device.type == 'cuda'


def test_no_cpu_kernel():
    import torecsys_b200  # noqa: F401
    with pytest.raises(NotImplementedError):
        torch.ops.torecsys_b200.fm(torch.zeros(2, 3, 4))


@pytest.mark.gpu
def test_registered_ops_match_the_wrappers_and_differentiate():
    from torecsys_b200 import autograd as A, ops
    t = torch.ops.torecsys_b200
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(33, 13, 16, device='cuda', generator=g)
    assert torch.equal(t.fm(x), ops.fm(x)) and torch.equal(t.ipn(x), ops.ipn(x))
    w = torch.randn(500, 16, device='cuda', generator=g)
    idx = torch.randint(0, 30, (33, 13), device='cuda')
    off = (torch.arange(13, device='cuda') * 30)
    assert torch.equal(t.embedding_gather(w, idx, off), ops.embedding_gather(w, idx, off))
    cw, cb = torch.randn(3, 16, 16, device='cuda', generator=g) * 0.2, torch.randn(3, 16, device='cuda', generator=g) * 0.1
    assert torch.equal(t.cross(x, cw, cb), ops.cross(x, cw, cb))
    # gradients: the registered formulas call the same backward kernels as the autograd.Function path of the modules
    for fn_new, fn_old, args in (
            (t.fm, A.FmFn.apply, (x,)),
            (t.ipn, A.IpnFn.apply, (x,)),
            (t.cross, A.CrossFn.apply, (x, cw, cb)),
            (lambda a, b: t.bilinear(a, b, None, False), lambda a, b: A.BilinearFn.apply(a, b, None, False),
             (x, torch.randn(16, 16, device='cuda', generator=g) * 0.2))):
        a1 = [a.clone().requires_grad_(True) for a in args]
        a2 = [a.clone().requires_grad_(True) for a in args]
        o1, o2 = fn_new(*a1), fn_old(*a2)
        assert torch.equal(o1, o2)
        go = torch.randn_like(o1)
        g1 = torch.autograd.grad(o1, a1, go)
        g2 = torch.autograd.grad(o2, a2, go)
        for u, v in zip(g1, g2):   # (weight gradients are summed with atomics: equal up to the order of the additions)
            assert torch.allclose(u, v, rtol=1e-5, atol=1e-5 * float(v.abs().max()))
    wg = w.clone().requires_grad_(True)
    out = t.embedding_gather(wg, idx, off)
    (gw,) = torch.autograd.grad(out, [wg], torch.ones_like(out))
    want = torch.zeros_like(w).index_add_(0, (idx + off).reshape(-1), torch.ones(33 * 13, 16, device='cuda'))
    assert torch.equal(gw, want)


@pytest.mark.gpu
def test_full_graph_capture_keeps_one_node_per_op():
    import torecsys_b200  # noqa: F401
    t = torch.ops.torecsys_b200
    w = torch.randn(500, 16, device='cuda')
    idx = torch.randint(0, 30, (64, 13), device='cuda')
    off = torch.arange(13, device='cuda') * 30

    def f(w, idx, off):
        x = t.embedding_gather(w, idx, off)
        return t.fm(x).sum(dim=1, keepdim=True) + t.ipn(x).sum(dim=1, keepdim=True)

    want = f(w, idx, off)
    compiled = torch.compile(f, backend='aot_eager', fullgraph=True)
    assert torch.allclose(compiled(w, idx, off), want, rtol=0, atol=0)

    class M(torch.nn.Module):
        def forward(self, w, idx, off):
            return f(w, idx, off)
    ep = torch.export.export(M(), (w, idx, off))
    targets = [str(n.target) for n in ep.graph.nodes if n.op == 'call_function']
    assert any('torecsys_b200.embedding_gather' in s for s in targets) and any('torecsys_b200.fm' in s for s in targets)
