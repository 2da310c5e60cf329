import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torecsys_b200 import ops
torch.set_printoptions(precision=4, linewidth=200, sci_mode=False)
for (b, n, e, each) in [(1, 3, 8, False), (1, 4, 16, False), (17, 3, 8, True)]:
    gen = torch.Generator().manual_seed(14)
    x = torch.randn(b, n, e, generator=gen, dtype=torch.float64, requires_grad=True)
    i, j = torch.triu_indices(n, n, offset=1)
    pairs = i.numel()
    w = (torch.randn(*((pairs, e, e) if each else (e, e)), generator=gen, dtype=torch.float64) / e ** 0.5).requires_grad_()
    g = torch.randn(b, pairs, e, generator=gen, dtype=torch.float64)
    y = torch.matmul(x[:, i].unsqueeze(-2), w).squeeze(-2) if each else torch.matmul(x[:, i], w)
    ((y * x[:, j]) * g).sum().backward()
    gx, gw, gb = ops.bilinear_backward(x.detach().float().cuda(), w.detach().float().cuda(), g.float().cuda(), each)
    d = (gx.cpu().double() - x.grad)
    print('case', b, n, e, each, 'max err per (sample, field):')
    print(d.abs().amax(-1))
    print('dW err', float((gw.cpu().double() - w.grad).abs().max()), 'db err', float((gb.cpu().double() - g.sum(0).reshape(gb.shape) if each else gb.cpu().double() - g.sum((0, 1))).abs().max()))
    if b == 1:
        print('got\n', gx.cpu()[0]); print('want\n', x.grad[0].float())
        # decompose: i-role and j-role parts of the reference
        xi, xj = x.detach()[:, i], x.detach()[:, j]
        W = w.detach()
        t = g * xj
        irole = torch.zeros_like(x.detach()); jrole = torch.zeros_like(x.detach())
        yy = torch.matmul(xi, W)
        for p in range(pairs):
            irole[:, i[p]] += t[:, p] @ W.T
            jrole[:, j[p]] += g[:, p] * yy[:, p]
        print('i-role part\n', irole[0].float()); print('j-role part\n', jrole[0].float())
