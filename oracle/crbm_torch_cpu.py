"""Multithreaded CPU restatement of the reference's Theano graph — TEST/BENCH INFRASTRUCTURE, not product code.

Purpose: the CPU arm of bench.py (`cpu_baseline`, `--impl reference`). The reference executes its graph through
Theano (CorrMM: im2col + BLAS sgemm on CPU); Theano is not installable offline, so the same graph is expressed here
with torch-CPU ops that lower the same way (F.conv2d / conv_transpose2d -> im2col + MKL sgemm), fp32, all host
threads. Each function cites the reference lines it follows (secomo/convRBM.py). It is checked against the NumPy
oracle in tests/test_oracle_torch_cpu.py. Only bench.py and tests import it.
"""
import numpy as np
import torch
import torch.nn.functional as F


class TorchCpuCRBM(object):
    def __init__(self, num_motifs, motif_length, doublestranded=True, batchsize=20, learning_rate=0.1, momentum=0.95,
                 pooling=1, cd_k=5, rho=0.01, lambda_rate=0.1, seed=0, W=None, b=None, c=None):
        g = torch.Generator().manual_seed(seed)
        K, M = num_motifs, motif_length
        self.K, self.M, self.P, self.ds, self.cd_k, self.bs = K, M, pooling, doublestranded, cd_k, batchsize
        self.lr, self.mu, self.rho, self.lam = learning_rate, momentum, rho, lambda_rate
        self.W = torch.randn(K, 1, 4, M, generator=g) if W is None else torch.as_tensor(W, dtype=torch.float32).clone()
        if b is None:
            from scipy.stats import norm
            b = np.zeros((1, K), dtype=np.float32) + norm.ppf(rho, 0, np.sqrt(M))
        self.b = torch.as_tensor(b, dtype=torch.float32).clone()
        self.c = torch.zeros(1, 4) if c is None else torch.as_tensor(c, dtype=torch.float32).clone()
        self.vW, self.vb, self.vc = torch.zeros_like(self.W), torch.zeros_like(self.b), torch.zeros_like(self.c)
        self.fh = torch.zeros(batchsize, K, 1, 200)                      # :168
        self.fhp = torch.zeros(batchsize, K, 1, 200) if doublestranded else None
        self.gen = g

    # :238-243
    def act(self, D, flip=False):
        W = torch.flip(self.W, dims=(2, 3)) if flip else self.W
        return F.conv2d(D, W) + self.b.reshape(1, -1, 1, 1)

    # :245-257
    def prob(self, x):
        N, K, _, Lh = x.shape
        g = x.reshape(N, K, 1, Lh // self.P, self.P)
        e = torch.exp(g)
        return (e / torch.sum(1.0 + e, dim=4, keepdim=True)).reshape(N, K, 1, Lh)

    @staticmethod
    def _multinomial(rows, u):
        csum = torch.cumsum(rows, dim=1)
        hit = csum > u[:, None]
        return (hit & (torch.cumsum(hit.to(torch.int32), dim=1) == 1)).to(rows.dtype)

    # :259-267
    def sample_h(self, p):
        N, K, _, Lh = p.shape
        rows = p.reshape(-1, self.P)
        u = torch.rand(rows.shape[0], generator=self.gen)
        return self._multinomial(rows, u).reshape(N, K, 1, Lh)

    # :277-292  (full-border true convolution == conv_transpose2d with the same filters)
    def v_act(self, h, hp):
        out = F.conv_transpose2d(h, self.W)
        if hp is not None:
            out = out + F.conv_transpose2d(hp, torch.flip(self.W, dims=(2, 3)))
        return out + self.c.reshape(1, 1, 4, 1)

    # :294-315
    def v_sample(self, act):
        p = torch.softmax(act, dim=2)
        N, _, A, Lv = p.shape
        rows = p.permute(0, 1, 3, 2).reshape(N * Lv, A)
        u = torch.rand(rows.shape[0], generator=self.gen)
        s = self._multinomial(rows, u)
        return s.reshape(N, 1, Lv, A).permute(0, 1, 3, 2).contiguous()

    # :327-371
    def stats(self, ph, php, V):
        def vh(p):
            out = F.conv2d(V.permute(1, 0, 2, 3), p.permute(1, 0, 2, 3))          # (1,K,4,M)
            return out.permute(1, 0, 2, 3) / float(p.shape[0] * p.shape[2] * p.shape[3])
        gW, gb = vh(ph), ph.mean(dim=(0, 2, 3))[None]
        if php is not None:
            gW = (gW + torch.flip(vh(php), dims=(2, 3))) / 2.0
            gb = (gb + php.mean(dim=(0, 2, 3))[None]) / 2.0
        a = V.mean(dim=(0, 1, 3))[None]
        return gW, gb, a + torch.flip(a, dims=(1,))

    # :440-451 (autograd, as the reference uses T.grad)
    def sparsity_grad(self, D):
        W = self.W.clone().requires_grad_(True)
        b = self.b.clone().requires_grad_(True)
        x = F.conv2d(D, W) + b.reshape(1, -1, 1, 1)
        N, K, _, Lh = x.shape
        g = x.reshape(N, K, 1, Lh // self.P, self.P)
        e = torch.exp(g)
        p = (e / torch.sum(1.0 + e, dim=4, keepdim=True)).reshape(N, K, 1, Lh).mean(dim=(0, 2, 3))
        r = -torch.mean(self.rho * torch.log(p) + (1 - self.rho) * torch.log(1 - p))
        gW, gb = torch.autograd.grad(r, (W, b))
        return gW, gb

    # :373-438
    def train_step(self, D):
        with torch.no_grad():
            ph = self.prob(self.act(D))
            php = self.prob(self.act(D, True)) if self.ds else None
            gWd, gbd, gcd = self.stats(ph, php, D)
            h, hp = self.fh, self.fhp
            for _ in range(self.cd_k):
                v = self.v_sample(self.v_act(h, hp))
                pm = self.prob(self.act(v))
                h = self.sample_h(pm)
                if self.ds:
                    pmp = self.prob(self.act(v, True))
                    hp = self.sample_h(pmp)
                else:
                    pmp, hp = None, None
            gWm, gbm, gcm = self.stats(pm, pmp, v)
        rW, rb = self.sparsity_grad(D)
        with torch.no_grad():
            self.vW = self.mu * self.vW + self.lr * (gWd - gWm - self.lam * rW)
            self.vb = self.mu * self.vb + self.lr * (gbd - gbm - self.lam * rb)
            self.vc = self.mu * self.vc + self.lr * (gcd - gcm)
            self.W, self.b, self.c = self.W + self.vW, self.b + self.vb, self.c + self.vc
            self.fh, self.fhp = h, hp

    # :507-514
    def hit_probs(self, D):
        with torch.no_grad():
            if self.ds:
                return self.prob(self.act(D))
            return self.prob(self.act(D) + self.act(D, True))


def random_onehot(N, L, seed):
    g = torch.Generator().manual_seed(seed)
    codes = torch.randint(0, 4, (N, L), generator=g)
    return F.one_hot(codes, 4).permute(0, 2, 1).reshape(N, 1, 4, L).to(torch.float32).contiguous()
