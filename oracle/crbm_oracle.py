"""NumPy restatement of the reference cRBM hot path — TEST INFRASTRUCTURE, not product code.

Oracle for `crbm_b200`: a line-by-line CPU restatement of `secomo/convRBM.py` (schulter/crbm 1.1.0) in plain
NumPy, with every stochastic draw taken from caller-supplied uniforms so that results are reproducible and
can be replayed into the CUDA kernels. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module; nothing under `crbm_b200/` does.

PINNING STATUS. Theano (the library that executes the reference's arithmetic; un-pinned in the reference's
setup.py:20, last release 1.0.5) is not installable here, so the reference cannot run natively. Instead:
  * `oracle/theano_shim` implements the Theano calls convRBM.py makes; the reference's own test-suite
    (tests/testcrbm.py — NumPy control loops for conv2d valid/full/flip, sigmoid, 4-way softmax, one-hot
    sampling) passes on it unmodified (`oracle/run_reference_tests.sh`, 17 passed).
  * `tests/golden/make_golden.py` imports the UNMODIFIED /root/reference/secomo/convRBM.py on that shim and
    records inputs, injected uniforms and outputs of every compiled function (train step incl. T.grad sparsity
    term, evaluateData/Params, freeEnergy, fePerMotif, getHitProbs) as fixtures under tests/golden/.
  * `tests/test_oracle_golden.py` checks this oracle against those fixtures, and `tests/test_oracle_kat.py`
    against the reference tests' control loops restated in-test.
What remains "parity unpinned": anything that depends on Theano internals the shim restates from
documentation rather than executes — the MRG31k3p random stream (only the distribution is reproduced) and
fp32 rounding/summation order of Theano's CorrMM lowering.

Layouts follow the reference: data (N,1,4,L) one-hot, W (K,1,4,M), b (1,K), c (1,4), hidden (N,K,1,Lh).
"""
import numpy as np

LOG2_4 = 2.0


def _f(x, dtype):
    return np.asarray(x, dtype=dtype)


# ------------------------------------------------------------------------------------------------------
# A1  convRBM.py:238-243  _bottomUpActivity
# ------------------------------------------------------------------------------------------------------
def bottom_up_activity(D, W, b, flip=False):
    """x[n,k,0,j] = b[0,k] + sum_{a,m} D[n,0,a,j+m] * Wf[k,0,a,m];  Wf = W (flip False, cross-correlation,
    conv(..., filter_flip=False)) or W[:,:,::-1,::-1] (flip True: the reverse-complement motif)."""
    dtype = np.result_type(D.dtype, W.dtype)
    Wf = W[:, :, ::-1, ::-1] if flip else W
    M = W.shape[3]
    win = np.lib.stride_tricks.sliding_window_view(D[:, 0], M, axis=2)       # (N,4,Lh,M)
    x = np.einsum('najm,kam->nkj', win.astype(dtype), Wf[:, 0].astype(dtype), optimize=True)
    x = x + b[0][None, :, None]
    return x[:, :, None, :]


# A2  convRBM.py:245-257  _bottomUpProbability
def bottom_up_probability(x, pooling):
    """Groups of `pooling` consecutive positions: p = exp(x) / sum_group(1 + exp(x))  (normaliser is
    P + sum exp, convRBM.py:252). No max-subtraction (as in the reference)."""
    N, K, _, Lh = x.shape
    if Lh % pooling:
        raise ValueError("hidden length %d not divisible by pooling %d (reference reshape fails)" % (Lh, pooling))
    g = x.reshape(N, K, 1, Lh // pooling, pooling)
    norm = np.sum(1.0 + np.exp(g), axis=4, keepdims=True)
    return (np.exp(g) / norm).reshape(N, K, 1, Lh)


def multinomial_from_uniform(pvals, u):
    """Theano MultinomialFromUniform, n=1: first outcome whose running sum exceeds u; all-zero row if none."""
    csum = np.cumsum(pvals, axis=1)
    hit = csum > np.asarray(u, dtype=pvals.dtype)[:, None]
    first = hit & (np.cumsum(hit, axis=1) == 1)
    return first.astype(pvals.dtype)


# A3  convRBM.py:259-267  _bottomUpSample
def bottom_up_sample(p, pooling, u):
    """One uniform per pooling group, rows ordered (n,k,group) (reshape at convRBM.py:263-264)."""
    N, K, _, Lh = p.shape
    rows = p.reshape(N * K * (Lh // pooling), pooling)
    return multinomial_from_uniform(rows, u).reshape(N, K, 1, Lh)


# A5  convRBM.py:277-292  _topDownActivity
def top_down_activity(h, hprime, W, c):
    """v[n,0,a,i] = c[0,a] + sum_{k,m} h[n,k,0,i-m] W[k,0,a,m] + sum_{k,m} h'[n,k,0,i-m] W[k,0,3-a,M-1-m]
    (full-border true convolution, i.e. the transpose of A1)."""
    N, K, _, Lh = h.shape
    M = W.shape[3]
    dtype = np.result_type(h.dtype, W.dtype)
    v = np.zeros((N, 4, Lh + M - 1), dtype=dtype)
    Wr = W[:, :, ::-1, ::-1]
    for m in range(M):
        v[:, :, m:m + Lh] += np.einsum('nkj,ka->naj', h[:, :, 0], W[:, 0, :, m])
        if hprime is not None:
            v[:, :, m:m + Lh] += np.einsum('nkj,ka->naj', hprime[:, :, 0], Wr[:, 0, :, m])
    v = v + c[0][None, :, None]
    return v[:, None]


# A6  convRBM.py:294-299,699-702  _topDownProbability / _softmax
def top_down_probability(act):
    e = np.exp(act)
    return e / e.sum(axis=2, keepdims=True)


# A6  convRBM.py:301-315  _topDownSample
def top_down_sample(prob, u):
    """One uniform per (n,i); rows ordered n-major (dimshuffle(0,1,3,2).reshape at convRBM.py:304-306)."""
    N, _, A, Lv = prob.shape
    rows = prob.transpose(0, 1, 3, 2).reshape(N * Lv, A)
    s = multinomial_from_uniform(rows, u)
    return s.reshape(N, 1, Lv, A).transpose(0, 1, 3, 2)


# A8  convRBM.py:327-337  _collectVHStatistics
def collect_vh(prob_h, V):
    N, K, _, Lh = prob_h.shape
    M = V.shape[3] - Lh + 1
    out = np.zeros((K, 1, 4, M), dtype=np.result_type(prob_h.dtype, V.dtype))
    for m in range(M):
        out[:, 0, :, m] = np.einsum('naj,nkj->ka', V[:, 0, :, m:m + Lh], prob_h[:, :, 0])
    return out / float(N * 1 * Lh)


# A9  convRBM.py:349-356 / 339-347
def collect_h(prob_h):
    return prob_h.mean(axis=(0, 2, 3))[None, :]


def collect_v(V):
    a = V.mean(axis=(0, 1, 3))[None, :]
    return a + a[:, ::-1]          # inc_subtensor: frequency PLUS its reverse (convRBM.py:345)


# A10 convRBM.py:358-371
def collect_update_statistics(prob_h, prob_hp, V):
    vh = collect_vh(prob_h, V)
    hh = collect_h(prob_h)
    if prob_hp is not None:
        vh = (vh + collect_vh(prob_hp, V)[:, :, ::-1, ::-1]) / 2.0
        hh = (hh + collect_h(prob_hp)) / 2.0
    return vh, hh, collect_v(V)


# A11 convRBM.py:440-451  _gradientSparsityConstraintEntropy, closed form of the two T.grad calls
def sparsity_gradient(D, W, b, pooling, rho):
    """R = -mean_k[rho log p_k + (1-rho) log(1-p_k)], p_k = mean_{n,j} P(h|D) on the FORWARD strand only.
    dR/dx[n,k,j] = g_k * Q[n,k,j], Q = p_j (1 - sum_{group} p_i), g_k = -(rho/p_k - (1-rho)/(1-p_k))/(K N Lh);
    dR/dW = g_k * sum Q (x) V, dR/db = g_k * sum Q."""
    p = bottom_up_probability(bottom_up_activity(D, W, b, False), pooling)
    N, K, _, Lh = p.shape
    pk = p.mean(axis=(0, 2, 3))
    gk = -(rho / pk - (1.0 - rho) / (1.0 - pk)) / float(K * N * Lh)
    grp = p.reshape(N, K, 1, Lh // pooling, pooling)
    Q = (grp * (1.0 - grp.sum(axis=4, keepdims=True))).reshape(N, K, 1, Lh)
    gW = collect_vh(Q, D) * float(N * Lh) * gk[:, None, None, None]
    gb = (Q.sum(axis=(0, 2, 3)) * gk)[None, :]
    return gW, gb


# A13 convRBM.py:657-697
def free_energy_per_motif(D, W, b, c, pooling, doublestranded):
    def part(flip):
        x = bottom_up_activity(D, W, b, flip)
        N, K, _, Lh = x.shape
        g = x.reshape(N, K, 1, Lh // pooling, pooling)
        return -np.sum(np.log(1.0 + np.sum(np.exp(g), axis=4)), axis=(2, 3))
    fe = part(False)
    if doublestranded:
        fe = fe + part(True)
    return fe - np.sum(D * c[0][None, None, :, None], axis=(1, 2, 3))[:, None]


def free_energy(D, W, b, c, pooling, doublestranded):
    fe = free_energy_per_motif(D, W, b, c, pooling, doublestranded)
    K = W.shape[0]
    cterm = np.sum(D * c[0][None, None, :, None], axis=(1, 2, 3))
    # per-motif broadcasts the c-term to every k (convRBM.py:695); the per-sequence form counts it once (:674)
    return (fe.sum(axis=1) + (K - 1) * cterm) / float(D.shape[3])


# A7 convRBM.py:507-514
def hit_probs(D, W, b, pooling, doublestranded):
    if doublestranded:
        return bottom_up_probability(bottom_up_activity(D, W, b, False), pooling)
    return bottom_up_probability(bottom_up_activity(D, W, b, False) + bottom_up_activity(D, W, b, True), pooling)


# A14 convRBM.py:475-485
def evaluate_params(W):
    e = np.exp(W)
    pwm = e / e.sum(axis=2, keepdims=True)
    ent = np.sum(-pwm * np.log2(pwm), axis=2)                 # (K,1,M)
    wnorm = np.sqrt(np.mean(W ** 2))
    ic = LOG2_4 - np.mean(ent)
    medic = LOG2_4 - np.mean(np.sort(ent, axis=2)[:, :, ent.shape[2] // 2])
    return wnorm, ic, medic


def get_pfms(W):
    """convRBM.py:640-655."""
    out = []
    for m in W:
        e = np.exp(np.asarray(m[0], dtype=np.float64))
        out.append(e / e.sum(axis=0, keepdims=True))
    return out


def truncate_for_pooling(L, M, P):
    """fit(): convRBM.py:586-591."""
    return int((L - M + 1) / P) * P + M - 1


def iterate_batch_indices(total, batch):
    """convRBM.py:722-726."""
    return [[i, i + batch] if i + batch <= total else [i, total] for i in range(total)[0::batch]]


class OracleCRBM(object):
    """State + train step (A12, convRBM.py:373-438) with injected uniforms."""

    FANTASY_LEN = 200  # convRBM.py:168

    def __init__(self, W, b, c, doublestranded=True, batchsize=20, learning_rate=0.1, momentum=0.95,
                 pooling=1, cd_k=5, rho=0.01, lambda_rate=0.1, dtype=np.float64):
        self.dtype = dtype
        self.W, self.b, self.c = _f(W, dtype).copy(), _f(b, dtype).copy(), _f(c, dtype).copy()
        self.vW, self.vb, self.vc = np.zeros_like(self.W), np.zeros_like(self.b), np.zeros_like(self.c)
        K = self.W.shape[0]
        self.ds, self.P, self.cd_k = bool(doublestranded), int(pooling), int(cd_k)
        self.lr, self.mu, self.rho, self.lam = learning_rate, momentum, rho, lambda_rate
        self.batchsize = batchsize
        self.fantasy_h = np.zeros((batchsize, K, 1, self.FANTASY_LEN), dtype=dtype)
        self.fantasy_hp = np.zeros((batchsize, K, 1, self.FANTASY_LEN), dtype=dtype) if self.ds else None

    def h_given_v(self, V, flip):
        return bottom_up_probability(bottom_up_activity(V, self.W, self.b, flip), self.P)

    def train_step(self, D, draw):
        """`draw(tag, n_rows)` returns the uniforms for one multinomial call; tags are 'v<i>', 'h<i>', 'hp<i>'
        for Gibbs iteration i. Returns a dict of the intermediate statistics (for parity tests)."""
        D = _f(D, self.dtype)
        # positive phase (convRBM.py:377-388); the positive-phase H sample is never used
        ph = self.h_given_v(D, False)
        php = self.h_given_v(D, True) if self.ds else None
        G_W_data, G_b_data, G_c_data = collect_update_statistics(ph, php, D)
        # negative phase from the persistent chain (convRBM.py:391-413)
        h, hp = self.fantasy_h, self.fantasy_hp
        for it in range(self.cd_k):
            pv = top_down_probability(top_down_activity(h, hp, self.W, self.c))
            v = top_down_sample(pv, draw('v%d' % it, pv.shape[0] * pv.shape[3]))
            pm = self.h_given_v(v, False)
            h = bottom_up_sample(pm, self.P, draw('h%d' % it, pm.size // self.P))
            if self.ds:
                pmp = self.h_given_v(v, True)
                hp = bottom_up_sample(pmp, self.P, draw('hp%d' % it, pmp.size // self.P))
            else:
                pmp, hp = None, None
        G_W_model, G_b_model, G_c_model = collect_update_statistics(pm, pmp, v)
        reg_W, reg_b = sparsity_gradient(D, self.W, self.b, self.P, self.rho)
        # momentum update (convRBM.py:415-436)
        self.vW = self.mu * self.vW + self.lr * (G_W_data - G_W_model - self.lam * reg_W)
        self.vb = self.mu * self.vb + self.lr * (G_b_data - G_b_model - self.lam * reg_b)
        self.vc = self.mu * self.vc + self.lr * (G_c_data - G_c_model)
        self.W, self.b, self.c = self.W + self.vW, self.b + self.vb, self.c + self.vc
        self.fantasy_h, self.fantasy_hp = h, hp
        return dict(G_W_data=G_W_data, G_b_data=G_b_data, G_c_data=G_c_data, G_W_model=G_W_model,
                    G_b_model=G_b_model, G_c_model=G_c_model, reg_W=reg_W, reg_b=reg_b, v_model=v)

    def evaluate_data(self, D, draw):
        """convRBM.py:466-472: [mean free energy, mean of a forward-strand H sample]."""
        D = _f(D, self.dtype)
        mfe = free_energy(D, self.W, self.b, self.c, self.P, self.ds).sum() / D.shape[0]
        p = self.h_given_v(D, False)
        h = bottom_up_sample(p, self.P, draw('eval_h', p.size // self.P))
        return mfe, h.mean()
