"""theano.sandbox.rng_mrg.MRG_RandomStreams.multinomial — TEST INFRASTRUCTURE ONLY.

Reproduces the op Theano lowers `multinomial(n=1, pvals)` to (`MultinomialFromUniform`): one uniform u per
row; the first outcome m whose running sum of pvals exceeds u is set to 1; if u >= the row sum the row stays
all-zero (this is how the sub-stochastic rows of convRBM.py:263-265 encode the "all off" pooling state).
The MRG31k3p generator itself is NOT reproduced: uniforms come from `uniform_source` (default: a NumPy
RandomState seeded like the stream), and every draw is appended to `draw_log` so fixtures can replay them.
"""
import numpy as np
import torch

from .._graph import Var, _ev

# Global hooks used by tests/golden/make_golden.py
uniform_source = None   # callable(node_index, n_rows) -> float64 array in [0,1), or None
draw_log = []           # list of dicts {node, uniforms, pvals}
_node_counter = [0]


def reset_hooks():
    global uniform_source
    uniform_source = None
    del draw_log[:]


class MRG_RandomStreams(object):
    def __init__(self, seed=12345, use_cuda=None):
        self._rs = np.random.RandomState(int(seed) % (2 ** 32))

    def multinomial(self, size=None, n=1, pvals=None, ndim=None, dtype='int64', nstreams=None):
        if n != 1:
            raise NotImplementedError("shim multinomial: n != 1")
        node = _node_counter[0]
        _node_counter[0] += 1
        rs = self._rs

        def fn(env):
            p = _ev(pvals, env)
            rows = p.shape[0]
            if uniform_source is not None:
                u = np.asarray(uniform_source(node, rows), dtype=np.float64)
            else:
                u = rs.random_sample(rows)
            draw_log.append({'node': node, 'uniforms': u.copy(), 'pvals': p.detach().cpu().numpy().copy()})
            ut = torch.from_numpy(u).to(p.dtype).unsqueeze(1)
            csum = torch.cumsum(p, dim=1)
            hit = csum > ut                                  # candidates
            first = hit & (torch.cumsum(hit.to(torch.int64), dim=1) == 1)
            return first.to(torch.int64)
        return Var(fn)

    def uniform(self, size, low=0.0, high=1.0, dtype=None):
        raise NotImplementedError("shim: uniform not used by convRBM.py")
