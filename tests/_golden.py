"""Helpers to load the committed reference fixtures (tests/golden/case_*.npz, made by make_golden.py)."""
import ast
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')

CASE_NAMES = sorted(os.path.basename(p)[len('case_'):-len('.npz')]
                    for p in glob.glob(os.path.join(GOLDEN, 'case_*.npz')))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, 'case_%s.npz' % name))
    case = {k: z[k] for k in z.files}
    case['hyper'] = ast.literal_eval(str(case['hyper']))
    return case


def onehot(codes, dtype=np.float32):
    """(N,L) base codes -> (N,1,4,L) one-hot, the layout of secomo.seqToOneHot (sequences.py:20-31)."""
    N, L = codes.shape
    D = np.zeros((N, 1, 4, L), dtype=dtype)
    for a in range(4):
        D[:, 0, a, :] = (codes == a)
    return D


def draw_for(case, step):
    def draw(tag, rows):
        u = case['u_%s_%s' % (step, tag)]
        assert u.shape[0] == rows, (tag, u.shape, rows)
        return u.astype(np.float64)
    return draw
