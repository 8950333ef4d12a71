"""theano.tensor.nnet.conv2d — TEST INFRASTRUCTURE ONLY (see theano/__init__.py).

Semantics restated from the Theano 1.0 documentation of `theano.tensor.nnet.conv2d`:
input (batch, in_ch, rows, cols); filters (out_ch, in_ch, frows, fcols);
border_mode 'valid' -> output rows-frows+1 x cols-fcols+1; 'full' -> rows+frows-1 x cols+fcols-1 (zero padded);
filter_flip=True (default) flips the filters along both spatial axes before sliding (a true convolution),
filter_flip=False is a cross-correlation. These are exactly what the reference's tests pin with NumPy control
loops (tests/testcrbm.py:182-189 for valid/flip, :202-220 for full).
"""
import torch
import torch.nn.functional as F

from .._graph import Var, _ev


def conv2d(input, filters, input_shape=None, filter_shape=None, border_mode='valid',
           subsample=(1, 1), filter_flip=True, **kwargs):
    if subsample != (1, 1):
        raise NotImplementedError("shim conv2d: subsample")

    def fn(env):
        x = _ev(input, env)
        w = _ev(filters, env)
        if filter_flip:
            w = torch.flip(w, dims=(2, 3))
        if border_mode == 'valid':
            pad = (0, 0)
        elif border_mode == 'full':
            pad = (w.shape[2] - 1, w.shape[3] - 1)
        else:
            raise NotImplementedError("shim conv2d: border_mode=%r" % (border_mode,))
        dt = torch.promote_types(x.dtype, w.dtype)
        return F.conv2d(x.to(dt), w.to(dt), padding=pad)
    return Var(fn)
