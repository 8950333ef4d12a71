"""theano.tensor subset used by convRBM.py — TEST INFRASTRUCTURE ONLY (see theano/__init__.py)."""
import math
import numpy as np
import torch

from .._graph import Var, SharedVariable, _ev, _torch_dtype
from . import nnet  # noqa: F401
from . import shared_randomstreams  # noqa: F401


def tensor4(name=None, dtype=None):
    def fn(env):
        raise RuntimeError("unbound symbolic input %r" % name)
    return Var(fn, name=name)


matrix = vector = tensor3 = tensor4


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (tuple, list)):
        return tuple(int(a) for a in axis)
    return (int(axis),)


def sum(x, axis=None, keepdims=False):  # noqa: A001
    def fn(env):
        t = _ev(x, env)
        ax = _axes(axis)
        return t.sum() if ax is None else t.sum(dim=ax, keepdim=keepdims)
    return Var(fn)


def mean(x, axis=None, keepdims=False):
    def fn(env):
        t = _ev(x, env)
        ax = _axes(axis)
        return t.mean() if ax is None else t.mean(dim=ax, keepdim=keepdims)
    return Var(fn)


def prod(x, axis=None):
    def fn(env):
        v = _ev(x, env)
        if isinstance(v, tuple):
            out = 1
            for e in v:
                out *= int(e)
            return out
        return v.prod() if axis is None else v.prod(dim=axis)
    return Var(fn)


def _unary(tfn, sfn):
    def op(x):
        def fn(env):
            v = _ev(x, env)
            return tfn(v) if isinstance(v, torch.Tensor) else sfn(v)
        return Var(fn)
    return op


exp = _unary(torch.exp, math.exp)
log = _unary(torch.log, math.log)
log2 = _unary(torch.log2, math.log2)
sqrt = _unary(torch.sqrt, math.sqrt)


def maximum(a, b):
    return Var(lambda env: torch.maximum(_ev(a, env), _ev(b, env)))


def cast(x, dtype):
    return Var(lambda env: _ev(x, env).to(_torch_dtype(dtype)))


def sort(x, axis=-1):
    return Var(lambda env: torch.sort(_ev(x, env), dim=axis).values)


def inc_subtensor(sub, y):
    """T.inc_subtensor(a[idx], y): a copy of `a` with y added into a[idx] (convRBM.py:345)."""
    base, idx = sub._sub_of

    def fn(env):
        t = _ev(base, env).clone()
        ix = _ev(idx, env)
        t[ix] = t[ix] + _ev(y, env)
        return t
    return Var(fn)


def grad(cost, wrt):
    """T.grad(cost, wrt) for a SharedVariable `wrt` (convRBM.py:447-450): torch.autograd on a private
    re-evaluation of the cost sub-graph with `wrt` as a differentiable leaf."""
    if not isinstance(wrt, SharedVariable):
        raise NotImplementedError("shim grad: wrt must be a shared variable")

    def fn(env):
        sub = dict(env.get('__bound__', {}))
        sub['__bound__'] = dict(sub)
        with torch.enable_grad():
            leaf = torch.from_numpy(np.ascontiguousarray(wrt._value)).clone().requires_grad_(True)
            sub[id(wrt)] = leaf
            c = _ev(cost, sub)
            (g,) = torch.autograd.grad(c, leaf)
        return g.detach()
    return Var(fn)
