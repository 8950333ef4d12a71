"""Lazy expression graph evaluated with torch-CPU. TEST INFRASTRUCTURE ONLY (see package docstring)."""
import math
import numpy as np
import torch

_config = None  # set by the package __init__


def _torch_dtype(name=None):
    name = name or _config.floatX
    return {'float32': torch.float32, 'float64': torch.float64,
            'int32': torch.int32, 'int64': torch.int64}[str(name)]


def _ev(x, env):
    """Evaluate a Var, or pass a Python/NumPy constant through (recursing into tuples/slices)."""
    if isinstance(x, Var):
        return x._eval(env)
    if isinstance(x, (tuple, list)):
        return type(x)(_ev(e, env) for e in x)
    if isinstance(x, slice):
        return slice(_ev(x.start, env), _ev(x.stop, env), _ev(x.step, env))
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x))
    return x


def _index(t, idx):
    """NumPy-style basic indexing on a torch tensor / tuple, including negative-step slices."""
    if isinstance(t, tuple):
        return t[idx]
    if not isinstance(idx, tuple):
        idx = (idx,)
    out = t
    dim = 0
    for ix in idx:
        if isinstance(ix, slice):
            step = 1 if ix.step is None else int(ix.step)
            if step < 0:
                n = out.shape[dim]
                # reverse, then apply the mirrored positive-step slice
                start = None if ix.start is None else n - 1 - int(ix.start)
                stop = None if ix.stop is None else n - 1 - int(ix.stop)
                out = torch.flip(out, dims=(dim,))
                sl = [slice(None)] * out.dim()
                sl[dim] = slice(start, stop, -step)
                out = out[tuple(sl)]
            else:
                sl = [slice(None)] * out.dim()
                sl[dim] = slice(None if ix.start is None else int(ix.start),
                                None if ix.stop is None else int(ix.stop), step)
                out = out[tuple(sl)]
            dim += 1
        elif ix is None:
            out = out.unsqueeze(dim)
            dim += 1
        else:
            out = out.select(dim, int(ix))
    return out


class Var(object):
    """A node of the lazy graph. `fn(env)` computes the value; results are memoised per call in `env`."""

    def __init__(self, fn, name=None, sub_of=None):
        self._fn = fn
        self.name = name
        self._sub_of = sub_of  # (base Var, index) when created by __getitem__ (for inc_subtensor)

    def _eval(self, env):
        k = id(self)
        if k not in env:
            env[k] = self._fn(env)
        return env[k]

    # ---- arithmetic -------------------------------------------------------------------------
    def _bin(self, other, op):
        return Var(lambda env: op(_ev(self, env), _ev(other, env)))

    def _rbin(self, other, op):
        return Var(lambda env: op(_ev(other, env), _ev(self, env)))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._rbin(o, lambda a, b: a + b)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._rbin(o, lambda a, b: a - b)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._rbin(o, lambda a, b: a * b)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._rbin(o, lambda a, b: a / b)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __floordiv__(self, o): return self._bin(o, lambda a, b: a // b)
    def __rfloordiv__(self, o): return self._rbin(o, lambda a, b: a // b)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return Var(lambda env: -_ev(self, env))

    # ---- tensor methods -----------------------------------------------------------------------
    @property
    def shape(self):
        return Var(lambda env: tuple(int(s) for s in _ev(self, env).shape))

    def __getitem__(self, idx):
        return Var(lambda env: _index(_ev(self, env), _ev(idx, env)), sub_of=(self, idx))

    def reshape(self, shape, ndim=None):
        return Var(lambda env: _ev(self, env).reshape(tuple(int(s) for s in _ev(tuple(shape), env))))

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (tuple, list)):
            pattern = tuple(pattern[0])

        def fn(env):
            t = _ev(self, env)
            kept = [p for p in pattern if p != 'x']
            dropped = [d for d in range(t.dim()) if d not in kept]
            for d in dropped:
                if t.shape[d] != 1:
                    raise ValueError("dimshuffle: cannot drop non-broadcastable dimension %d" % d)
            t = t.permute(*(kept + dropped))
            t = t.reshape(t.shape[:len(kept)])
            for pos, p in enumerate(pattern):
                if p == 'x':
                    t = t.unsqueeze(pos)
            return t
        return Var(fn)

    def astype(self, dtype):
        return Var(lambda env: _ev(self, env).to(_torch_dtype(dtype)))

    def sum(self, axis=None, keepdims=False):
        from .tensor import sum as _sum
        return _sum(self, axis=axis, keepdims=keepdims)

    def mean(self, axis=None, keepdims=False):
        from .tensor import mean as _mean
        return _mean(self, axis=axis, keepdims=keepdims)


class SharedVariable(Var):
    """theano.shared: a graph leaf holding a NumPy value (convRBM.py:133 etc.)."""

    def __init__(self, value, name=None, borrow=False):
        Var.__init__(self, None, name=name)
        self._value = np.array(value, copy=not borrow)

    def _eval(self, env):
        k = id(self)
        if k not in env:
            env[k] = torch.from_numpy(np.ascontiguousarray(self._value))
        return env[k]

    def get_value(self, borrow=False):
        return self._value if borrow else self._value.copy()

    def set_value(self, value, borrow=False):
        self._value = np.array(value, copy=not borrow)


def shared(value, name=None, borrow=False, **kwargs):
    return SharedVariable(value, name=name, borrow=borrow)


class _Function(object):
    """theano.function: NumPy in, NumPy out; `updates` are all computed before any is assigned."""

    def __init__(self, inputs, outputs=None, updates=None, givens=None, name=None, **kwargs):
        self.inputs = list(inputs)
        self.outputs = outputs
        self.updates = list(updates.items()) if isinstance(updates, dict) else list(updates or [])
        self.givens = list(givens.items()) if isinstance(givens, dict) else list(givens or [])
        self.name = name

    def __call__(self, *args):
        if len(args) != len(self.inputs):
            raise TypeError("%s: expected %d inputs, got %d" % (self.name, len(self.inputs), len(args)))
        env = {}
        for var, val in zip(self.inputs, args):
            env[id(var)] = torch.from_numpy(np.ascontiguousarray(np.asarray(val)))
        for var, repl in self.givens:
            env[id(var)] = _ev(repl, env)
        env['__bound__'] = dict(env)  # input bindings, reused by grad nodes for their private sub-evaluation
        with torch.no_grad():
            if self.outputs is None:
                outs = None
            elif isinstance(self.outputs, (list, tuple)):
                outs = [_to_numpy(_ev(o, env)) for o in self.outputs]
            else:
                outs = _to_numpy(_ev(self.outputs, env))
            new_vals = [(sv, _to_numpy(_ev(nv, env))) for sv, nv in self.updates]
        for sv, val in new_vals:
            sv.set_value(val.astype(sv._value.dtype, copy=False))
        return outs


def _to_numpy(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy().copy()
    return np.asarray(v)


def function(inputs, outputs=None, updates=None, givens=None, name=None, **kwargs):
    return _Function(inputs, outputs, updates=updates, givens=givens, name=name, **kwargs)
