"""Minimal lazy-graph stand-in for the `theano` package — TEST INFRASTRUCTURE ONLY.

Why this exists: the reference (schulter/crbm, `/root/reference/secomo/convRBM.py`) does all of its
arithmetic through Theano, which is neither installed nor installable offline in this image. To pin the
NumPy oracle (`oracle/crbm_oracle.py`) against the reference's OWN code rather than against a second
restatement, this package implements exactly the subset of the Theano API that `convRBM.py` and
`tests/testcrbm.py` touch, with Theano's documented op semantics, on top of torch-CPU tensors:

  theano.shared / .get_value / .set_value / .dimshuffle        (convRBM.py:133-173, 233-235, 242, 279)
  theano.function(inputs, outputs, updates=, givens=)          (convRBM.py:459-511)
  theano.tensor: tensor4, sum, mean, exp, log, log2, sqrt, prod, cast, sort, grad, inc_subtensor
  theano.tensor.nnet.conv2d(border_mode in {valid, full}, filter_flip)   (convRBM.py:241,280,285,333)
  theano.sandbox.rng_mrg.MRG_RandomStreams.multinomial(n=1, pvals)       (convRBM.py:265,307)

The unmodified reference file is then imported from /root/reference by `tests/golden/make_golden.py`
(in the build container only) to generate the committed fixtures under `tests/golden/`, and the reference's
own test-suite (`/root/reference/tests/testcrbm.py`) is run against it by `oracle/run_reference_tests.sh`
— its NumPy control loops validate this shim's conv2d semantics.

It is NOT Theano: graph optimisation, C code generation and the MRG31k3p generator are not reproduced. The
multinomial op reproduces `MultinomialFromUniform` (first outcome whose running sum exceeds the uniform;
an all-zero row when the uniform is >= the row sum) and draws its uniforms from an injectable source so
fixtures can record them.

Nothing under crbm_b200/ may import this package.
"""
import math
import numpy as np
import torch

from . import _graph
from ._graph import Var, SharedVariable, function, shared  # noqa: F401


class _Config(object):
    floatX = 'float32'


config = _Config()
_graph._config = config

from . import tensor  # noqa: E402,F401
from . import sandbox  # noqa: E402,F401

__version__ = '1.0.5-shim'
