#!/usr/bin/env python
"""Writes tests/golden/reference_model.pkl with the UNMODIFIED reference's own `CRBM.saveModel`
(/root/reference/secomo/convRBM.py:177-204, joblib pickle protocol 2) running on `oracle/theano_shim`, plus
tests/golden/reference_model_expected.npz (the parameters and hyper-parameters that went in). Build-container only.

    python tests/golden/make_reference_pickle.py

The pickle is the interchange fixture of SURVEY 8 f-3: `crbm_b200.CRBM.loadModel` must read a file the reference
wrote (tests/test_host_cpu.py), and the reference's `loadModel` must read ours (checked here when run).
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'theano_shim'))
sys.path.insert(0, ROOT)

import theano  # noqa: E402  (the shim)

theano.config.floatX = 'float32'
spec = importlib.util.spec_from_file_location('ref_convRBM', '/root/reference/secomo/convRBM.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

HYPER = dict(num_motifs=7, motif_length=9, epochs=3, input_dims=4, doublestranded=False, batchsize=11,
             learning_rate=0.05, momentum=0.9, pooling=2, cd_k=3, rho=0.02, lambda_rate=0.25)


def main():
    rs = np.random.RandomState(20261018)
    model = ref.CRBM(**HYPER)
    W = rs.randn(7, 1, 4, 9).astype(np.float32)
    b = rs.randn(1, 7).astype(np.float32)
    c = rs.randn(1, 4).astype(np.float32)
    model.motifs.set_value(W)
    model.bias.set_value(b)
    model.c.set_value(c)
    out = os.path.join(HERE, 'reference_model.pkl')
    model.saveModel(out)
    np.savez(os.path.join(HERE, 'reference_model_expected.npz'), W=W, b=b, c=c,
             hyper=np.array(repr(dict(HYPER, spmethod='entropy'))))
    # the other direction: a file written by crbm_b200 is read by the reference's own loadModel
    import crbm_b200
    import tempfile
    ours = crbm_b200.CRBM(5, 6, epochs=2, doublestranded=True, batchsize=4, pooling=3, cd_k=2, rho=0.03, device='cpu')
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'ours.pkl')
        ours.saveModel(path)
        back = ref.CRBM.loadModel(path)
    np.testing.assert_array_equal(back.motifs.get_value(), ours.motifs.get_value())
    np.testing.assert_array_equal(back.bias.get_value(), ours.bias.get_value())
    assert (back.num_motifs, back.motif_length, back.pooling, back.cd_k, back.doublestranded) == (5, 6, 3, 2, True)
    print("wrote", out, os.path.getsize(out), "bytes; the reference's loadModel read a crbm_b200 pickle")


if __name__ == '__main__':
    main()
