#!/bin/bash
# Runs the reference's OWN test-suite for the hot path (tests/testcrbm.py, tests/testloadsamples.py), unmodified,
# from /root/reference on top of oracle/theano_shim (+ import stubs for the absent plotting/Biopython deps).
# Build-container only: /root/reference does not exist on the GPU boxes. Its NumPy control loops
# (testcrbm.py:182-189, :202-220) are what validate the shim's conv2d / softmax / multinomial semantics.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
cd /tmp
PYTHONPATH="$here/theano_shim:$here/ref_stubs:/root/reference" python -m pytest -q -p no:cacheprovider \
    /root/reference/tests/testcrbm.py /root/reference/tests/testloadsamples.py "$@"
