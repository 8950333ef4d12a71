"""CPU-only tests of the host side: the secomo-compatible interface (constructor validation, persistence, sequence
input) and the C-ABI library (loads, exports every declared symbol). No compute calls: there is no CPU fallback."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

import crbm_b200
from crbm_b200 import CRBM, _lib, load_sample, readSeqsFromFasta, seqToOneHot, splitTrainingTest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'crbm_b200.h')).read()
    declared = set(re.findall(r'\b(crbm_[a-z0-9_]+)\s*\(', header))
    declared -= {'crbm_train_args', 'crbm_stats_layout'}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.crbm_abi_version() == 1
    assert lib.crbm_packed_row_words(200) % 4 == 0 and lib.crbm_packed_row_words(200) >= 1 + 13 + 3
    assert lib.crbm_hidden_words(100) == 4
    assert lib.crbm_train_stats_floats(100, 15) == 3 * 6000 + 4 * 100 + 8 + 4
    lay = _lib.StatsLayout()
    assert lib.crbm_train_stats_layout(100, 15, ctypes.byref(lay)) == 0 and lay.total == lib.crbm_train_stats_floats(100, 15)


def test_struct_mirror_matches_header_field_order():
    header = open(os.path.join(ROOT, 'include', 'crbm_b200.h')).read()
    body = header[header.index('typedef struct crbm_train_args {'):header.index('} crbm_train_args;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for stmt in body.split('{', 1)[1].split(';'):
        stmt = stmt.strip()
        if not stmt:
            continue
        parts = stmt.split(',')
        for i, p in enumerate(parts):
            names.append(re.findall(r'([A-Za-z_][A-Za-z0-9_]*)\s*$', p.strip())[0])
    assert names == [f[0] for f in _lib.TrainArgs._fields_]


def test_crbm_parametervalues():
    """Constructor validation (reference tests/testcrbm.py:14-56)."""
    bad = [dict(num_motifs=0, motif_length=5), dict(num_motifs=1, motif_length=0),
           dict(num_motifs=1, motif_length=1, epochs=-1), dict(num_motifs=1, motif_length=1, input_dims=-1),
           dict(num_motifs=1, motif_length=1, batchsize=0), dict(num_motifs=1, motif_length=1, learning_rate=0.0),
           dict(num_motifs=1, motif_length=1, momentum=-1.), dict(num_motifs=1, motif_length=1, momentum=1.1),
           dict(num_motifs=1, motif_length=1, pooling=0), dict(num_motifs=1, motif_length=1, cd_k=0),
           dict(num_motifs=1, motif_length=1, rho=-1.), dict(num_motifs=1, motif_length=1, rho=1.1),
           dict(num_motifs=1, motif_length=1, lambda_rate=-.1)]
    for kw in bad:
        with pytest.raises(Exception):
            CRBM(**kw)
    with pytest.warns(UserWarning):
        CRBM(num_motifs=1, motif_length=1, input_dims=3)


def test_crbm_creation_and_reload(tmpdir):
    """Shapes, bias initialisation and save/load round trip (reference tests/testcrbm.py:58-98, convRBM.py:143-148)."""
    from scipy.stats import norm
    model = CRBM(num_motifs=2, motif_length=5)
    W, b, c = model.motifs.get_value(), model.bias.get_value(), model.c.get_value()
    assert W.shape == (2, 1, 4, 5) and b.shape == (1, 2) and c.shape == (1, 4)
    np.testing.assert_allclose(b, norm.ppf(0.01, 0, np.sqrt(5)), rtol=1e-6)
    np.testing.assert_array_equal(c, 0)
    path = os.path.join(tmpdir.strpath, "model.pkl")
    model.saveModel(path)
    model2 = CRBM.loadModel(path)
    for attr in ('num_motifs', 'motif_length', 'epochs', 'input_dims', 'doublestranded', 'batchsize', 'momentum',
                 'pooling', 'cd_k', 'rho', 'lambda_rate', 'learning_rate', 'spmethod'):
        assert getattr(model, attr) == getattr(model2, attr)
    np.testing.assert_allclose(model.motifs.get_value(), model2.motifs.get_value())
    np.testing.assert_allclose(model.bias.get_value(), model2.bias.get_value())
    np.testing.assert_allclose(model.c.get_value(), model2.c.get_value())
    # pickle layout is the reference's: ((W,b,c), 13 hyper-parameters) (convRBM.py:186-204)
    import joblib
    params, hyper = joblib.load(path)
    assert len(params) == 3 and len(hyper) == 13 and hyper[-1] == 'entropy' and hyper[0] == 2 and hyper[1] == 5


def test_rho_auto_and_pfms():
    m = CRBM(num_motifs=10, motif_length=15, rho=0)
    assert m.rho == pytest.approx(1. / (10 * 15) / 2.)
    m = CRBM(num_motifs=10, motif_length=15, rho=0, doublestranded=False)
    assert m.rho == pytest.approx(1. / (10 * 15))
    pfms = m.getPFMs()
    assert len(pfms) == 10
    for p in pfms:
        assert p.shape == (4, 15)
        np.testing.assert_allclose(p.sum(), 15)
    assert 'Number of motifs: 10' in repr(m)


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = CRBM(num_motifs=2, motif_length=5)
    D = seqToOneHot(['ACGTACGTAC'])
    for call in (lambda: m.fit(D), lambda: m.motifHitProbs(D), lambda: m.freeEnergy(D), lambda: m._bottomUpActivity(D)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_load_sample_shape_and_onehot():
    x = load_sample()
    assert x.shape == (3997, 1, 4, 200) and x.dtype == np.float32      # reference tests/testloadsamples.py:6
    np.testing.assert_array_equal(x.sum(axis=2), 1.0)
    # first record of the Oct4 sample starts CACAGCAGAG
    first = ''.join('ACGT'[i] for i in x[0, 0].argmax(axis=0)[:10])
    assert first == 'CACAGCAGAG'


def test_fasta_reader_and_onehot(tmpdir, capsys):
    path = os.path.join(tmpdir.strpath, 'x.fa')
    with open(path, 'w') as f:
        f.write(">s1 first\nACGT\nacgt\n>s2\nACNTACGT\n>s3\nTTTTGGGG\n\n")
    seqs = readSeqsFromFasta(path)
    assert [s.id for s in seqs] == ['s1', 's3']
    assert "skip sequence containing N" in capsys.readouterr().out
    oh = seqToOneHot(seqs)
    assert oh.shape == (2, 1, 4, 8) and oh.dtype == np.float32
    np.testing.assert_array_equal(oh[0, 0].argmax(axis=0), [0, 1, 2, 3, 0, 1, 2, 3])
    np.testing.assert_array_equal(oh[1, 0].argmax(axis=0), [3, 3, 3, 3, 2, 2, 2, 2])
    with pytest.raises(KeyError):
        seqToOneHot(['ACGU'])


def test_split_training_test(tmpdir):
    """reference tests/testsequences.py: the non-random split concatenates back to the original; random keeps counts."""
    src = crbm_b200.sample_fasta_path()
    path = os.path.join(tmpdir.strpath, 'oct4.fa')
    with open(src) as f, open(path, 'w') as g:
        g.write(''.join(f.readlines()[:400]))
    splitTrainingTest(path, 0.1, randomize=False)
    tr = seqToOneHot(readSeqsFromFasta(os.path.join(tmpdir.strpath, 'oct4_train.fa')))
    te = seqToOneHot(readSeqsFromFasta(os.path.join(tmpdir.strpath, 'oct4_test.fa')))
    full = seqToOneHot(readSeqsFromFasta(path))
    np.testing.assert_array_equal(np.concatenate([te, tr]), full)
    splitTrainingTest(path, 0.25, num_top_regions=100)
    tr = readSeqsFromFasta(os.path.join(tmpdir.strpath, 'oct4_train.fa'))
    te = readSeqsFromFasta(os.path.join(tmpdir.strpath, 'oct4_test.fa'))
    assert len(te) == 25 and len(tr) == 75


def test_batch_indices_and_truncation():
    m = CRBM(num_motifs=2, motif_length=20, pooling=5, batchsize=20)
    assert m._iterateBatchIndices(45, 20) == [[0, 20], [20, 40], [40, 45]]
    assert m._truncated_length(1000) == 999


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm: oracle/crbm_torch_cpu.py on the host cores) prints exactly one JSON line on
    stdout with the keys the driver reads; nothing else may reach stdout."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1'],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'nt/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': 'nt/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('C2')


def test_save_motifs_jaspar_and_pfm(tmpdir):
    """secomo/utils.py:16-47 `saveMotifs`: one `<name><i>.pfm` per motif in the layout of Bio.motifs.jaspar.write
    (header `>id name`, rows `A [ ... ]` with %6.2f values); reference tests/testtutorial.py counts the files."""
    from crbm_b200 import saveMotifs, formatMotif

    class Stub(object):
        def getPFMs(self):
            rs = np.random.RandomState(0)
            out = []
            for _ in range(3):
                x = rs.rand(4, 5)
                out.append(x / x.sum(axis=0, keepdims=True))
            return out

    d = os.path.join(tmpdir.strpath, 'pfms', 'nested')
    saveMotifs(Stub(), d)
    files = sorted(os.listdir(d))
    assert files == ['mot0.pfm', 'mot1.pfm', 'mot2.pfm']
    lines = open(os.path.join(d, 'mot1.pfm')).read().splitlines()
    assert lines[0] == '>mot2 mot2' and [l[0] for l in lines[1:]] == list('ACGT')
    vals = np.array([[float(v) for v in l[l.index('[') + 1:l.index(']')].split()] for l in lines[1:]])
    np.testing.assert_allclose(vals, Stub().getPFMs()[1], atol=0.005)
    np.testing.assert_allclose(vals.sum(axis=0), 1.0, atol=0.02)
    pfm = formatMotif(Stub().getPFMs()[0], 'x', 'pfm').splitlines()
    assert len(pfm) == 4 and all(len(r.split()) == 5 for r in pfm)
    with pytest.raises(ValueError):
        saveMotifs(Stub(), d, fformat='tab')


def test_host_packing_matches_the_device_row_layout():
    """packCodesHost: 2 bit/base rows in the layout the kernels read (front pad word, data words, zero padding, stride a multiple
    of 4 words = crbm_packed_row_words); round trip through unpackCodesHost; ragged tail words zero-padded."""
    from crbm_b200 import packCodesHost, unpackCodesHost, packedRowWords
    lib = _lib.load()
    rs = np.random.RandomState(3)
    for Ln in (1, 15, 16, 17, 200, 999):
        assert packedRowWords(Ln) == lib.crbm_packed_row_words(Ln)
        codes = rs.randint(0, 4, size=(5, Ln)).astype(np.uint8)
        words = packCodesHost(codes)
        assert words.shape == (5, packedRowWords(Ln)) and words.dtype == np.uint32
        assert not words[:, 0].any() and not words[:, 1 + (Ln + 15) // 16:].any()
        for n, i in ((0, 0), (4, Ln - 1), (2, Ln // 2)):
            assert (int(words[n, 1 + i // 16]) >> (2 * (i % 16))) & 3 == codes[n, i]
        if Ln % 16:
            assert int(words[0, (Ln + 15) // 16]) >> (2 * (Ln % 16)) == 0
        np.testing.assert_array_equal(unpackCodesHost(words, Ln), codes)
    with pytest.raises(ValueError):
        packCodesHost(np.array([[0, 4]], dtype=np.uint8))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under crbm_b200/ may import, call or execute anything under oracle/ (and there is
    no CPU fallback to route to: compute entry points raise without a CUDA device, see test_compute_fails_loudly_without_gpu)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'crbm_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn), errors='replace').read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), fn
                assert 'crbm_oracle' not in text and 'crbm_torch_cpu' not in text, fn


def test_header_is_plain_c(tmpdir):
    """include/crbm_b200.h is the drop-in boundary: it must compile as C99 and as C++ with no CUDA or torch headers."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(tmpdir.strpath, 'hdr.c')
    with open(src, 'w') as f:
        f.write('#include "crbm_b200.h"\nint main(void){ crbm_train_args a; crbm_p2p_args p; crbm_stats_layout l; (void)a; (void)p; (void)l;'
                ' return CRBM_PHASE_APPLY == 8 ? 0 : 1; }\n')
    inc = os.path.join(root, 'include')
    if shutil.which('gcc'):
        r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', inc, '-fsyntax-only', src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    if shutil.which('g++'):
        r = subprocess.run(['g++', '-std=c++17', '-Wall', '-I', inc, '-fsyntax-only', '-x', 'c++', src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
