"""The secomo.CRBM call sequences the reference's own tests and tutorial exercise (reference tests/testcrbm.py:100-146,
257-363, 507-519; secomo/tutorial.py:13-44), run on the GPU path, plus graph-replay == eager training."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def data():
    import crbm_b200
    return crbm_b200.load_sample()


def test_training_runs_both_strand_modes(data, capsys):
    """reference test_training: fit(data[:100]) and fit(data, data[:10]) for ds in (True, False)."""
    from crbm_b200 import CRBM
    for ds in (True, False):
        model = CRBM(num_motifs=10, motif_length=15, doublestranded=ds, epochs=1, seed=4)
        w0 = model.motifs.get_value()
        model.fit(data[:100])
        model.fit(data[:95], data[:10])        # 95: exercises the short last batch
        assert np.all(np.isfinite(model.motifs.get_value())) and not np.allclose(w0, model.motifs.get_value())
    out = capsys.readouterr().out
    assert "BatchSize: 20" in out and "Epoch 0: FE=" in out and "Training finished after:" in out
    assert "NumH=" in out and "WNorm=" in out and "medIC=" in out


def test_motifhitprobs_dims_and_free_energy_dims(data):
    from crbm_b200 import CRBM
    d = data[:100]
    for ds in (True, False):
        model = CRBM(num_motifs=10, motif_length=15, doublestranded=ds, epochs=1)
        pred = model.motifHitProbs(d)
        assert pred.shape == (100, 10, 1, 200 - 15 + 1) and pred.dtype == np.float32
        assert np.all((pred >= 0) & (pred <= 1))
        assert model.freeEnergy(d[:11]).shape == (11,)
        assert model.freeEnergy(d[:11], permotif=True).shape == (11, 10)
        assert model.getHitProbs(d[:3]).shape == (3, 10, 1, 186)


def test_topdown_full_pipeline_properties(data):
    """reference test_topdown_*_full: V | H(D) has the data's shape, probabilities and samples sum to N*L."""
    from crbm_b200 import CRBM
    d = data[:11]
    model = CRBM(num_motifs=10, motif_length=5, seed=5)
    _, h1 = model._computeHgivenV(d, False)
    _, h2 = model._computeHgivenV(d, True)
    for hp in (None, h2):
        p, s = model._computeVgivenH(h1, hp)
        assert p.shape == d.shape and s.shape == d.shape
        np.testing.assert_allclose(p.sum(), d.shape[0] * d.shape[3], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(s.sum(), d.shape[0] * d.shape[3], rtol=1e-4, atol=1e-4)
    zeros = np.zeros((11, 10, 1, 196), dtype=np.float32)
    p, _ = model._computeVgivenH(zeros, zeros)
    np.testing.assert_allclose(p, 0.25, rtol=1e-5, atol=1e-5)     # c == 0 at construction


def test_host_packed_rows_and_fasta_ingest_equal_device_packing(tmpdir):
    """rows packed on the host (packCodesHost) are bit-identical to crbm_pack_codes / crbm_pack_onehot output, and
    PackedSequences.from_fasta == packing the one-hot array of the same file (N-containing records skipped, case folded)."""
    import crbm_b200
    from crbm_b200 import PackedSequences, packCodesHost
    rs = np.random.RandomState(5)
    for Ln in (17, 200, 999):
        codes = rs.randint(0, 4, size=(9, Ln)).astype(np.uint8)
        dev = PackedSequences.from_codes(codes)
        host = PackedSequences.from_packed_host(packCodesHost(codes), Ln)
        np.testing.assert_array_equal(host.words.cpu().numpy(), dev.words.cpu().numpy())
    path = os.path.join(tmpdir.strpath, 'x.fa')
    with open(path, 'w') as f:
        f.write('>a desc\nACGTacgtAC\n>b\nACGTNCGTAC\n>c\nttttGGGGcc\n')
    got = PackedSequences.from_fasta(path)
    ref = PackedSequences.from_onehot(crbm_b200.seqToOneHot(crbm_b200.readSeqsFromFasta(path)))
    assert got.N == 2 and got.L == 10
    np.testing.assert_array_equal(got.words.cpu().numpy(), ref.words.cpu().numpy())


def test_tutorial_function(tmpdir):
    """crbm_b200.tutorial.tutorial == secomo/tutorial.py without the plots (reference tests/testtutorial.py checks the outputs)."""
    from crbm_b200.tutorial import tutorial
    path = tmpdir.mkdir('run')
    out = tutorial(path.strpath)
    outputs = [o.basename for o in path.listdir()]
    assert 'oct4_model_params.pkl' in outputs and 'pfms' in outputs and 'motif_hit_summary.npz' in outputs
    assert len(os.listdir(os.path.join(path.strpath, 'pfms'))) == 10
    assert out['matches'].shape == (100, 10, 1, 186) and out['summary']['max'].shape == (3997, 10)
    assert out['positional_density'].shape == (10, 186)
    np.testing.assert_allclose(out['summary']['max'][:100], out['matches'][:, :, 0, :].max(axis=2), rtol=2e-6, atol=2e-7)


def test_tutorial_flow(tmpdir, data):
    """secomo/tutorial.py:13-44 without the plotting calls."""
    from crbm_b200 import CRBM
    model = CRBM(num_motifs=10, motif_length=15, epochs=1, seed=6)
    model.fit(data)
    path = os.path.join(tmpdir.strpath, 'oct4_model_params.pkl')
    model.saveModel(path)
    model2 = CRBM.loadModel(path)
    np.testing.assert_array_equal(model.motifs.get_value(), model2.motifs.get_value())
    assert len(model2.getPFMs()) == 10
    matches = model2.motifHitProbs(data[:100])
    assert matches.shape == (100, 10, 1, 186)
    np.testing.assert_array_equal(matches, model.motifHitProbs(data[:100]))


def test_graph_replay_equals_eager_steps(data):
    """Replaying the captured step graph must give bit-identical parameters to launching every step eagerly."""
    import crbm_b200
    from crbm_b200 import CRBM
    seqs = crbm_b200.PackedSequences.from_onehot(data[:20 * 13 + 7])
    out = []
    for graphs in (True, False):
        np.random.seed(1)
        m = CRBM(num_motifs=12, motif_length=9, epochs=2, cd_k=2, seed=77)
        m._use_graphs = graphs
        m.fit(seqs)
        out.append((m.motifs.get_value(), m.bias.get_value(), m.c.get_value(), m.fantasy_h.get_value()))
    for a, b in zip(*out):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('P,cd_k', [(1, 1), (1, 3), (4, 2)])
def test_step_issued_in_phases_equals_whole_step(data, P, cd_k):
    """crbm_train_phases: NEGATIVE, then POSITIVE | REDUCE, then APPLY as three calls == crbm_train_step, bit for bit
    (the partial-sum buffers are fixed, so the parts of a step may be issued separately)."""
    import torch
    import crbm_b200
    from crbm_b200 import CRBM, _lib
    lib = _lib.load()
    Lt = 15 - 1 + (186 // P) * P
    seqs = crbm_b200.PackedSequences.from_onehot(data[:60, :, :, :Lt])
    out = []
    for split in (False, True):
        np.random.seed(2)
        m = CRBM(num_motifs=20, motif_length=15, pooling=P, cd_k=cd_k, seed=9, batchsize=20)
        for i in range(3):
            a = m._fill_args(seqs.words.data_ptr() + 20 * i * seqs.row_words * 4, seqs.row_words, 20, seqs.L)
            a.step = i
            s = _lib.stream_handle()
            if split:
                _lib.check(lib.crbm_train_phases(a, _lib.PHASE_NEGATIVE, s))
                _lib.check(lib.crbm_train_phases(a, _lib.PHASE_POSITIVE | _lib.PHASE_REDUCE, s))
                _lib.check(lib.crbm_train_phases(a, _lib.PHASE_APPLY, s))
            else:
                _lib.check(lib.crbm_train_step(a, s))
        torch.cuda.synchronize()
        out.append((m.motifs.get_value(), m.bias.get_value(), m.c.get_value(), m.fantasy_h.get_value(), m.fantasy_h_prime.get_value()))
    for x, y in zip(*out):
        np.testing.assert_array_equal(x, y)
    with pytest.raises(ValueError):
        _lib.check(lib.crbm_train_phases(a, 0, _lib.stream_handle()))


def test_peer_memory_exchange_loopback_equals_plain_step(data, monkeypatch):
    """The data-parallel kernels (publishing reduction + polling, rank-ordered sum inside the update, csrc/train.cu) with this
    rank as its only peer (CRBM_P2P_LOOPBACK=1) must reproduce the single-rank step bit for bit — graph replay, eager packed
    steps and the host-fed one-hot call alike. (The multi-rank check is tests/test_gpu_multi.py.)"""
    import crbm_b200
    from crbm_b200 import CRBM
    seqs = crbm_b200.PackedSequences.from_onehot(data[:20 * 11 + 5])
    out = []
    for loopback in ('1', ''):
        monkeypatch.setenv('CRBM_P2P_LOOPBACK', loopback)
        np.random.seed(4)
        m = CRBM(num_motifs=14, motif_length=11, epochs=2, cd_k=2, seed=21)
        assert m._dp == (loopback == '1')
        m._train_full_batches(seqs, 0, 11)        # (not fit(): its evaluation pass advances the Philox step of the plain path only)
        for i in range(3):
            m._trainingFct(data[20 * i:20 * i + 20])
        m._train_packed(seqs, 0, 20)
        out.append((m.motifs.get_value(), m.bias.get_value(), m.c.get_value(), m.fantasy_h.get_value()))
        if loopback:
            m._p2p_check()
    for a, b in zip(*out):
        np.testing.assert_array_equal(a, b)


def test_unsupported_shapes_fail_loudly(data):
    from crbm_b200 import CRBM
    m = CRBM(num_motifs=2, motif_length=34)                 # beyond the tuned kernels: served by the shape-generic ones
    assert m.motifHitProbs(data[:2]).shape == (2, 2, 1, 167)
    m = CRBM(num_motifs=2, motif_length=5, pooling=3)       # 196 % 3 != 0: the reference's reshape fails too
    with pytest.raises(ValueError):
        m.motifHitProbs(data[:2])
    m = CRBM(num_motifs=2, motif_length=5, pooling=7)       # 7 does not divide the 200-long persistent chain
    with pytest.raises(ValueError):
        m.fit(data[:20])


def test_statistics_vector_is_additive_across_shards(data):
    """Data-parallel contract on the real kernels, emulating two ranks on one GPU: summing the statistics vectors that
    crbm_train_accumulate produces for two shards (half the minibatch + half the fantasy particles each) and applying the
    update gives the parameters of one step on the whole minibatch (same injected uniforms)."""
    import torch
    import crbm_b200
    from crbm_b200 import CRBM, _lib
    lib = _lib.load()
    K, M, B, N = 40, 15, 8, 10
    rs = np.random.RandomState(0)
    W = (0.3 * rs.randn(K, 1, 4, M)).astype(np.float32)
    b = (-3 + 0.3 * rs.randn(1, K)).astype(np.float32)
    fh = (rs.rand(B, K, 1, 200) < 0.03).astype(np.float32)
    fhp = (rs.rand(B, K, 1, 200) < 0.03).astype(np.float32)
    lv = 199 + M
    u_v, u_h, u_hp = rs.rand(B, lv).astype(np.float32), rs.rand(B, K * 200).astype(np.float32), rs.rand(B, K * 200).astype(np.float32)
    d = data[:N]

    def make(bs, fsl):
        m = CRBM(K, M, batchsize=bs, cd_k=1, seed=1)
        m.motifs.set_value(W); m.bias.set_value(b)
        m.fantasy_h.set_value(fh[fsl]); m.fantasy_h_prime.set_value(fhp[fsl])
        return m
    full = make(B, slice(0, B))
    full._trainingFct(d, uniforms={'v': u_v.ravel(), 'h': u_h.ravel(), 'hp': u_hp.ravel()})

    shards, stats = [], []
    for r in range(2):
        fsl, dsl = slice(r * B // 2, (r + 1) * B // 2), slice(r * N // 2, (r + 1) * N // 2)
        m = make(B // 2, fsl)
        seqs = crbm_b200.PackedSequences.from_onehot(d[dsl])
        dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
        uv, uh, uhp = dev(u_v[fsl].ravel()), dev(u_h[fsl].ravel()), dev(u_hp[fsl].ravel())
        a = m._fill_args(seqs.words.data_ptr(), seqs.row_words, seqs.N, seqs.L, uv, uh, uhp)
        _lib.check(lib.crbm_train_accumulate(a, _lib.stream_handle()), 'accumulate')
        torch.cuda.synchronize()
        shards.append((m, a, seqs, (uv, uh, uhp)))
        stats.append(m._train_state['stats'].clone())
    total = stats[0] + stats[1]                     # what the NCCL all-reduce computes
    for m, a, _, _ in shards:
        m._train_state['stats'].copy_(total)
        _lib.check(lib.crbm_train_apply(a, _lib.stream_handle()), 'apply')
        np.testing.assert_allclose(m.motifs.get_value(), full.motifs.get_value(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(m.bias.get_value(), full.bias.get_value(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(m.c.get_value(), full.c.get_value(), rtol=1e-5, atol=2e-6)
    got = np.concatenate([shards[0][0].fantasy_h.get_value(), shards[1][0].fantasy_h.get_value()])
    np.testing.assert_array_equal(got, full.fantasy_h.get_value())


@pytest.mark.parametrize('K,M,L,ds,P', [(100, 15, 200, True, 1), (33, 9, 77, False, 1), (16, 25, 500, True, 1),
                                        (128, 20, 999, True, 5), (24, 12, 83, True, 6), (20, 8, 71, False, 4)])
def test_fused_scan_reductions_match_dense_scan(K, M, L, ds, P):
    """max / argmax / mean over positions computed inside the scan kernel == the same reductions of the
    dense motifHitProbs tensor (argmax exact wherever the maximum is unique at fp32 resolution, numpy first-maximum semantics)."""
    import crbm_b200
    rs = np.random.RandomState(K + L)
    m = crbm_b200.CRBM(K, M, doublestranded=ds, pooling=P, seed=1)
    m.motifs.set_value((0.4 * rs.randn(K, 1, 4, M)).astype(np.float32))
    m.bias.set_value((-3 + rs.randn(1, K)).astype(np.float32))
    seqs = crbm_b200.PackedSequences.random(37, L, seed=3)
    dense = m.motifHitProbs(seqs)[:, :, 0, :]
    got = m.motifHitSummary(seqs)
    # the dense scan runs on tcgen05 (fp16 hi/lo split of W), the fused reductions on mma.sync (bf16 x3 split): the two agree
    # to fp32 rounding, so the maxima are compared at 2e-6 and the hit position wherever the maximum is not a near-tie
    np.testing.assert_allclose(got['max'], dense.max(axis=2), rtol=2e-6, atol=2e-7)
    srt = np.sort(dense, axis=2)
    clear = (srt[:, :, -1] - srt[:, :, -2]) > 4e-6 * np.maximum(srt[:, :, -1], 1e-3)
    assert clear.mean() > 0.95
    np.testing.assert_array_equal(got['argmax'][clear], dense.argmax(axis=2)[clear])
    at = np.take_along_axis(dense, got['argmax'][:, :, None].astype(np.int64), axis=2)[:, :, 0]
    np.testing.assert_allclose(at, dense.max(axis=2), rtol=4e-6, atol=4e-7)     # near-ties: the reported position is a maximum
    np.testing.assert_allclose(got['mean'], dense.mean(axis=2), rtol=1e-5, atol=1e-7)
    # maximum and position alone (no mean: the un-pooled kernel then works on the activities only): the same values bit for bit
    part = m.motifHitSummary(seqs, want=('max', 'argmax'))
    assert sorted(part) == ['argmax', 'max']
    np.testing.assert_array_equal(part['max'], got['max'])
    np.testing.assert_array_equal(part['argmax'], got['argmax'])
    only = m.motifHitSummary(seqs, want=('mean',))
    np.testing.assert_array_equal(only['mean'], got['mean'])
    with pytest.raises(ValueError):
        m.motifHitSummary(seqs, want=('median',))


def test_streamlined_onehot_step_equals_packed_step_and_validates(data):
    """`_trainingFct(one-hot array)` (staging copy + graph replay) gives the same parameters as the explicit packed path,
    and rejects non-one-hot input (reported at the next synchronisation point)."""
    import crbm_b200
    from crbm_b200 import CRBM
    out = []
    for fast in (True, False):
        np.random.seed(3)
        m = CRBM(num_motifs=12, motif_length=9, cd_k=1, seed=5, batchsize=20)
        for i in range(5):
            batch = data[20 * i:20 * (i + 1)]
            if fast:
                m._trainingFct(batch)
            else:
                m._train_packed(crbm_b200.PackedSequences.from_onehot(batch), 0, 20)
        out.append((m.motifs.get_value(), m.bias.get_value(), m.c.get_value(), m.fantasy_h.get_value()))
    for a, b in zip(*out):
        np.testing.assert_array_equal(a, b)
    m = CRBM(num_motifs=4, motif_length=5, seed=1, batchsize=20)
    bad = data[:20].copy()
    bad[3, 0, :, 7] = 0.0
    m._trainingFct(data[:20])
    m._trainingFct(bad)
    with pytest.raises(ValueError):
        m.motifs.get_value()


@pytest.mark.parametrize('kind', ['numpy', 'pinned', 'cuda', 'strided', 'copy', 'zerocopy', 'zerocopy-numpy'])
@pytest.mark.parametrize('L', [200, 197])
def test_hostfed_step_reads_the_batch_where_it_lies(data, kind, L, monkeypatch):
    """The host-fed step: by default (crbm_train_step_staged) the batch is copied by the copy engine into a ring of device staging
    buffers ahead of the step, whose packing kernel waits for the copy's flag; device-resident input and CRBM_HOSTFED=zerocopy
    (crbm_train_step_hostfed) are read by the packing kernel where they lie (device memory / page-locked host memory over PCIe);
    pageable / strided input goes through a ring of pinned buffers; lengths that are not a multiple of four floats take the scalar
    loads. Same parameters as the explicit packed path, bit for bit. `copy` is the first version (CRBM_HOSTFED=copy)."""
    import torch
    import crbm_b200
    from crbm_b200 import CRBM
    if kind == 'copy':
        monkeypatch.setenv('CRBM_HOSTFED', 'copy')
    if kind.startswith('zerocopy'):
        monkeypatch.setenv('CRBM_HOSTFED', 'zerocopy')
    d = np.ascontiguousarray(data[:240, :, :, :L])
    out = []
    for fast in (True, False):
        np.random.seed(3)
        m = CRBM(num_motifs=12, motif_length=9, cd_k=1, seed=5, batchsize=20)
        keep = []
        for i in range(12):                       # more steps than the ring of staging buffers has slots
            batch = d[20 * i:20 * (i + 1)]
            if not fast:
                m._train_packed(crbm_b200.PackedSequences.from_onehot(batch), 0, 20)
                continue
            if kind in ('pinned', 'zerocopy'):
                src = torch.from_numpy(batch.copy()).pin_memory()
            elif kind == 'cuda':
                src = torch.from_numpy(batch.copy()).cuda()
            elif kind == 'strided':
                src = torch.from_numpy(np.repeat(batch, 2, axis=3))[:, :, :, ::2]
                assert not src.is_contiguous()
            else:
                src = batch
            keep.append(src)                  # pinned / device input must stay alive and unmodified until the step has run
            m._trainingFct(src)
        out.append((m.motifs.get_value(), m.bias.get_value(), m.c.get_value(), m.fantasy_h.get_value()))
    for a, b in zip(*out):
        np.testing.assert_array_equal(a, b)


def test_visible_bias_trace_in_pinned_host_memory(data):
    """crbm_train_args.c_trace: every update kernel (host-fed eager step, packed eager step, graph replay) stores the step's c into
    slot step % slots of a page-locked host array."""
    import torch
    import crbm_b200
    from crbm_b200 import CRBM
    np.random.seed(3)
    m = CRBM(num_motifs=12, motif_length=9, cd_k=1, seed=5, batchsize=20)
    trace = m.traceVisibleBias(8)
    seen = []
    for i in range(3):
        m._trainingFct(data[20 * i:20 * (i + 1)])
        seen.append(m.c.get_value().ravel().copy())
    seqs = crbm_b200.PackedSequences.from_onehot(data[:200])
    m._train_packed(seqs, 60, 80)
    seen.append(m.c.get_value().ravel().copy())
    torch.cuda.synchronize()
    for i in range(4):
        np.testing.assert_array_equal(trace[i], seen[i])
    m._train_full_batches(seqs, 0, 10)                  # steps 4..13 by graph replay: slots 6, 7, 0..5 are rewritten
    torch.cuda.synchronize()
    np.testing.assert_array_equal(trace[13 % 8], m.c.get_value().ravel())
    assert len({tuple(r) for r in trace}) == 8
    assert m.traceVisibleBias(0) is None


def test_streamed_fasta_ingest_matches_host_packing(tmp_path):
    """PackedSequences.from_fasta streams blocks through pinned staging buffers and packs on the device: same rows as packing the
    record reader's codes, for block sizes that split the file many times."""
    import crbm_b200
    from crbm_b200.sequences import readSeqsFromFasta, seqToCodes, sample_fasta_path
    p = sample_fasta_path()
    codes = seqToCodes(readSeqsFromFasta(p))
    want = crbm_b200.PackedSequences.from_codes(codes).words.cpu().numpy()
    for cb in (32 << 20, 40000, 3000):
        got = crbm_b200.PackedSequences.from_fasta(p, chunk_bytes=cb)
        assert (got.N, got.L) == codes.shape
        np.testing.assert_array_equal(got.words.cpu().numpy(), want)


def test_empty_input_gives_empty_results(data):
    """No sequence in: arrays of the reference's shapes with a zero first axis out (nothing is launched)."""
    from crbm_b200 import CRBM
    m = CRBM(num_motifs=5, motif_length=7, epochs=1, seed=1)
    D = data[:0]
    assert m.motifHitProbs(D).shape == (0, 5, 1, 194)
    assert m.freeEnergy(D).shape == (0,) and m.freeEnergy(D, permotif=True).shape == (0, 5)
    s = m.motifHitSummary(D)
    assert s['max'].shape == (0, 5) and s['argmax'].shape == (0, 5) and s['mean'].shape == (0, 5)
    assert m.motifHitProbs(data[:3]).shape == (3, 5, 1, 194)
