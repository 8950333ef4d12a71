"""`CRBM` — the secomo.CRBM interface (reference secomo/convRBM.py:25-726) driving the sm_100a kernels.

Same constructor hyper-parameters, defaults, validation errors and method names as the reference class, so code
written against `secomo.CRBM` runs unchanged; the Theano graph is replaced by calls through the C ABI of
libcrbm_b200.so (include/crbm_b200.h). PyTorch only provides device memory, streams and torch.distributed.
There is no CPU fallback: anything that computes raises if no CUDA device / library is available.

Differences a caller can observe (all documented in DESIGN.md):
  * random numbers come from Philox, not MRG31k3p (same distributions; `seed=` keyword added, default clock like :155);
  * exp() is evaluated with the group maximum subtracted, so results stay finite where the reference overflows;
  * any motif_length / pooling (beyond 33 / 16 on slower shape-generic kernels); input_dims must be 4 to compute;
  * `fit`/`motifHitProbs`/`freeEnergy` also accept a GPU-resident `PackedSequences` (2 bit/base);
  * under torch.distributed (one process per GPU) `fit` is data-parallel: every minibatch and the persistent chain
    are sharded across ranks and one all-reduce of the statistics vector per step keeps all replicas identical.
"""
import ctypes
import os
import time
import warnings

import numpy as np

from . import _lib
from .sequences import PackedSequences

FANTASY_LEN = _lib.FANTASY_LEN


def _torch():
    import torch
    return torch


def torch_int32():
    return _torch().int32


class _SharedParam(object):
    """Stand-in for a `theano.shared` variable (reference :133,149,152): get_value() / set_value() in host layout."""

    def __init__(self, owner, name, shape):
        self._owner, self.name, self._shape = owner, name, tuple(shape)

    def _tensor(self):
        return getattr(self._owner, '_t_' + self.name)

    def get_value(self, borrow=False):
        self._owner._check_deferred()
        return self._tensor().detach().cpu().numpy().reshape(self._shape).copy()

    def set_value(self, value, borrow=False):
        torch = _torch()
        value = np.asarray(value, dtype=np.float32)
        if value.shape != self._shape:
            raise ValueError("%s: expected shape %r, got %r" % (self.name, self._shape, value.shape))
        self._tensor().copy_(torch.from_numpy(np.ascontiguousarray(value)).reshape(self._tensor().shape))


class _FantasyParam(object):
    """The persistent chain as the reference exposes it: dense (batchsize,K,1,200) float (reference :168-173)."""

    def __init__(self, owner, strand):
        self._owner, self._strand = owner, strand

    def get_value(self, borrow=False):
        o = self._owner
        o._require_cuda()
        torch = _torch()
        lib = _lib.load()
        dense = torch.empty((o._local_batch, o.num_motifs, 1, FANTASY_LEN), dtype=torch.float32, device=o._device)
        _lib.check(lib.crbm_chain_to_dense(_lib.ptr(o._chain()[self._strand]), o._local_batch, o.num_motifs, o.motif_length,
                                           o.pooling, _lib.ptr(dense), _lib.stream_handle()), 'chain_to_dense')
        return dense.cpu().numpy()

    def set_value(self, value, borrow=False):
        o = self._owner
        o._require_cuda()
        torch = _torch()
        lib = _lib.load()
        value = np.ascontiguousarray(value, dtype=np.float32)
        if value.shape != (o._local_batch, o.num_motifs, 1, FANTASY_LEN):
            raise ValueError("fantasy particles: expected %r" % ((o._local_batch, o.num_motifs, 1, FANTASY_LEN),))
        dense = torch.from_numpy(value).to(o._device)
        _lib.check(lib.crbm_chain_from_dense(_lib.ptr(dense), o._local_batch, o.num_motifs, o.motif_length, o.pooling,
                                             _lib.ptr(o._chain()[self._strand]), _lib.stream_handle()), 'chain_from_dense')


class CRBM(object):
    """Convolutional restricted Boltzmann machine for DNA motifs (reference secomo/convRBM.py:25).

    Parameters are those of the reference (num_motifs, motif_length, epochs=100, input_dims=4, doublestranded=True,
    batchsize=20, learning_rate=0.1, momentum=0.95, pooling=1, cd_k=5, rho=0.01, lambda_rate=0.1), plus keyword-only
    `seed` (Philox seed; default: clock, as reference :155) and `device`.
    """

    def __init__(self, num_motifs, motif_length, epochs=100, input_dims=4,
                 doublestranded=True, batchsize=20, learning_rate=0.1,
                 momentum=0.95, pooling=1, cd_k=5,
                 rho=0.01, lambda_rate=0.1, *, seed=None, device=None):
        # sanity checks: same conditions and messages as reference :73-108
        if num_motifs <= 0:
            raise Exception("Number of motifs must be positive.")
        if motif_length <= 0:
            raise Exception("Motif length must be positive.")
        if epochs < 0:
            raise Exception("Epochs must be non-negative.")
        if input_dims <= 0:
            raise Exception("input_dims must be positive.")
        elif input_dims != 4:
            warnings.warn("input_dims != 4 was not comprehensively \
                tested yet. Be careful when interpreting the results.",
                          UserWarning)
        if batchsize <= 0:
            raise Exception("batchsize must be positive.")
        if learning_rate <= 0.0:
            raise Exception("learning_rate must be positive.")
        if not (momentum >= 0.0 and momentum < 1.):
            raise Exception("momentum must be between zero and one.")
        if pooling <= 0:
            raise Exception("pooling must be positive.")
        if cd_k <= 0:
            raise Exception("cd_k must be positive.")
        if not (rho >= 0.0 and rho < 1.):
            raise Exception("rho must be between zero and one.")
        if lambda_rate < 0.:
            raise Exception("lambda_rate must be non-negative.")

        self.num_motifs = num_motifs
        self.motif_length = motif_length
        self.input_dims = input_dims
        self.doublestranded = doublestranded
        self.batchsize = batchsize
        self.learning_rate = learning_rate
        self.momentum = momentum
        self.rho = rho
        self.lambda_rate = lambda_rate
        self.pooling = pooling
        self.cd_k = cd_k
        self.epochs = epochs
        self.spmethod = 'entropy'

        torch = _torch()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        self._device = torch.device(device)

        # data-parallel layout: `batchsize` is the global minibatch; each rank owns batchsize/world particles
        self._rank, self._world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._rank, self._world = dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        if batchsize % self._world != 0:
            raise Exception("batchsize must be divisible by the number of data-parallel ranks.")
        self._local_batch = batchsize // self._world
        self._p2p_loopback = self._world == 1 and os.environ.get('CRBM_P2P_LOOPBACK', '') == '1'
        self._dp = self._world > 1 or self._p2p_loopback           # steps go through the data-parallel exchange

        x = np.random.randn(num_motifs, 1, input_dims, motif_length).astype(np.float32)        # reference :127
        if not rho:                                                                             # reference :136-140
            rho = 1. / (num_motifs * motif_length)
            if doublestranded:
                rho = rho / 2.
            self.rho = rho
        from scipy.stats import norm
        b = (np.zeros((1, num_motifs)) + norm.ppf(self.rho, 0, np.sqrt(motif_length))).astype(np.float32)   # :143-148
        c = np.zeros((1, input_dims), dtype=np.float32)

        dev = self._device
        self._t_W = torch.from_numpy(x).to(dev).contiguous()
        self._t_bias = torch.from_numpy(b).to(dev).contiguous()
        self._t_c = torch.from_numpy(c).to(dev).contiguous()
        self._t_vW = torch.zeros_like(self._t_W)
        self._t_vb = torch.zeros_like(self._t_bias)
        self._t_vc = torch.zeros_like(self._t_c)
        self.motifs = _SharedParam(self, 'W', x.shape)
        self.bias = _SharedParam(self, 'bias', b.shape)
        self.c = _SharedParam(self, 'c', c.shape)
        self.motif_velocity = _SharedParam(self, 'vW', x.shape)
        self.bias_velocity = _SharedParam(self, 'vb', b.shape)
        self.c_velocity = _SharedParam(self, 'vc', c.shape)

        self._seed = int(time.time()) if seed is None else int(seed)                           # reference :155
        self._step = 0
        self._kw = (num_motifs + 31) // 32
        self._t_fantasy = None      # the persistent chain (bit planes), allocated with the first compute call (_chain)
        self.fantasy_h = _FantasyParam(self, 0)
        if doublestranded:
            self.fantasy_h_prime = _FantasyParam(self, 1)
        self._train_state = None
        self._engine = None
        self._use_graphs = True
        self._p2p = None            # peer-memory all-reduce state (data parallel), set up lazily
        self._dp_seq = 0            # number of data-parallel training steps done (identical on every rank)
        self._sync_replicas()                             # data parallel: every rank starts from rank 0's draw of W
        self.theano_getHitProbs = self.motifHitProbs      # the attribute the north star calls getHitProbs (:508)
        self.getHitProbs = self.motifHitProbs
        self.theano_freeEnergy = lambda data: self.freeEnergy(data)
        self.theano_fePerMotif = lambda data: self.freeEnergy(data, permotif=True)
        self.theano_evaluateData = self._evaluateData
        self.theano_evaluateParams = self._evaluateParams
        self.theano_trainingFct = self._trainingFct

    # ------------------------------------------------------------------------------------------------------
    def _sync_replicas(self):
        """Data parallel: parameters and velocities of every rank := rank 0's. The reference draws W from the unseeded NumPy
        generator (:127), so the replicas would otherwise start — and, since every rank applies the same summed update, stay —
        different. Called by the constructor and at the start of fit() (after any set_value / loadModel of the caller)."""
        if self._world <= 1:
            return
        import torch.distributed as dist
        for t in (self._t_W, self._t_bias, self._t_c, self._t_vW, self._t_vb, self._t_vc):
            dist.broadcast(t, 0)

    def _require_cuda(self):
        if self._device.type != 'cuda':
            raise RuntimeError("crbm_b200 computes only on a CUDA device (sm_100a); there is no CPU fallback")
        if self.input_dims != 4:
            raise NotImplementedError("crbm_b200 kernels are written for the DNA alphabet (input_dims == 4)")
        return _lib.load()

    def _chain(self):
        """The persistent chain (reference :168-173) as bit planes [2][local_batch][words]; the word layout follows the kernels
        that train this model shape (include/crbm_b200.h: crbm_chain_layout)."""
        if self._t_fantasy is None:
            lib = self._require_cuda()
            words = int(lib.crbm_chain_words(self.num_motifs, self.motif_length, self.pooling))
            self._t_fantasy = _torch().zeros((2, self._local_batch, words), dtype=torch_int32(), device=self._device)
        return self._t_fantasy

    def _as_packed(self, data):
        self._require_cuda()
        self._check_deferred()
        if isinstance(data, PackedSequences):
            return data
        data = np.asarray(data) if not hasattr(data, 'shape') else data
        return PackedSequences.from_onehot(data, device=self._device)

    # ---- persistence (reference :177-236): identical pickle layout ------------------------------------------------
    def saveModel(self, filename):
        """Save the model parameters and additional hyper-parameters."""
        import joblib
        numpyParams = (self.motifs.get_value(), self.bias.get_value(), self.c.get_value())
        hyperparams = (self.num_motifs, self.motif_length, self.input_dims, self.doublestranded, self.batchsize,
                       self.learning_rate, self.momentum, self.rho, self.lambda_rate, self.pooling, self.cd_k,
                       self.epochs, self.spmethod)
        joblib.dump((numpyParams, hyperparams), filename, protocol=2)

    @classmethod
    def loadModel(cls, filename):
        """Load a model from a given pickle file."""
        import joblib
        numpyParams, hyperparams = joblib.load(filename)
        (num_motifs, motif_length, input_dims, doublestranded, batchsize, learning_rate, momentum, rho, lambda_rate,
         pooling, cd_k, epochs, spmethod) = hyperparams
        obj = cls(num_motifs, motif_length, epochs=epochs, input_dims=input_dims, doublestranded=doublestranded,
                  batchsize=batchsize, learning_rate=learning_rate, momentum=momentum, pooling=pooling, cd_k=cd_k,
                  rho=rho, lambda_rate=lambda_rate)
        motifs, bias, c = numpyParams
        obj.motifs.set_value(motifs)
        obj.bias.set_value(bias)
        obj.c.set_value(c)
        return obj

    # ---- bottom-up builders (reference :238-275), eager ------------------------------------------------------------
    def _v2h_device(self, seqs, strand_mode, out_kind, start=0, end=None, buf=None):
        """(end-start, K, 1, Lh) device tensor of activities / probabilities of rows [start, end) of `seqs`.
        It is a view of a buffer whose rows are padded to a multiple of 8 floats: the tcgen05 scan kernel then writes
        whole 32-byte sectors only; the values are those of the contiguous layout."""
        lib = self._require_cuda()
        torch = _torch()
        K, M = self.num_motifs, self.motif_length
        Lh = seqs.L - M + 1
        if Lh <= 0:
            raise ValueError("sequences of length %d are shorter than the motif (%d)" % (seqs.L, M))
        end = seqs.N if end is None else end
        n = end - start
        pitch = self._scan_pitch(Lh, out_kind)
        if buf is None:
            buf = torch.empty((n, K, 1, pitch), dtype=torch.float32, device=self._device)
        else:
            buf = buf[:n]
        _lib.check(lib.crbm_v2h_dense_strided(_lib.ptr(seqs.words[start:]), seqs.row_words, n, seqs.L, _lib.ptr(self._t_W),
                                              _lib.ptr(self._t_bias), K, M, self.pooling, strand_mode, out_kind,
                                              _lib.ptr(buf), pitch, _lib.stream_handle()), 'v2h_dense')
        return buf[:, :, :, :Lh]

    def _scan_pitch(self, Lh, out_kind):
        pooled = self.pooling != 1 and out_kind == _lib.OUT_PROBABILITY
        return Lh if (pooled and os.environ.get('CRBM_POOL_KERNEL', '')[:1] == 'c') else (Lh + 7) // 8 * 8

    def _staging(self, name, shape, dtype):
        """Pinned host staging buffers, kept on the model and reused across calls (allocating pinned memory costs about as much as
        copying through it once; a result array in pageable memory is filled from them with a host memcpy)."""
        torch = _torch()
        cache = self.__dict__.setdefault('_pinned', {})
        key = (name, tuple(shape), dtype)
        if key not in cache:
            for k in [k for k in cache if k[0] == name]:
                del cache[k]
            cache[key] = torch.empty(shape, dtype=dtype, pin_memory=True)
        return cache[key]

    def _v2h(self, data, strand_mode, out_kind, max_chunk_bytes=128 << 20):
        """Scan `data` and return the (N,K,1,Lh) float32 array on the host. The sequences are scanned in chunks into two device
        buffers; a chunk is copied device -> pinned staging buffer on a second stream while the next chunk is scanned, and from
        there into the result array by the host."""
        seqs = self._as_packed(data)
        K, M = self.num_motifs, self.motif_length
        Lh = seqs.L - M + 1
        if Lh <= 0:
            raise ValueError("sequences of length %d are shorter than the motif (%d)" % (seqs.L, M))
        torch = _torch()
        rows = max(1, min(seqs.N, int(max_chunk_bytes // (4 * K * Lh))))
        nbuf = min(2, -(-seqs.N // rows))
        pitch = self._scan_pitch(Lh, out_kind)
        main = torch.cuda.current_stream(self._device)
        side = torch.cuda.Stream(device=self._device)
        bufs = [torch.empty((rows, K, 1, pitch), dtype=torch.float32, device=self._device) for _ in range(nbuf)]
        # Results up to CRBM_PINNED_RESULT_MAX bytes (default 4 GB) are returned in page-locked memory: the chunks are copied device ->
        # result directly (no staging buffer, no host memcpy into freshly faulted pageable pages, which was the slowest stage: ~5 GB/s),
        # and torch's caching host allocator hands the pages out again once the caller has dropped the previous result.
        nbytes = 4 * seqs.N * K * Lh
        out_t = None
        if nbytes <= int(os.environ.get('CRBM_PINNED_RESULT_MAX', str(4 << 30))):
            try:
                out_t = torch.empty((seqs.N, K, 1, Lh), dtype=torch.float32, pin_memory=True)
            except RuntimeError:
                out_t = None
        if out_t is not None:
            done = [None] * nbuf                                # per device buffer: its last copy has read it
            for i, s in enumerate(range(0, seqs.N, rows)):
                e, j = min(seqs.N, s + rows), i % nbuf
                if done[j] is not None:
                    main.wait_event(done[j])                    # the scan into this buffer follows the copy that read it last
                view = self._v2h_device(seqs, strand_mode, out_kind, s, e, buf=bufs[j])
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    out_t[s:e].copy_(view, non_blocking=True)
                    done[j] = torch.cuda.Event()
                    done[j].record(side)
            side.synchronize()
            main.wait_stream(side)
            return out_t.numpy()
        out = np.empty((seqs.N, K, 1, Lh), dtype=np.float32)
        stage = [self._staging('scan%d' % i, (rows, K, 1, Lh), torch.float32) for i in range(nbuf)]
        copied = [None, None]
        pending = []                                            # (chunk start, end, staging index) whose copies are in flight

        def drain(upto):
            while len(pending) > upto:
                s0, e0, j = pending.pop(0)
                copied[j].synchronize()
                np.copyto(out[s0:e0], stage[j].numpy()[:e0 - s0])

        for i, s in enumerate(range(0, seqs.N, rows)):
            e = min(seqs.N, s + rows)
            drain(nbuf - 1)                                     # the staging buffer this chunk will use has been emptied
            view = self._v2h_device(seqs, strand_mode, out_kind, s, e, buf=bufs[i % nbuf])
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                stage[i % nbuf][:e - s].copy_(view, non_blocking=True)
                copied[i % nbuf] = torch.cuda.Event()
                copied[i % nbuf].record(side)
            main.wait_event(copied[i % nbuf])                   # (orders the next scan into this device buffer after the copy)
            pending.append((s, e, i % nbuf))
        drain(0)
        return out

    def _bottomUpActivity(self, data, flip_motif=False):
        """x = conv(data, W, filter_flip=flip_motif) + b (reference :238-243)."""
        return self._v2h(data, _lib.STRAND_RC if flip_motif else _lib.STRAND_FWD, _lib.OUT_ACTIVITY)

    def _bottomUpProbability(self, activities):
        """exp(x) / sum_group(1 + exp(x)) (reference :245-257)."""
        lib = self._require_cuda()
        torch = _torch()
        act = torch.from_numpy(np.ascontiguousarray(activities, dtype=np.float32)).to(self._device)
        N, K, _, Lh = act.shape
        out = torch.empty_like(act)
        _lib.check(lib.crbm_pool_probability(_lib.ptr(act), N, K, Lh, self.pooling, _lib.ptr(out), _lib.stream_handle()),
                   'pool_probability')
        return out.cpu().numpy()

    def _bottomUpSample(self, probs, uniforms=None):
        """Multinomial sample within every pooling group (reference :259-267)."""
        lib = self._require_cuda()
        torch = _torch()
        p = torch.from_numpy(np.ascontiguousarray(probs, dtype=np.float32)).to(self._device)
        N, K, _, Lh = p.shape
        u = None if uniforms is None else torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float32)).to(self._device)
        out = torch.empty_like(p)
        self._step += 1
        _lib.check(lib.crbm_sample_hidden(_lib.ptr(p), N, K, Lh, self.pooling, _lib.ptr(u), self._seed,
                                          self._step << 32, _lib.ptr(out), _lib.stream_handle()), 'sample_hidden')
        return out.cpu().numpy()

    def _computeHgivenV(self, data, flip_motif=False, uniforms=None):
        """[probability, sample] of the hidden layer (reference :269-275)."""
        prob = self._v2h(data, _lib.STRAND_RC if flip_motif else _lib.STRAND_FWD, _lib.OUT_PROBABILITY)
        return [prob, self._bottomUpSample(prob, uniforms)]

    # ---- top-down builders (reference :277-325), eager --------------------------------------------------------------
    def _topDownActivity(self, h, hprime):
        lib = self._require_cuda()
        torch = _torch()
        hd = torch.from_numpy(np.ascontiguousarray(h, dtype=np.float32)).to(self._device)
        hpd = None if hprime is None else torch.from_numpy(np.ascontiguousarray(hprime, dtype=np.float32)).to(self._device)
        N, K, _, Lh = hd.shape
        if K != self.num_motifs:
            raise ValueError("hidden tensor has %d motif channels, model has %d" % (K, self.num_motifs))
        M = self.motif_length
        out = torch.empty((N, 1, 4, Lh + M - 1), dtype=torch.float32, device=self._device)
        _lib.check(lib.crbm_h2v_dense(_lib.ptr(hd), _lib.ptr(hpd), N, K, Lh, _lib.ptr(self._t_W), _lib.ptr(self._t_c), M,
                                      _lib.ptr(out), _lib.stream_handle()), 'h2v_dense')
        return out.cpu().numpy()

    def _visible(self, activity, want_prob, want_sample, uniforms=None):
        lib = self._require_cuda()
        torch = _torch()
        act = torch.from_numpy(np.ascontiguousarray(activity, dtype=np.float32)).to(self._device)
        N, _, A, Lv = act.shape
        prob = torch.empty_like(act) if want_prob else None
        samp = torch.empty_like(act) if want_sample else None
        u = None if uniforms is None else torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float32)).to(self._device)
        self._step += 1
        _lib.check(lib.crbm_visible_softmax_sample(_lib.ptr(act), N, Lv, _lib.ptr(prob), _lib.ptr(u), self._seed,
                                                   self._step << 32, _lib.ptr(samp), _lib.stream_handle()),
                   'visible_softmax_sample')
        return (prob.cpu().numpy() if want_prob else None), (samp.cpu().numpy() if want_sample else None)

    def _topDownProbability(self, activity, softmaxdown=True):
        if not softmaxdown:
            raise NotImplementedError("softmaxdown=False is dead code in the reference (:299) and is not built")
        return self._visible(activity, True, False)[0]

    def _topDownSample(self, probability, softmaxdown=True, uniforms=None):
        if not softmaxdown:
            raise NotImplementedError("softmaxdown=False is dead code in the reference (:312-314) and is not built")
        # sampling from probabilities == sampling from softmax(log p)
        with np.errstate(divide='ignore'):
            logp = np.log(np.asarray(probability, dtype=np.float64)).astype(np.float32)
        return self._visible(np.maximum(logp, -80.0), False, True, uniforms)[1]

    def _computeVgivenH(self, H_sample, H_sample_prime, softmaxdown=True, uniforms=None):
        """[probability, sample] of the visible layer (reference :317-325)."""
        act = self._topDownActivity(H_sample, H_sample_prime)
        prob, samp = self._visible(act, True, True, uniforms)
        return [prob, samp]

    # ---- scan / free energy (reference :507-568, 657-697) --------------------------------------------------------------
    def motifHitProbs(self, data):
        """Per-position motif match probabilities, (N,K,1,L-M+1) float32.

        As in the reference (:507-514): doublestranded=True scans the forward strand only; doublestranded=False sums
        the activities of both strands before the pooling sigmoid."""
        mode = _lib.STRAND_FWD if self.doublestranded else _lib.STRAND_SUM
        return self._v2h(data, mode, _lib.OUT_PROBABILITY)

    def motifHitSummary(self, data, chunk_rows=None, want=('max', 'argmax', 'mean')):
        """Per sequence and motif: max over positions of motifHitProbs, its position and the mean over positions —
        the reductions secomo.utils applies to motifHitProbs (utils.py:155,245-247,306) — computed inside the scan kernel,
        so the (N,K,1,Lh) tensor is never materialised. Returns dict(max=(N,K) float32, argmax=(N,K) int32, mean=(N,K)).
        `want` selects the results: without 'mean' an un-pooled model is scanned on the activities alone (the maximum of the
        sigmoid is the sigmoid of the maximum activity: one exponential per sequence and motif instead of one per position).
        Long inputs are scanned in chunks of `chunk_rows` sequences (two sets of device buffers; the device->host copy of a
        chunk into pinned staging memory overlaps the scan of the next one)."""
        lib = self._require_cuda()
        torch = _torch()
        seqs = self._as_packed(data)
        K, M = self.num_motifs, self.motif_length
        N = seqs.N
        rows = max(1, min(N, int(chunk_rows or max(1024, (64 << 20) // (4 * K)))))
        nbuf = min(2, -(-N // rows))
        names = tuple(t for t in (('max', torch.float32, np.float32), ('argmax', torch.int32, np.int32), ('mean', torch.float32, np.float32))
                      if t[0] in want)
        if not names or any(w not in ('max', 'argmax', 'mean') for w in want):
            raise ValueError("motifHitSummary: want must be a non-empty subset of ('max', 'argmax', 'mean')")
        bufs = [{nm: torch.empty((rows, K), dtype=td, device=self._device) for nm, td, _ in names} for _ in range(nbuf)]
        mode = _lib.STRAND_FWD if self.doublestranded else _lib.STRAND_SUM
        main = torch.cuda.current_stream(self._device)
        side = torch.cuda.Stream(device=self._device)

        def scan(s, e, b):
            _lib.check(lib.crbm_scan_reduce(_lib.ptr(seqs.words[s:]), seqs.row_words, e - s, seqs.L, _lib.ptr(self._t_W),
                                            _lib.ptr(self._t_bias), K, M, self.pooling, mode, _lib.OUT_PROBABILITY,
                                            _lib.ptr(b.get('max')), _lib.ptr(b.get('argmax')), _lib.ptr(b.get('mean')), _lib.stream_handle()),
                       'scan_reduce')

        # results up to CRBM_PINNED_RESULT_MAX bytes (default 4 GB) are returned in page-locked memory, the chunks copied device ->
        # result directly (see _v2h)
        out_t = None
        if 4 * N * K * len(names) <= int(os.environ.get('CRBM_PINNED_RESULT_MAX', str(4 << 30))):
            try:
                out_t = {nm: torch.empty((N, K), dtype=td, pin_memory=True) for nm, td, _ in names}
            except RuntimeError:
                out_t = None
        if out_t is not None:
            done = [None] * nbuf
            for i, s in enumerate(range(0, N, rows)):
                e, j = min(N, s + rows), i % nbuf
                if done[j] is not None:
                    main.wait_event(done[j])                    # the scan into these buffers follows the copies that read them last
                scan(s, e, bufs[j])
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    for nm, _, _ in names:
                        out_t[nm][s:e].copy_(bufs[j][nm][:e - s], non_blocking=True)
                    done[j] = torch.cuda.Event()
                    done[j].record(side)
            side.synchronize()
            main.wait_stream(side)
            return {nm: t.numpy() for nm, t in out_t.items()}
        out = {nm: np.empty((N, K), dtype=npd) for nm, _, npd in names}
        stage = [{nm: self._staging('sum_%s%d' % (nm, i), (rows, K), td) for nm, td, _ in names} for i in range(nbuf)]
        copied = [None, None]
        pending = []

        def drain(upto):
            while len(pending) > upto:
                s0, e0, j = pending.pop(0)
                copied[j].synchronize()
                for nm, _, _ in names:
                    np.copyto(out[nm][s0:e0], stage[j][nm].numpy()[:e0 - s0])

        for i, s in enumerate(range(0, N, rows)):
            e = min(N, s + rows)
            j = i % nbuf
            drain(nbuf - 1)
            b = bufs[j]
            scan(s, e, b)
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                for nm, _, _ in names:
                    stage[j][nm][:e - s].copy_(b[nm][:e - s], non_blocking=True)
                copied[j] = torch.cuda.Event()
                copied[j].record(side)
            main.wait_event(copied[j])
            pending.append((s, e, j))
        drain(0)
        return out

    def motifHitPositionMean(self, data, max_chunk_bytes=256 << 20):
        """Mean hit probability per (motif, position) over the sequences, (K, L-M+1) float32: what
        `secomo.utils.positionalDensityPlot` draws (utils.py:113-116, `motifHitProbs(seqs).mean(axis=(0, 2))`). The sequences are
        scanned in chunks into one reused device buffer and summed over the sequence axis on the device (crbm_position_sum), so
        neither the host nor the device ever holds the (N,K,1,Lh) tensor."""
        lib = self._require_cuda()
        torch = _torch()
        seqs = self._as_packed(data)
        K, M = self.num_motifs, self.motif_length
        Lh = seqs.L - M + 1
        if Lh <= 0:
            raise ValueError("sequences of length %d are shorter than the motif (%d)" % (seqs.L, M))
        mode = _lib.STRAND_FWD if self.doublestranded else _lib.STRAND_SUM
        pitch = self._scan_pitch(Lh, _lib.OUT_PROBABILITY)
        rows = max(1, min(seqs.N, int(max_chunk_bytes // (4 * K * pitch))))
        buf = torch.empty((rows, K, 1, pitch), dtype=torch.float32, device=self._device)
        acc = torch.zeros((K, Lh), dtype=torch.float32, device=self._device)
        for s in range(0, seqs.N, rows):
            e = min(seqs.N, s + rows)
            self._v2h_device(seqs, mode, _lib.OUT_PROBABILITY, s, e, buf=buf)
            _lib.check(lib.crbm_position_sum(_lib.ptr(buf), e - s, K, Lh, pitch, _lib.ptr(acc), _lib.stream_handle()), 'position_sum')
        return (acc / float(seqs.N)).cpu().numpy()

    def _free_energy_device(self, seqs, want_seq, want_motif, want_count, uniforms=None):
        lib = self._require_cuda()
        torch = _torch()
        K, M = self.num_motifs, self.motif_length
        fe_seq = torch.empty(seqs.N, dtype=torch.float32, device=self._device) if want_seq else None
        fe_mot = torch.empty((seqs.N, K), dtype=torch.float32, device=self._device) if want_motif else None
        cnt = torch.empty(seqs.N, dtype=torch.float32, device=self._device) if want_count else None
        if seqs.N == 0:                                           # no sequence: empty results, nothing to launch
            return fe_seq, fe_mot, cnt
        u = None if uniforms is None else torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float32)).to(self._device)
        self._step += 1
        _lib.check(lib.crbm_free_energy(_lib.ptr(seqs.words), seqs.row_words, seqs.N, seqs.L, _lib.ptr(self._t_W),
                                        _lib.ptr(self._t_bias), _lib.ptr(self._t_c), K, M, self.pooling,
                                        1 if self.doublestranded else 0, _lib.ptr(fe_seq), _lib.ptr(fe_mot), _lib.ptr(cnt),
                                        _lib.ptr(u), self._seed, self._step, _lib.stream_handle()), 'free_energy')
        return fe_seq, fe_mot, cnt

    def freeEnergy(self, data, permotif=False):
        """Free energy per sequence (N,), or per sequence and motif (N,K) (reference :549-568)."""
        seqs = self._as_packed(data)
        fe_seq, fe_mot, _ = self._free_energy_device(seqs, not permotif, permotif, False)
        return (fe_mot if permotif else fe_seq).cpu().numpy()

    # ---- training (reference :373-438, 570-634) ---------------------------------------------------------------------------
    def _ensure_train_state(self):
        if self._train_state is None:
            lib = self._require_cuda()
            torch = _torch()
            K, M = self.num_motifs, self.motif_length
            lv = FANTASY_LEN + M - 1
            st = {}
            st['v_row_words'] = lib.crbm_packed_row_words(lv)
            st['v_model'] = torch.zeros((self._local_batch, st['v_row_words']), dtype=torch.int32, device=self._device)
            st['stats'] = torch.zeros(int(lib.crbm_train_stats_floats(K, M)), dtype=torch.float32, device=self._device)
            st['ws_bytes'] = int(lib.crbm_train_workspace_bytes_for(K, M, self._local_batch))
            st['workspace'] = torch.empty(st['ws_bytes'], dtype=torch.uint8, device=self._device)
            lay = _lib.StatsLayout()
            _lib.check(lib.crbm_train_stats_layout(K, M, lay), 'stats_layout')
            st['layout'] = lay
            self._train_state = st
        return self._train_state

    def _fill_args(self, words_ptr, row_words, n_data, L, u_v=None, u_h=None, u_hp=None):
        st = self._ensure_train_state()
        a = _lib.TrainArgs()
        a.W, a.b, a.c = self._t_W.data_ptr(), self._t_bias.data_ptr(), self._t_c.data_ptr()
        a.vW, a.vb, a.vc = self._t_vW.data_ptr(), self._t_vb.data_ptr(), self._t_vc.data_ptr()
        a.K, a.M, a.pooling = self.num_motifs, self.motif_length, self.pooling
        a.doublestranded, a.cd_k = (1 if self.doublestranded else 0), self.cd_k
        a.learning_rate, a.momentum, a.rho, a.lambda_rate = self.learning_rate, self.momentum, self.rho, self.lambda_rate
        a.data, a.data_row_words, a.n_data, a.L = words_ptr, row_words, n_data, L
        a.batchsize = self._local_batch
        a.fantasy_bits, a.v_model, a.v_row_words = self._chain().data_ptr(), st['v_model'].data_ptr(), st['v_row_words']
        a.seed, a.step = self._seed + 7919 * self._rank, self._step
        a.u_v = None if u_v is None else u_v.data_ptr()
        a.u_h = None if u_h is None else u_h.data_ptr()
        a.u_hp = None if u_hp is None else u_hp.data_ptr()
        a.step_dev, a.batch_index_dev, a.n_batches = None, None, 0
        a.stats, a.workspace, a.workspace_bytes = st['stats'].data_ptr(), st['workspace'].data_ptr(), st['ws_bytes']
        a.c_trace, a.c_trace_slots = self._trace_args()
        return a

    # ---- data parallel: one-shot all-reduce over NVLink peer memory, fused into the update --------------------------------
    def _p2p_setup(self):
        """Expose this rank's double-buffered statistics vector + flags to the peers (torch symmetric memory = CUDA VMM /
        IPC handles exchanged through the process group). Returns None (-> NCCL all-reduce) if that is not possible."""
        if self._p2p is not None:
            return self._p2p or None
        self._p2p = False
        torch = _torch()
        lib = _lib.load()
        if self._world == 1 and self._p2p_loopback:
            # single-rank loop-back (CRBM_P2P_LOOPBACK=1): the peer-memory kernels run with this rank as its only peer, in
            # ordinary device memory — exercises the publish / poll / rank-ordered sum path on a one-GPU box
            nfl = int(lib.crbm_train_stats_floats(self.num_motifs, self.motif_length))
            stride = (nfl + 63) // 64 * 64
            buf = torch.zeros(2 * stride + 64, dtype=torch.float32, device=self._device)
            self._ensure_train_state()
            self._p2p = dict(buf=buf, hdl=None, stride=stride, peer_stats=[buf.data_ptr()], peer_flags=[buf.data_ptr() + 2 * stride * 4],
                             ticket=torch.zeros(1, dtype=torch.int32, device=self._device),
                             error=torch.zeros(1, dtype=torch.int32, device=self._device), local_stats=buf.data_ptr())
            ll = int(lib.crbm_p2p_ll_elems(self.num_motifs, self.motif_length, self.pooling))
            if ll and os.environ.get('CRBM_DP_LL', '1') != '0':
                self._p2p['ll_buf'] = torch.zeros(2 * ll, dtype=torch.int64, device=self._device)
                self._p2p['ll_stride'], self._p2p['peer_ll'] = ll, [self._p2p['ll_buf'].data_ptr()]
            return self._p2p
        if self._world == 1 or self._world > _lib.MAX_PEERS or os.environ.get('CRBM_DP', 'p2p') != 'p2p':
            return None
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            st = self._ensure_train_state()
            nfl = int(lib.crbm_train_stats_floats(self.num_motifs, self.motif_length))
            stride = (nfl + 63) // 64 * 64
            buf = symm_mem.empty(2 * stride + 64, dtype=torch.float32, device=self._device)
            buf.zero_()
            torch.cuda.synchronize(self._device)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            assert len(ptrs) == self._world and ptrs[self._rank] == buf.data_ptr()
            flag_off = 2 * stride * 4
            p = dict(buf=buf, hdl=hdl, stride=stride, peer_stats=ptrs, peer_flags=[x + flag_off for x in ptrs],
                     ticket=torch.zeros(1, dtype=torch.int32, device=self._device),
                     error=torch.zeros(1, dtype=torch.int32, device=self._device), local_stats=buf.data_ptr())
            # receive buffers of the fused exchange (reduce_update_ll_kernel): 2 parities x crbm_p2p_ll_elems 8-byte elements
            ll = int(lib.crbm_p2p_ll_elems(self.num_motifs, self.motif_length, self.pooling))
            if ll and os.environ.get('CRBM_DP_LL', '1') != '0':
                llbuf = symm_mem.empty(2 * ll, dtype=torch.int64, device=self._device)
                llbuf.zero_()
                torch.cuda.synchronize(self._device)
                llh = symm_mem.rendezvous(llbuf, dist.group.WORLD.group_name)
                p['ll_buf'], p['ll_hdl'], p['ll_stride'] = llbuf, llh, ll
                p['peer_ll'] = [int(x) for x in llh.buffer_ptrs]
            dist.barrier()
            self._p2p = p
            return p
        except Exception as e:                              # pragma: no cover - depends on the machine
            warnings.warn("crbm_b200: peer-memory all-reduce unavailable (%s); falling back to NCCL all_reduce" % (e,))
            return None

    def _p2p_args(self, p, parity, sequence, ll_ok=True):
        a = _lib.P2PArgs()
        a.n_ranks, a.rank, a.parity, a.sequence = self._world, self._rank, parity, sequence
        for r in range(self._world):
            a.peer_stats[r], a.peer_flags[r] = p['peer_stats'][r], p['peer_flags'][r]
        a.stats_stride, a.error_flag, a.ticket = p['stride'], p['error'].data_ptr(), p['ticket'].data_ptr()
        if ll_ok and 'peer_ll' in p:
            # fused exchange inside the reduction + update launch; only when every rank has rows in this step (all ranks must
            # take the same branch)
            for r in range(self._world):
                a.peer_ll[r] = p['peer_ll'][r]
            a.ll_parity_stride = p['ll_stride']
        return a

    def _p2p_step(self, args, p, parity, sequence, ll_ok=True):
        """One data-parallel step on the current stream: statistics into buffer `parity`, exchange over peer memory, update
        (one launch for reduction + exchange + update where crbm_train_step_p2p fuses them)."""
        lib = _lib.load()
        s = _lib.stream_handle()
        args.stats = p['local_stats'] + parity * p['stride'] * 4
        pa = self._p2p_args(p, parity, sequence, ll_ok)
        _lib.check(lib.crbm_train_step_p2p(args, pa, s), 'train_step_p2p')

    def _p2p_check(self, collective=False):
        """Raise if the peer-memory exchange of an earlier step timed out on this rank (the kernel then poisons the
        parameters it could not update with NaN instead of applying a partial sum). Called at every host synchronisation
        point: get_value / saveModel (local flag) and the end of every epoch (`collective`: the flags of all ranks are
        combined first, so that every rank raises, not only the ones that waited)."""
        p = self._p2p
        if not p:
            return
        err = p['error']
        if collective and self._world > 1:
            import torch.distributed as dist
            err = err.clone()
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if int(err.item()) != 0:
            raise RuntimeError("crbm_b200: a data-parallel peer did not publish its statistics in time; the parameters of the "
                               "affected step were not updated and are marked invalid (NaN)")

    def _train_packed(self, seqs, start, end, uniforms=None, local_shard=False):
        """One training step on rows [start,end) of `seqs` (this rank's shard is cut here unless `local_shard`)."""
        lib = self._require_cuda()
        torch = _torch()
        n = end - start
        ll_ok = True
        if self._world > 1 and not local_shard:
            per = (n + self._world - 1) // self._world
            lo = min(n, self._rank * per)
            hi = min(n, lo + per)
            start, end = start + lo, start + hi
            ll_ok = (self._world - 1) * per < n          # no rank's shard of this step is empty
        u_v = u_h = u_hp = None
        if uniforms is not None:
            conv = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self._device)
            u_v, u_h, u_hp = conv(uniforms.get('v')), conv(uniforms.get('h')), conv(uniforms.get('hp'))
        row_bytes = seqs.row_words * 4
        a = self._fill_args(seqs.words.data_ptr() + start * row_bytes, seqs.row_words, end - start, seqs.L, u_v, u_h, u_hp)
        s = _lib.stream_handle()
        if self._dp:
            p = self._p2p_setup()
            if p is not None:
                a.step = self._dp_seq
                self._p2p_step(a, p, self._dp_seq & 1, self._dp_seq + 1, ll_ok and end > start)
            else:
                import torch.distributed as dist
                _lib.check(lib.crbm_train_accumulate(a, s), 'train_accumulate')
                dist.all_reduce(self._train_state['stats'])
                _lib.check(lib.crbm_train_apply(a, s), 'train_apply')
            self._dp_seq += 1
        else:
            _lib.check(lib.crbm_train_step(a, s), 'train_step')
        self._step += 1

    # ---- CUDA-graph replay of whole training steps ------------------------------------------------------------------
    def _local_shard(self, seqs):
        """Rows of the full minibatches this rank trains on, contiguous, batch-major: [n_full][local_batch]."""
        lb = self._local_batch
        if self._world == 1 or getattr(seqs, 'is_local_shard', False):
            n_full = seqs.N // lb
            return seqs.words[:n_full * lb], n_full
        n_full = seqs.N // self.batchsize
        w = seqs.words[:n_full * self.batchsize].view(n_full, self.batchsize, seqs.row_words)
        return w[:, self._rank * lb:(self._rank + 1) * lb].reshape(n_full * lb, seqs.row_words).contiguous(), n_full

    def _graph_engine(self, seqs, steps_per_graph):
        """Capture `steps_per_graph` consecutive training steps (accumulate -> all-reduce -> apply) into one CUDA graph.
        The minibatch cursor and the Philox step live in device memory, so the same graph is replayed for every step."""
        torch = _torch()
        lib = self._require_cuda()
        key = (seqs.N, seqs.L, steps_per_graph)
        eng = getattr(self, '_engine', None)
        # the engine keeps the sequences it was captured on alive and is matched by identity: a data pointer may be handed
        # out again by the caching allocator for a different dataset of the same shape (the shard is a copy under DP)
        if eng is not None and eng['seqs'] is seqs and eng['key'] == key:
            return eng
        shard, n_full = self._local_shard(seqs)
        eng = dict(key=key, seqs=seqs, shard=shard, n_full=n_full, graphs={})
        eng['step_dev'] = torch.zeros(1, dtype=torch.int64, device=self._device)
        eng['batch_dev'] = torch.zeros(1, dtype=torch.int32, device=self._device)
        a = self._fill_args(shard.data_ptr(), seqs.row_words, self._local_batch, seqs.L)
        a.step_dev, a.batch_index_dev, a.n_batches = eng['step_dev'].data_ptr(), eng['batch_dev'].data_ptr(), n_full
        eng['args'] = a
        self._engine = eng
        return eng

    def _run_steps_once(self, eng, part='both'):
        lib = _lib.load()
        s = _lib.stream_handle()
        a = eng['args']
        if part == 'both' and self._world == 1:
            _lib.check(lib.crbm_train_step(a, s), 'train_step')
            return
        if part.startswith('p2p'):
            self._p2p_step(a, self._p2p, int(part[3:]), 0)       # sequence comes from *step_dev inside the kernels
            return
        if part in ('both', 'accumulate'):
            _lib.check(lib.crbm_train_accumulate(a, s), 'train_accumulate')
        if part == 'both' and self._world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._train_state['stats'])
        if part in ('both', 'apply'):
            _lib.check(lib.crbm_train_apply(a, s), 'train_apply')

    def _graph_for(self, eng, nsteps, part='both'):
        torch = _torch()
        key = (nsteps, part)
        g = eng['graphs'].get(key)
        if g is None:
            # warm-up outside capture (first-use cudaFuncSetAttribute), state restored after
            tensors = (self._t_W, self._t_bias, self._t_c, self._t_vW, self._t_vb, self._t_vc,
                       self._chain(), eng['step_dev'], eng['batch_dev'])
            saved = [t.clone() for t in tensors]
            side = torch.cuda.Stream(device=self._device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                if part != 'p2p':                     # (p2p kernels were warmed up by the eager alignment steps)
                    self._run_steps_once(eng, part)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self._device)
            for t, v in zip(tensors, saved):
                t.copy_(v)
            eng['cursor'] = None                      # (the warm-up ran on the cursors and restored them: be explicit)
            g = torch.cuda.CUDAGraph()
            c0 = _lib.load().crbm_launch_count()
            with torch.cuda.graph(g):
                for i in range(nsteps):
                    self._run_steps_once(eng, ('p2p%d' % (i & 1)) if part == 'p2p' else part)
            if part in ('both', 'accumulate', 'p2p'):
                eng['launches_per_step'] = int(_lib.load().crbm_launch_count() - c0) // nsteps + (1 if part == 'accumulate' else 0)
            eng['graphs'][key] = g
        return g

    def _train_full_batches(self, seqs, first_batch, count, steps_per_graph=8):
        """Train on `count` consecutive full minibatches starting at minibatch `first_batch`, by graph replay.
        Single GPU: `steps_per_graph` whole steps per graph. Data parallel: the NCCL all-reduce of the statistics
        vector stays outside the graphs (accumulate graph -> all_reduce -> apply graph), one step at a time."""
        if count <= 0:
            return
        eng = self._graph_engine(seqs, steps_per_graph)
        p = self._p2p_setup() if self._dp else None
        if p is not None:
            # whole steps (accumulate -> publish -> peer all-reduce + update) replayed as graphs of an even number of
            # steps starting at an even sequence number, so that the buffer parity of every graph node is static
            lb, rw = self._local_batch, seqs.row_words
            done = 0

            def eager(nsteps):
                nonlocal done
                for _ in range(nsteps):
                    a = self._fill_args(eng['shard'].data_ptr() + (first_batch + done) * lb * rw * 4, rw, lb, seqs.L)
                    a.step = self._dp_seq
                    self._p2p_step(a, p, self._dp_seq & 1, self._dp_seq + 1)
                    self._dp_seq += 1
                    done += 1
            spg = steps_per_graph - (steps_per_graph & 1)
            eager(min(count, (self._dp_seq & 1) + (0 if ('p2p_warm' in eng) else 2)))
            eng['p2p_warm'] = True
            if spg >= 2 and count - done >= 2:
                # the device cursors are only written when they do not already stand where this call starts (a caller that continues
                # where the previous call ended launches nothing but the graphs)
                if eng.get('cursor') != ('p2p', self._dp_seq, (first_batch + done) % eng['n_full']):
                    eng['step_dev'].fill_(self._dp_seq)
                    eng['batch_dev'].fill_(first_batch + done)
                while count - done >= 2:
                    # graphs of spg steps, then ONE graph of the remaining even number of steps (cached per length); only a last
                    # odd step runs eagerly
                    n = spg if count - done >= spg else (count - done) & ~1
                    g = self._graph_for(eng, n, 'p2p')
                    nrep = (count - done) // n
                    for _ in range(nrep):
                        g.replay()
                    self._dp_seq += nrep * n
                    done += nrep * n
                eng['cursor'] = ('p2p', self._dp_seq, (first_batch + done) % eng['n_full'])
            eager(count - done)
            self._step += count
            return
        if eng.get('cursor') != ('local', self._step, first_batch % eng['n_full']):
            eng['step_dev'].fill_(self._step)
            eng['batch_dev'].fill_(first_batch)
        eng['cursor'] = ('local', self._step + count, (first_batch + count) % eng['n_full'])
        if self._world > 1:
            import torch.distributed as dist
            ga, gb = self._graph_for(eng, 1, 'accumulate'), self._graph_for(eng, 1, 'apply')
            stats = self._train_state['stats']
            for _ in range(count):
                ga.replay()
                dist.all_reduce(stats)
                gb.replay()
        else:
            big, rest = divmod(count, steps_per_graph)
            if big:
                g = self._graph_for(eng, steps_per_graph)
                for _ in range(big):
                    g.replay()
            if rest:
                self._graph_for(eng, rest).replay()      # the remainder as ONE graph of `rest` steps (cached per length)
        self._step += count

    def _train_epoch(self, seqs):
        """All minibatches of one epoch in order (no shuffling, reference :613-615); a short last batch runs eagerly."""
        n_full = seqs.N // self.batchsize
        if self._use_graphs:
            self._train_full_batches(seqs, 0, n_full)
        else:
            for i in range(n_full):
                self._train_packed(seqs, i * self.batchsize, (i + 1) * self.batchsize)
        if seqs.N % self.batchsize:
            self._train_packed(seqs, n_full * self.batchsize, seqs.N)
        self._p2p_check(collective=True)

    def _trainingFct(self, data, uniforms=None, local_shard=False):
        """Train on one mini-batch (reference :524-526 -> theano_trainingFct). Under data parallelism `data` is the
        global minibatch (every rank passes the same array) unless local_shard=True (this rank's rows only).

        One-hot input (NumPy or torch, host or device) takes the streamlined path: one call into the library issues the
        chain sweeps (h2v -> v2h of the persistent particles), copies the batch into a persistent staging buffer and packs it
        on a second stream meanwhile, then positive phase -> reduction -> update. The one-hot validation
        result is checked at the next call / synchronisation point."""
        if (uniforms is None and self._use_graphs and not isinstance(data, PackedSequences) and getattr(data, 'ndim', 0) == 4
                and (not self._dp or self._p2p_setup() is not None)):
            ll_ok = True
            if self._world > 1 and not local_shard:         # this rank's rows of the global minibatch
                n = int(data.shape[0])
                per = (n + self._world - 1) // self._world
                lo = min(n, self._rank * per)
                data = data[lo:min(n, lo + per)]
                ll_ok = (self._world - 1) * per < n
            if int(data.shape[0]) > 0:
                return self._train_step_onehot(data, ll_ok)
        seqs = self._as_packed(data)
        self._train_packed(seqs, 0, seqs.N, uniforms, local_shard)

    def _check_deferred(self):
        """Raise if the one-hot validation of an earlier streamlined training call failed (the kernels keep a running
        count on the device; it is read here, at the synchronisation points of the API, not once per minibatch)."""
        self._p2p_check()
        eng = getattr(self, '_onehot_engine', None)
        if eng is not None and eng.get('pending'):
            eng['pending'] = False
            nbad = int(eng['bad'].item())
            if nbad != 0:
                eng['bad'].zero_()
                raise ValueError("input is not one-hot at %d positions (expected exactly one 1 among A,C,G,T)" % nbad)

    def _train_step_onehot(self, data, ll_ok=True):
        shape = data.shape
        n, L = int(shape[0]), int(shape[3])
        eng = getattr(self, '_onehot_engine', None)
        if eng is None or eng['key'] != (n, L):
            eng = self._onehot_engine_for(data, n, L)
        torch = eng['torch']
        a, pa = eng['args'], None
        if self._dp:
            # data parallel: statistics into buffer `parity` of the peer-mapped double buffer, sequence number per step
            p = self._p2p
            a.step = self._dp_seq
            a.stats = p['local_stats'] + (self._dp_seq & 1) * p['stride'] * 4
            pa = self._p2p_args(p, self._dp_seq & 1, self._dp_seq + 1, ll_ok)
            self._dp_seq += 1
        else:
            a.step = self._step
        a.c_trace, a.c_trace_slots = self._trace_args()
        stream = torch.cuda.current_stream(self._device).cuda_stream
        if eng['hostfed']:
            on_device = (not isinstance(data, np.ndarray)) and data.is_cuda and data.dtype == torch.float32 and data.is_contiguous()
            if eng['staged'] and not on_device:
                # crbm_train_step_staged: the batch crosses PCIe on the copy engine AHEAD of the step (ring of device staging
                # buffers, one per step in flight; the step's packing kernel waits for a flag the copy engine writes). Nothing
                # but kernels on the compute stream; a host that runs ahead of the device hides the transfer.
                i = eng['staged_pos']
                eng['staged_pos'] = i + 1
                R = len(eng['stage'])
                slot = i % R
                if i >= R and i % eng['gate_every'] == 0:
                    # the event recorded after step i - R + 3: the slots of steps i .. i + 3 are free
                    eng['gate_events'][((i - R + 4) // eng['gate_every']) % len(eng['gate_events'])].synchronize()
                if isinstance(data, np.ndarray) or not data.is_pinned() or data.dtype != torch.float32 or not data.is_contiguous():
                    hb = eng['ring'][slot]
                    if isinstance(data, np.ndarray):
                        np.copyto(hb['np'], data, casting='unsafe')
                    else:
                        hb['buf'].copy_(data)
                    src_ptr = hb['buf'].data_ptr()
                else:
                    src_ptr = data.data_ptr()
                eng['seq_np'][slot] = (i + 1) & 0x7fffffff
                rc = eng['fn_staged'](a, pa, src_ptr, eng['stage'][slot].data_ptr(), eng['flags_ptr'] + 4 * slot, eng['seq_ptr'] + 4 * slot,
                                      eng['c_bad'], eng['c_side'], stream)
                if rc != _lib.OK:
                    _lib.check(rc, 'train_step_staged')
                if (i + 1) % eng['gate_every'] == 0:
                    # recorded every few steps: the slot of step i is reused R steps later, after this event has been waited for
                    eng['gate_events'][((i + 1) // eng['gate_every']) % len(eng['gate_events'])].record()
            # One C call, one stream, no copy node (crbm_train_step_hostfed): the packing kernel reads the one-hot array where it
            # lies — device memory, or page-locked host memory over PCIe — while the chain sweeps run.
            elif isinstance(data, np.ndarray) or not (data.is_cuda or data.is_pinned()) or data.dtype != torch.float32 or not data.is_contiguous():
                # pageable / strided / non-fp32 input: one host-side copy into a ring of page-locked buffers (a slot is reused only
                # after the step that read it has executed)
                slot = eng['ring'][eng['ring_pos'] % len(eng['ring'])]
                eng['ring_pos'] += 1
                slot['event'].synchronize()
                src = slot['buf']
                if isinstance(data, np.ndarray):
                    np.copyto(slot['np'], data, casting='unsafe')
                else:
                    src.copy_(data)
                rc = eng['fn_hostfed'](a, pa, src.data_ptr(), eng['c_bad'], stream)
                slot['event'].record()
            else:
                # (a pinned HOST tensor or a device tensor is read by the kernels of this step: it must stay unmodified until the
                # step has executed on the device — after a stream synchronisation, or once a later trace slot has been written)
                rc = eng['fn_hostfed'](a, pa, data.data_ptr(), eng['c_bad'], stream)
            if rc != _lib.OK:
                _lib.check(rc, 'train_step_hostfed')
        else:
            src = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)) if isinstance(data, np.ndarray) else data
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.to(torch.float32).contiguous()
            if src.is_cuda:
                # the library copies `src` on the side stream: order that copy after whatever produced the tensor on the current
                # stream, and keep its memory from being reused before the copy has run
                eng['side'].wait_stream(torch.cuda.current_stream(self._device))
                src.record_stream(eng['side'])
            # CRBM_HOSTFED=copy: the first version of this call (crbm_train_step_onehot): chain sweeps on the current stream; copy of
            # the batch into a staging buffer + packing + validation on a side stream; then positive phase, reduction, update.
            rc = eng['fn'](a, pa, src.data_ptr(), eng['c_staging'], eng['c_bad'], eng['c_side'], eng['c_done'], eng['c_ready'], stream)
            if rc != _lib.OK:
                _lib.check(rc, 'train_step_onehot')
        eng['pending'] = True
        self._step += 1

    def _onehot_engine_for(self, data, n, L):
        """Persistent buffers of the streamlined one-hot training call for batches of (n, L)."""
        lib = self._require_cuda()
        torch = _torch()
        if tuple(data.shape[1:3]) != (1, 4):
            raise ValueError("expected one-hot data of shape (N,1,4,L), got %r" % (tuple(data.shape),))
        self._check_deferred()
        rw = lib.crbm_packed_row_words(L)
        eng = dict(key=(n, L), pending=False, torch=torch, fn=lib.crbm_train_step_onehot, fn_hostfed=lib.crbm_train_step_hostfed, rw=rw)
        mode = os.environ.get('CRBM_HOSTFED', 'staged')
        eng['hostfed'] = mode != 'copy'
        eng['staged'] = mode == 'staged'
        eng['fn_staged'] = lib.crbm_train_step_staged
        eng['words'] = torch.zeros((n, rw), dtype=torch.int32, device=self._device)
        eng['bad'] = torch.zeros(1, dtype=torch.int32, device=self._device)
        eng['args'] = self._fill_args(eng['words'].data_ptr(), rw, n, L)
        eng['c_bad'] = eng['bad'].data_ptr()
        if eng['hostfed']:
            R = 8 if eng['staged'] else 4
            eng['ring'], eng['ring_pos'] = [], 0
            for _ in range(R):
                buf = torch.empty((n, 1, 4, L), dtype=torch.float32).pin_memory()
                ev = torch.cuda.Event()
                ev.record()
                eng['ring'].append(dict(buf=buf, np=buf.numpy(), event=ev))
            if eng['staged']:
                eng['stage'] = [torch.empty((n, 1, 4, L), dtype=torch.float32, device=self._device) for _ in range(R)]
                eng['flags'] = torch.zeros(R, dtype=torch.int32, device=self._device)
                eng['seq'] = torch.zeros(R, dtype=torch.int32).pin_memory()
                eng['seq_np'] = eng['seq'].numpy()
                eng['flags_ptr'], eng['seq_ptr'] = eng['flags'].data_ptr(), eng['seq'].data_ptr()
                eng['staged_pos'], eng['gate_every'] = 0, 4
                eng['gate_events'] = [torch.cuda.Event() for _ in range(2 * R // 4 + 2)]
                eng['side'] = torch.cuda.Stream(device=self._device)
                eng['c_side'] = eng['side'].cuda_stream
                torch.cuda.synchronize(self._device)
        else:
            eng['staging'] = torch.empty((n, 1, 4, L), dtype=torch.float32, device=self._device)
            main = torch.cuda.current_stream(self._device)
            eng['side'] = torch.cuda.Stream(device=self._device)
            eng['ready'], eng['done'] = torch.cuda.Event(), torch.cuda.Event()
            eng['ready'].record(eng['side'])
            eng['done'].record(main)
            eng['c_staging'] = eng['staging'].data_ptr()
            eng['c_side'], eng['c_done'], eng['c_ready'] = eng['side'].cuda_stream, eng['done'].cuda_event, eng['ready'].cuda_event
        self._onehot_engine = eng
        return eng

    # ---- per-step trace of the visible bias in page-locked host memory ------------------------------------------------------
    def traceVisibleBias(self, slots=64):
        """Have every training step also store the updated visible bias c into slot (step % slots) of a page-locked host
        array, written by the update kernel itself over PCIe (crbm_train_args.c_trace): a training loop can log c per step
        without a device->host copy in the stream. Returns the (slots, 4) float32 NumPy view; entries are valid after the next
        device synchronisation. slots=0 switches the trace off."""
        torch = _torch()
        self._require_cuda()
        if slots <= 0:
            self._c_trace = None
            return None
        t = torch.zeros((int(slots), 4), dtype=torch.float32).pin_memory()
        self._c_trace = t
        self._engine = None          # captured graphs hold the old argument block
        return t.numpy()

    def _trace_args(self):
        t = getattr(self, '_c_trace', None)
        return (t.data_ptr(), int(t.shape[0])) if t is not None else (None, 0)

    def _evaluateData(self, data, uniforms=None):
        """[mean free energy, mean of a sampled forward-strand H] (reference :466-472, 517-522)."""
        seqs = self._as_packed(data)
        fe, _, cnt = self._free_energy_device(seqs, True, False, True, uniforms)
        lh = seqs.L - self.motif_length + 1
        return [float(fe.sum().item() / seqs.N), float(cnt.sum().item() / (seqs.N * self.num_motifs * lh))]

    def _evaluateParams(self):
        """[sqrt(mean W^2), IC, median IC] (reference :475-485, 528-533)."""
        W = self.motifs.get_value().astype(np.float64)
        e = np.exp(W)
        pwm = e / e.sum(axis=2, keepdims=True)
        entropy = np.sum(-pwm * np.log2(pwm), axis=2)
        twn = np.sqrt(np.mean(W ** 2))
        ic = np.log2(W.shape[2]) - np.mean(entropy)
        medic = np.log2(W.shape[2]) - np.mean(np.sort(entropy, axis=2)[:, :, entropy.shape[2] // 2])
        return [np.float32(twn), np.float32(ic), np.float32(medic)]

    def _truncated_length(self, L):
        return int((L - self.motif_length + 1) / self.pooling) * self.pooling + self.motif_length - 1

    def _prepare(self, data):
        """fit()'s truncation (reference :586-591) + packing."""
        if isinstance(data, PackedSequences):
            return data.truncated(self._truncated_length(data.L))
        nseq = self._truncated_length(data.shape[3])
        return self._as_packed(data[:, :, :, :nseq])

    def _epoch_metrics(self, test):
        """Unweighted mean over test batches of [mean FE, mean H sample] (reference :616-627). Under data parallelism the
        batches are dealt out in contiguous runs, every rank evaluates its run in one launch, and the two sums of batch means
        are all-reduced (eval metrics sharded + one small all-reduce per epoch)."""
        torch = _torch()
        idx = self._iterateBatchIndices(test.N, self.batchsize)
        nb = len(idx)
        lo, hi = (self._rank * nb) // self._world, ((self._rank + 1) * nb) // self._world
        sums = torch.zeros(2, dtype=torch.float64, device=self._device)
        if hi > lo:
            s0, e0 = idx[lo][0], idx[hi - 1][1]
            part = test.rows(s0, e0)
            fe, _, cnt = self._free_energy_device(part, True, False, True)
            lh = test.L - self.motif_length + 1
            starts = torch.tensor([s - s0 for s, _ in idx[lo:hi]] + [e0 - s0], device=self._device)
            cfe = torch.cat([torch.zeros(1, device=self._device, dtype=torch.float64), fe.double().cumsum(0)])
            ccn = torch.cat([torch.zeros(1, device=self._device, dtype=torch.float64), cnt.double().cumsum(0)])
            sizes = (starts[1:] - starts[:-1]).double()
            sums[0] = ((cfe[starts[1:]] - cfe[starts[:-1]]) / sizes).sum()
            sums[1] = ((ccn[starts[1:]] - ccn[starts[:-1]]) / (sizes * self.num_motifs * lh)).sum()
        if self._world > 1:
            import torch.distributed as dist
            dist.all_reduce(sums)
        meanfe, meannmh = (sums / nb).tolist()
        return meanfe, meannmh

    def fit(self, training_data, test_data=None):
        """Fits the cRBM to the provided training sequences (reference :570-634)."""
        self._require_cuda()
        self._sync_replicas()
        training = self._prepare(training_data)
        test = self._prepare(test_data) if test_data is not None else training
        verbose = self._rank == 0
        if verbose:
            print(("BatchSize: " + str(self.batchsize)))
            print("Start training the model...")
        starttime = time.time()
        for epoch in range(self.epochs):
            self._train_epoch(training)
            meanfe, meannmh = self._epoch_metrics(test)
            [twn_, ic_, medic_] = self._evaluateParams()
            if verbose:
                print(("Epoch {:d}: ".format(epoch) +
                       "FE={:1.3f} ".format(meanfe) +
                       "NumH={:1.4f} ".format(meannmh) +
                       "WNorm={:2.2f} ".format(float(twn_)) +
                       "IC={:1.3f} medIC={:1.3f}".format(float(ic_), float(medic_))))
        _torch().cuda.synchronize(self._device)
        if verbose:
            print(("Training finished after: {:5.2f} seconds!".format(time.time() - starttime)))

    def getPFMs(self):
        """Weight matrices converted to position frequency matrices (reference :640-655)."""
        def softmax_(x):
            x_exp = np.exp(x)
            y = np.zeros(x.shape)
            for i in range(x.shape[1]):
                y[:, i] = x_exp[:, i] / np.sum(x_exp[:, i])
            return y
        return [softmax_(m[0, :, :]) for m in self.motifs.get_value()]

    def __repr__(self):
        st = "Parameters:\n\n"
        st += "Number of motifs: {}\n".format(self.num_motifs)
        st += "Motif length: {}\n".format(self.motif_length)
        st += "\n"
        st += "Hyper-parameters:\n\n"
        st += "input dims: {:d}".format(self.input_dims)
        st += "doublestranded: {}".format(self.doublestranded)
        st += "batchsize: {:d}".format(self.batchsize)
        st += "learning rate: {:1.3f}".format(self.learning_rate)
        st += "momentum: {:1.3f}".format(self.momentum)
        st += "rho: {:1.4f}".format(self.rho)
        st += "lambda: {:1.3f}".format(self.lambda_rate)
        st += "pooling: {:d}".format(self.pooling)
        st += "cd_k: {:d}".format(self.cd_k)
        st += "epochs: {:d}".format(self.epochs)
        return st

    def _iterateBatchIndices(self, totalsize, nbatchsize):
        """Returns indices in batches (reference :722-726)."""
        return [[i, i + nbatchsize] if i + nbatchsize <= totalsize
                else [i, totalsize] for i in range(totalsize)[0::nbatchsize]]
