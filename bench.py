#!/usr/bin/env python
"""Headline benchmark: CD-1 training throughput (nucleotides/s) of the cRBM on BASELINE.json's configs[1]
("synthetic 1M x 200 bp one-hot, 100 motifs x 15 bp, pooling=1, CD-1, batch 100 on 1 B200"), with the other BASELINE
configurations as sub-records of the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4] [--only-headline]

One "step" = one minibatch update (`theano_trainingFct`, reference secomo/convRBM.py:373-438) on `batch` sequences per
GPU. Weak scaling: every rank trains `batch` sequences per step, one exchange of the statistics vector per step.
  value        whole-job data nucleotides/s of the headline workload, packed dataset resident in HBM (CUDA-graph replay)
  e2e          same metric through the reference-facing call `CRBM._trainingFct(one-hot batch in host memory)`
  roofline     the dominant launch of the step timed alone (tensor pipe) + whole-step fractions; see DESIGN.md "Measurement"
  c3, c4       BASELINE configs 3 and 4 (training; K=128/M=20/pooling 5/1 kb and K=512/M=25/CD-5/500 bp), same fields
  c5           BASELINE config 5: the 512 x 25 motif scan, both getHitProbs strand branches, dense (HBM roofline) and with the
               callers' reductions fused over the whole 3e9-nt job (sharded over the ranks)
  scan         getHitProbs at the headline shape (dense, fused reductions, and end to end with host buffers)
  eval         the per-epoch evaluation pass (free energy + sampled hidden mean, convRBM.py:466-472,616-624)
  ingest       FASTA file -> packed sequences resident in HBM (threaded host parser, all host threads and one)
  fit          N = 1: the whole `CRBM.fit(one-hot NumPy array)` call of one epoch (packing, steps, evaluation pass)
  cpu_baseline the reference's train step restated with torch-CPU ops, on the host cores (N = 1 only)
`--impl reference` times that CPU restatement (oracle/crbm_torch_cpu.py; Theano itself is not installable offline) with
the same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1..3]; n_seq is the per-rank synthetic dataset (each minibatch is read once per epoch)
    'c2': dict(name='C2: synthetic 1M x 200 bp, K=100, M=15, pooling=1, CD-1, batch 100/GPU', K=100, M=15, P=1, cd_k=1,
               batch=100, L=200, n_seq=1000000, default_steps=10000, sub_steps=2048),
    'c3': dict(name='C3: synthetic x 1 kb, K=128, M=20, pooling=5, CD-1, batch 128/GPU', K=128, M=20, P=5, cd_k=1,
               batch=128, L=1000, n_seq=262144, default_steps=2048, sub_steps=1024),
    'c4': dict(name='C4: synthetic x 500 bp, K=512, M=25, pooling=1, CD-5, batch 128/GPU', K=512, M=25, P=1, cd_k=5,
               batch=128, L=500, n_seq=65536, default_steps=512, sub_steps=256),
}
C5 = dict(name='C5: getHitProbs scan of a 3e9-nt synthetic genome (3M x 1 kb), K=512, M=25, both strand branches', K=512, M=25, P=1,
          L=1000, total_nt=3.0e9)
METRIC = 'CD-1 train nt/s'


def load_json(name):
    try:
        with open(os.path.join(ROOT, name)) as f:
            return json.load(f)
    except Exception:
        return None


def peaks():
    p = load_json('MEASURED_PEAKS.json')
    if p:
        return float(p['hbm_gbs']), 'MEASURED_PEAKS.json', float(p.get('sm_max_mhz', 1965.0))
    return 6650.0, 'fallback (B200_PROFILING.md)', 1965.0


def tensor_peak():
    p = load_json('MEASURED_PEAKS.json')
    if p:
        return float(p['bf16_tflops']), 'MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)'
    return 1590.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu captures (profiles/r02_traffic.json)."""
    t = load_json(os.path.join('profiles', 'r02_traffic.json')) or {}
    e = t.get(key)
    return (e['dram_bytes_read'] + e['dram_bytes_write']) if e else None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------------------
def run_reference(args, wl, rank, world):
    """CPU arm: the reference's train step restated with torch-CPU ops, all host threads, bounded sample."""
    if rank != 0:
        return
    import torch
    from oracle.crbm_torch_cpu import TorchCpuCRBM, random_onehot
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = TorchCpuCRBM(wl['K'], wl['M'], pooling=wl['P'], cd_k=wl['cd_k'], batchsize=wl['batch'], seed=0)
    Lt = int((wl['L'] - wl['M'] + 1) / wl['P']) * wl['P'] + wl['M'] - 1
    nb = 8
    data = [random_onehot(wl['batch'], Lt, 100 + i) for i in range(nb)]
    steps, warm = args.steps, args.warmup
    # bounded sample: a step is one minibatch update; if K + W full minibatches would not finish within ~150 s of CPU time, every
    # step trains on the first `rows` sequences of its minibatch instead (stated in `sample`)
    m.train_step(data[0])
    t0 = time.perf_counter()
    m.train_step(data[1])
    t_full = time.perf_counter() - t0
    rows = wl['batch']
    budget = float(os.environ.get('CRBM_REFERENCE_BUDGET_S', '150'))
    if (steps + warm) * t_full > budget:
        rows = max(1, int(wl['batch'] * budget / ((steps + warm) * t_full)))
        data = [d[:rows] for d in data]
    for i in range(warm):
        m.train_step(data[i % nb])
    t0 = time.perf_counter()
    for i in range(steps):
        m.train_step(data[i % nb])
    dt = time.perf_counter() - t0
    nts = steps * rows * Lt / dt
    line = {'metric': METRIC, 'value': nts, 'unit': 'nt/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warm,
            'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': wl['name'], 'batch_per_step': rows, 'seq_len': Lt},
            'cpu_baseline': {'value': nts, 'unit': 'nt/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d train steps of %d x %d bp%s (torch-CPU restatement of the Theano graph; Theano not '
                                       'installable offline)' % (steps, rows, Lt, '' if rows == wl['batch'] else
                                                                 ' (first %d of the %d sequences of every minibatch: time bound)' % (rows, wl['batch']))},
            'e2e': {'value': nts, 'unit': 'nt/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(wl, budget_s=12.0):
    import torch
    from oracle.crbm_torch_cpu import TorchCpuCRBM, random_onehot
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = TorchCpuCRBM(wl['K'], wl['M'], pooling=wl['P'], cd_k=wl['cd_k'], batchsize=wl['batch'], seed=0)
    Lt = int((wl['L'] - wl['M'] + 1) / wl['P']) * wl['P'] + wl['M'] - 1
    data = [random_onehot(wl['batch'], Lt, 100 + i) for i in range(4)]
    for i in range(3):
        m.train_step(data[i % 4])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and n < 5000:
        m.train_step(data[n % 4])
        n += 1
    dt = time.perf_counter() - t0
    return {'value': n * wl['batch'] * Lt / dt, 'unit': 'nt/s', 'cores': cores, 'kind': 'port',
            'sample': '%d train steps of %d x %d bp in %.1f s (torch-CPU restatement of the Theano graph)' % (n, wl['batch'], Lt, dt)}


# ---------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------------
class Ctx(object):
    def __init__(self, rank, world, local_rank, dev, dist):
        self.rank, self.world, self.local_rank, self.dev, self.dist = rank, world, local_rank, dev, dist

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def gather(self, value):
        """The value of every rank (list, on every rank), or None on a single GPU."""
        import torch
        if self.dist is None:
            return None
        t = torch.tensor([float(value)], device=self.dev)
        ts = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(ts, t)
        return [float(x.item()) for x in ts]

    def max_over_ranks(self, ms):
        import torch
        if self.dist is None:
            return ms
        t = torch.tensor([ms], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def train_workload(ctx, wl, steps, warmup, spg, sampler=None):
    """Time `steps` training steps of workload `wl` (graph replay, dataset resident in HBM). Returns (record, model, data)."""
    import numpy as np
    import torch
    import crbm_b200
    K, M, P, B, L = wl['K'], wl['M'], wl['P'], wl['batch'], wl['L']
    np.random.seed(0)
    model = crbm_b200.CRBM(K, M, epochs=1, batchsize=B * ctx.world, pooling=P, cd_k=wl['cd_k'], seed=1234 + ctx.rank, device=ctx.dev)
    Lt = model._truncated_length(L)
    n_seq = wl['n_seq']
    data = crbm_b200.PackedSequences.random(n_seq, Lt, seed=1234 + ctx.rank, device=ctx.dev)
    data.is_local_shard = True      # weak scaling: every rank owns its own synthetic dataset
    nb = n_seq // B

    def run_steps(count, first):
        model._train_full_batches(data, first % nb, count, steps_per_graph=spg)

    # priming (untimed, before the W warm-up steps): every CUDA graph the timed region replays is captured here — the spg-step
    # graph, the graph of the remainder steps % spg, and the data-parallel alignment steps
    prime = 2 * spg + 2
    run_steps(prime, 0)
    if steps % spg:
        run_steps(spg + steps % spg, prime)
        prime += spg + steps % spg
    run_steps(warmup, prime)
    # a rehearsal of the timed region's own graph launches (untimed): the W warm-up steps above replay a different (W-step) graph, and
    # the first replay of a graph after another one runs ~10 % slower (20 steps: 47 us per step cold against 43 us rehearsed)
    # ... repeated for ~30 ms of continuous load: after the idle gaps of the set-up (captures, synchronisations) the first
    # milliseconds of work run 5-7 % slow (20 steps: 41.7 us per step behind one rehearsal, 39.7 behind 10, 39.2 behind 40)
    rehearsal = min(steps, 64)
    nreh = max(1, int(os.environ.get('BENCH_REHEARSALS', str(max(1, 800 // rehearsal)))))
    for i in range(nreh):
        run_steps(rehearsal, prime + warmup + i * rehearsal)
    rehearsal *= nreh
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)
    ctx.barrier()
    if os.environ.get('BENCH_PREFLUSH', '1') != '0':
        # L2 flush (a write larger than the 126 MB L2) queued right before the start event: timing hygiene, and the whole GPU is
        # awake when the first timed kernel starts (after the idle gap behind the synchronisation the first ~100 us of a region run
        # slow: 842 us for 20 steps against 20 x 37.4 us marginal)
        flush.fill_(1)
    e0.record()
    run_steps(steps, prime + warmup + rehearsal)        # continues where the rehearsal ended: the timed region is graph launches only
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    if sampler is not None and ctx.world == 1:
        # short runs end before nvidia-smi's first 100 ms sample: keep the load on until a few samples exist
        t_end = time.time() + 1.5
        while len(sampler.rows) < 4 and time.time() < t_end:
            run_steps(spg, 0)
            torch.cuda.synchronize()
    by_rank = ctx.gather(ms / steps)
    ms = ctx.max_over_ranks(ms)
    rec = {'workload': wl['name'], 'value': steps * B * Lt * ctx.world / (ms * 1e-3), 'unit': 'nt/s', 'steps': steps,
           'ms_per_step': ms / steps, 'batch_per_gpu': B, 'seq_len': Lt,
           'gpu_launches': model._engine.get('launches_per_step', 0) * steps}
    rec['untimed_rehearsal_steps'] = rehearsal
    if by_rank is not None:
        rec['ms_per_step_by_rank'] = by_rank
    return rec, model, data


def work_per_step(wl, Lt):
    """gather-adds of one step (one add per (position, motif, motif base) of every contraction) and the dense-equivalent
    flops of the same contractions (2 flops per one-hot input, 4 inputs per motif base)."""
    K, M, B, cd_k = wl['K'], wl['M'], wl['batch'], wl['cd_k']
    Lh = Lt - M + 1
    pos_data = B * Lh                     # x, x', P(x)V, P'(x)V, Q(x)V
    pos_chain = B * 200                   # per sweep x, x'; last sweep also P(x)V, P'(x)V
    adds = K * M * (5 * pos_data + (2 * cd_k + 2) * pos_chain)
    return adds, 8.0 * adds


def phase_times(model, data, B, dev):
    """Device time of the phases of one training step, each captured 20x back to back in a CUDA graph (no host launch latency) and
    timed ALONE with CUDA events on the launching stream: chain sweeps (h2v + v2h, x cd_k), positive phase (one launch; two for
    motif_length > 16 on the mma.sync path), reduction + update, and the whole step."""
    import torch
    from crbm_b200 import _lib
    lib = _lib.load()
    reps, rounds = 20, 5
    full = model._fill_args(data.words.data_ptr(), data.row_words, B, data.L)

    def timed(fn):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rounds):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * rounds)

    s = _lib.stream_handle
    ph = lambda mask: (lambda: _lib.check(lib.crbm_train_phases(full, mask, s()), 'train_phases'))
    state = (model._t_W, model._t_bias, model._t_c, model._t_vW, model._t_vb, model._t_vc, model._chain())
    saved = [t.clone() for t in state]
    out = {}
    out['whole_step'] = timed(lambda: _lib.check(lib.crbm_train_step(full, s()), 'train_step'))
    for t, v in zip(state, saved):
        t.copy_(v)
    out['chain_sweeps'] = timed(ph(_lib.PHASE_NEGATIVE))
    out['positive_phase'] = timed(ph(_lib.PHASE_POSITIVE))
    out['reduce_update'] = timed(ph(_lib.PHASE_REDUCE | _lib.PHASE_APPLY))
    for t, v in zip(state, saved):              # the per-phase timing trains this rank alone: keep the replicas identical
        t.copy_(v)
    return out


def train_roofline(wl, Lt, rec, kern, key):
    K, M, P, B = wl['K'], wl['M'], wl['P'], wl['batch']
    Lh = Lt - M + 1
    from crbm_b200 import _lib
    tc = bool(_lib.load().crbm_chain_layout(K, M, P))
    pos_us = kern['positive_phase']
    # positive-phase launch: per hidden position and motif the one-hot contraction over 4 M inputs, 2 flops each, for x and x'
    # (up-pass) and for the P, P', Q statistics
    alg_flops = 2.0 * B * Lh * K * 4 * M * 5
    tpeak, tsrc = tensor_peak()
    hbm_peak, _, sm_max = peaks()
    tf = alg_flops / (pos_us * 1e-6) / 1e12 if pos_us else None
    adds, dense = work_per_step(wl, Lt)
    step_s = rec['ms_per_step'] * 1e-3
    name = ('v2h_train_tc_kernel<STAT_P|STAT_Q, P=%d> (tcgen05.mma, statistics accumulated in TMEM; positive-phase launch)' % P) if tc else \
        ('v2h_mma_kernel<KS=%d, %s%s> (mma.sync; positive-phase launch%s)' % ((M + 3) // 4, 'STAT_P|STAT_Q' if M <= 16 else 'STAT_P> + <STAT_Q',
                                                                             ', POOL' if P != 1 else '', '' if M <= 16 else 'es'))
    return {'bound': 'tensor', 'kernel': name, 'achieved': tf, 'peak': tpeak, 'unit': 'TFLOP/s', 'frac': (tf / tpeak) if tf else None,
            'traffic': ncu_traffic(key), 'peak_source': tsrc, 'launch_us': pos_us, 'algorithmic_flops_per_launch': alg_flops,
            'share_of_step': pos_us / kern['whole_step'] if kern.get('whole_step') else None,
            'note': 'dense-equivalent flops of the one-hot contractions (x, x\', and the P, P\', Q statistics); the fp16 hi/lo operand '
                    'splits and tile padding are overhead, not counted',
            'whole_step': {'gather_adds_per_step': adds, 'fp32_add_frac': adds / step_s / (148 * 128 * sm_max * 1e6),
                           'dense_equivalent_tflops': dense / step_s / 1e12, 'tensor_frac': dense / step_s / 1e12 / tpeak,
                           'us_per_step': rec['ms_per_step'] * 1e3,
                           'note': 'SURVEY 8d work of the whole step (positive phase, cd_k chain sweeps, sparsity statistic) over the '
                                   'measured step time; fp32_add_frac is against 148 SM x 128 lanes x %d MHz' % int(sm_max)}}


def timed_calls(ctx, fn, reps=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return ctx.max_over_ranks(e0.elapsed_time(e1) / reps)


def bench_scan(ctx, model, K, M, Lt, traffic_key):
    """getHitProbs kernel (tcgen05 / TMEM, csrc/scan_tc.cu) with fp32 output materialised in HBM; input and output larger
    than L2. Timed in the layout CRBM.motifHitProbs uses on the device (rows padded to a multiple of 8 floats, so every store
    covers whole 32-byte sectors) and in the reference's contiguous (N,K,1,Lh) layout. Algorithmic bytes count Lh floats
    per row in both cases."""
    import torch
    import crbm_b200
    from crbm_b200 import _lib
    lib = _lib.load()
    hbm_peak, peak_src, _ = peaks()
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    Lh = Lt - M + 1
    n = max(1024, min(200000, int(8e9 // (4 * K * Lh))))
    seqs = crbm_b200.PackedSequences.random(n, Lt, seed=99 + rank, device=dev)      # this rank's shard of the sequences
    s = _lib.stream_handle()
    out_rec = {}
    for mode_name, mode in (('forward', _lib.STRAND_FWD), ('both_strands_summed', _lib.STRAND_SUM)):
        ms_by_pitch = {}
        pitch_fast = (Lh + 7) // 8 * 8
        for pitch in sorted({pitch_fast, Lh}, reverse=True):
            if mode != _lib.STRAND_FWD and pitch != pitch_fast:
                continue
            out = torch.empty((n, K, 1, pitch), dtype=torch.float32, device=dev)
            ms_by_pitch[pitch] = timed_calls(ctx, lambda: _lib.check(lib.crbm_v2h_dense_strided(
                _lib.ptr(seqs.words), seqs.row_words, n, Lt, _lib.ptr(model._t_W), _lib.ptr(model._t_bias), K, M, model.pooling,
                mode, _lib.OUT_PROBABILITY, _lib.ptr(out), pitch, s), 'scan'))
            del out
        ms = ms_by_pitch[pitch_fast]
        nt = n * Lt * world                          # whole job
        alg = n * (Lt / 4.0) + 4.0 * n * K * Lh      # per GPU (the roofline is per device)
        rec = {'metric': 'motif-scan nt/s (getHitProbs, fp32 probabilities written)', 'value': nt / (ms * 1e-3), 'unit': 'nt/s',
               'sequences_per_gpu': n, 'n_gpus': world, 'ms': ms, 'output_GB_per_gpu': 4.0 * n * K * Lh / 1e9,
               'roofline': {'bound': 'hbm', 'kernel': lib.crbm_v2h_dense_kernel_name(K, M, model.pooling, _lib.OUT_PROBABILITY, pitch_fast).decode(),
                            'achieved': alg / (ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': alg / (ms * 1e-3) / 1e9 / hbm_peak,
                            'traffic': ncu_traffic(traffic_key) if (mode == _lib.STRAND_FWD and traffic_key) else None,
                            'algorithmic_bytes': alg, 'peak_source': peak_src, 'bytes_per_nt': alg / (n * Lt)}}
        if Lh in ms_by_pitch and Lh != pitch_fast:
            rec['layout'] = ('output rows padded from %d to %d floats (32-byte sectors); contiguous (N,K,1,Lh) layout: %.3f ms = %.3e nt/s'
                             % (Lh, pitch_fast, ms_by_pitch[Lh], nt / (ms_by_pitch[Lh] * 1e-3)))
        out_rec[mode_name] = rec
    # fused reductions (max / argmax / mean over positions per sequence and motif): nothing but (N,K) results is written
    mx = torch.empty((n, K), dtype=torch.float32, device=dev)
    am = torch.empty((n, K), dtype=torch.int32, device=dev)
    mean = torch.empty((n, K), dtype=torch.float32, device=dev)
    for mode_name, mode in (('forward', _lib.STRAND_FWD), ('both_strands_summed', _lib.STRAND_SUM)):
        ms_red = timed_calls(ctx, lambda: _lib.check(lib.crbm_scan_reduce(
            _lib.ptr(seqs.words), seqs.row_words, n, Lt, _lib.ptr(model._t_W), _lib.ptr(model._t_bias), K, M, model.pooling,
            mode, _lib.OUT_PROBABILITY, _lib.ptr(mx), _lib.ptr(am), _lib.ptr(mean), s), 'scan_reduce'), reps=3, warm=1)
        out_rec[mode_name]['fused_reductions'] = {
            'metric': 'motif-scan nt/s, max/argmax/mean over positions fused (no dense output)',
            'value': n * Lt * world / (ms_red * 1e-3), 'unit': 'nt/s', 'ms': ms_red}
        # maximum and its position alone (motifHitSummary(want=('max', 'argmax')); utils.py:155,245-247 use nothing else)
        ms_mx = timed_calls(ctx, lambda: _lib.check(lib.crbm_scan_reduce(
            _lib.ptr(seqs.words), seqs.row_words, n, Lt, _lib.ptr(model._t_W), _lib.ptr(model._t_bias), K, M, model.pooling,
            mode, _lib.OUT_PROBABILITY, _lib.ptr(mx), _lib.ptr(am), None, s), 'scan_reduce'), reps=3, warm=1)
        out_rec[mode_name]['fused_reductions']['max_argmax_only'] = {'value': n * Lt * world / (ms_mx * 1e-3), 'unit': 'nt/s', 'ms': ms_mx}
    out_rec['sharding'] = 'sequences sharded across ranks, weights replicated, no collective on the data path'
    return out_rec


def bench_scan_job(ctx, model, K, M, Lt, total_nt):
    """The whole C5 job through the API: `CRBM.motifHitSummary(packed sequences resident in HBM)` over total_nt / world
    nucleotides per rank, both strand branches (the branch is the model's `doublestranded` flag, convRBM.py:507-514)."""
    import torch
    import crbm_b200
    # every rank scans the shard it would own on 8 GPUs (375 000 sequences per GPU): the whole 3e9-nt job at N = 8,
    # N/8 of it at smaller N (weak scaling; `total_nt` states what was scanned)
    n = int(total_nt / 8 / Lt)
    seqs = crbm_b200.PackedSequences.random(n, Lt, seed=7 + ctx.rank, device=ctx.dev)
    warm = seqs.rows(0, min(n, 40000))          # (more than one chunk: the call's pinned staging buffers take their final shape)
    rec = {'metric': 'motif-scan nt/s, CRBM.motifHitSummary over the whole job (max/argmax/mean per sequence and motif)',
           'total_nt': float(n) * Lt * ctx.world, 'sequences_per_gpu': n, 'unit': 'nt/s'}
    # the first call of a process pins the pages of its result arrays (returned in page-locked memory so that the chunks are copied
    # device -> result directly); torch's host allocator hands them out again afterwards: timed once, reported apart
    t0 = time.time()
    first = model.motifHitSummary(seqs)
    torch.cuda.synchronize()
    rec['first_call_seconds'] = ctx.max_over_ranks((time.time() - t0) * 1e3) * 1e-3
    del first
    for name, ds, want in (('forward', True, ('max', 'argmax', 'mean')), ('both_strands_summed', False, ('max', 'argmax', 'mean')),
                           ('forward_max_argmax_only', True, ('max', 'argmax'))):
        model.doublestranded = ds
        model.motifHitSummary(warm, want=want)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = model.motifHitSummary(seqs, want=want)          # includes the device->host copy of the (N,K) results
        e1.record()
        torch.cuda.synchronize()
        ms = ctx.max_over_ranks(e0.elapsed_time(e1))
        rec[name] = {'value': float(n) * Lt * ctx.world / (ms * 1e-3), 'seconds': ms * 1e-3,
                     'd2h_bytes_per_gpu': int(sum(v.nbytes for v in res.values()))}
        del res
    model.doublestranded = True
    return rec


def bench_eval(ctx, model, K, M, Lt):
    """The per-epoch evaluation pass of fit() (convRBM.py:616-624 -> theano_evaluateData :466-472): mean free energy and the mean of
    a sampled forward-strand hidden layer over a test set, through CRBM._epoch_metrics (batches dealt out over the ranks)."""
    import crbm_b200
    n = 100000 // max(1, (Lt * K) // 20000) * ctx.world
    n = max(model.batchsize * ctx.world, n // model.batchsize * model.batchsize)
    test = crbm_b200.PackedSequences.random(n, Lt, seed=11, device=ctx.dev)       # the same test set on every rank
    ms = timed_calls(ctx, lambda: model._epoch_metrics(test), reps=3, warm=1)
    return {'metric': 'evaluation pass nt/s (mean free energy + sampled hidden mean over the test set)', 'value': n * Lt / (ms * 1e-3),
            'unit': 'nt/s', 'sequences': n, 'ms': ms, 'n_gpus': ctx.world}


def bench_ingest(ctx, Lt):
    """Streaming ingest: a synthetic FASTA file -> PackedSequences on the device (memory-mapped file, blocks parsed
    and packed to 2 bits per base by the host threads of crbm_fasta_pack_block_host into page-locked staging buffers, one
    asynchronous copy of 0.25 byte per base per block while the next block is parsed). Wall clock around the call, best of 3."""
    import tempfile
    import numpy as np
    import torch
    import crbm_b200
    n = 200000
    rs = np.random.RandomState(5 + ctx.rank)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, size=(n, Lt))]
    path = os.path.join(tempfile.gettempdir(), 'crbm_b200_bench_%d_%d.fa' % (os.getpid(), ctx.rank))
    rows = np.empty((n, Lt + 10), dtype=np.uint8)
    rows[:, :9] = np.frombuffer(b'>s0000000', dtype=np.uint8)
    rows[:, 9] = 10
    rows[:, 10:10 + Lt - 1 + 1] = letters
    with open(path, 'wb') as f:
        f.write(np.concatenate([rows, np.full((n, 1), 10, dtype=np.uint8)], axis=1).tobytes())

    def timed(**kw):
        best = None
        for _ in range(4):                                  # the first call pins the staging buffers
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            seqs = crbm_b200.PackedSequences.from_fasta(path, **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert seqs.N == n and seqs.L == Lt
            best = dt if best is None else min(best, dt)
            del seqs
        return best
    try:
        dt = timed()
        dt1 = timed(threads=1)
    finally:
        os.remove(path)
    ms = ctx.max_over_ranks(dt * 1e3)
    ms1 = ctx.max_over_ranks(dt1 * 1e3)
    return {'metric': 'FASTA -> packed sequences resident in HBM, nt/s (host parsing included)', 'value': n * Lt * ctx.world / (ms * 1e-3),
            'unit': 'nt/s', 'records_per_gpu': n, 'file_MB': (n * (Lt + 11)) / 1e6, 'ms': ms, 'host_threads': os.cpu_count(),
            'one_host_thread': {'value': n * Lt * ctx.world / (ms1 * 1e-3), 'ms': ms1}}


def bench_e2e_train(ctx, model, B, Lt, steps):
    """Through the reference-facing call with HOST buffers, every step: `CRBM._trainingFct(one-hot fp32 minibatch in page-locked host
    memory)` — the call enqueues the host -> device copy of the batch (16 B per nucleotide) on the copy engine into a ring of staging
    buffers and the step's kernels on the compute stream, whose packing kernel waits for the copy's flag (crbm_train_step_staged) — and
    the updated visible bias c (16 bytes) comes back device -> host every step, written by the update kernel into a ring of page-locked
    slots (CRBM.traceVisibleBias).
    `value`: the host consumes the ring after a stream synchronisation every `ring` steps, so it keeps issuing steps while the
    device works — the way a training loop that logs a metric runs.
    `blocking_read_value`: the host waits for every step's result before it issues the next step."""
    import numpy as np
    import torch
    nb, ring = 16, 64
    rs = np.random.RandomState(7)
    host = []
    for i in range(nb):
        codes = torch.from_numpy(rs.randint(0, 4, size=(B, Lt)))
        oh = torch.nn.functional.one_hot(codes, 4).permute(0, 2, 1).reshape(B, 1, 4, Lt).to(torch.float32).contiguous()
        host.append(oh.pin_memory())
    trace = model.traceVisibleBias(ring)
    sink = np.zeros(4)
    stream = torch.cuda.current_stream()

    def one(i, blocking):
        slot = model._step % ring
        model._trainingFct(host[i % nb], local_shard=True)          # pinned host one-hot batch in, read by the step itself
        if blocking:
            stream.synchronize()
            sink[:] += trace[slot]                                  # this step's result, already in host memory
        elif i % ring == ring - 1:
            stream.synchronize()                                    # the ring is consumed: its slots may be overwritten
            sink[:] += trace.sum(axis=0)

    def run(blocking, n):
        for i in range(5):
            one(i, True)
        if not blocking:
            for i in range(8 * ring):          # ~20 ms of untimed steps of the same kind (see train_workload: the first milliseconds run slow)
                one(i, False)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            one(i, blocking)
        e1.record()
        torch.cuda.synchronize()
        ms = ctx.max_over_ranks(e0.elapsed_time(e1))
        return n * B * Lt * ctx.world / (ms * 1e-3)

    steps = max(ring, steps // ring * ring)
    blocking = run(True, min(steps, 1024))
    value = run(False, steps)
    model.traceVisibleBias(0)
    assert np.all(np.isfinite(sink))
    return {'value': value, 'unit': 'nt/s', 'h2d_bytes_per_step': B * 4 * Lt * 4, 'd2h_bytes_per_step': 16, 'steps': steps,
            'blocking_read_value': blocking,
            'call': 'CRBM._trainingFct(one-hot fp32 batch in pinned host memory; copied host -> device by the copy engine inside the call, '
                    'the packing kernel of the step waits for its flag) + c of every step written device -> host by the update kernel into '
                    'a pinned ring (consumed after a stream synchronisation every %d steps; blocking_read_value: the host waits for every '
                    'step before issuing the next)' % ring}


def bench_fit(ctx, wl, n=200000):
    """The whole user call: CRBM(...).fit(one-hot NumPy array), one epoch — packing of the host array, every training step of the
    epoch, the evaluation pass over the (same) data and the parameter metrics, wall clock around the call."""
    import contextlib
    import numpy as np
    import torch
    import crbm_b200
    K, M, P, B, L = wl['K'], wl['M'], wl['P'], wl['batch'], wl['L']
    rs = np.random.RandomState(11)
    D = crbm_b200.codesToOneHot(rs.randint(0, 4, size=(n, L)).astype(np.uint8))
    np.random.seed(0)
    best = None
    model = crbm_b200.CRBM(K, M, epochs=1, batchsize=B, pooling=P, cd_k=wl['cd_k'], seed=99, device=ctx.dev)
    for _ in range(3):                                      # (the first call loads the kernels and pins the staging buffers)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):
            model.fit(D)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        assert np.all(np.isfinite(model.motifs.get_value()))
    del model
    Lt = int((L - M + 1) / P) * P + M - 1
    return {'metric': 'CRBM.fit(one-hot NumPy array), 1 epoch: packing + %d training steps + evaluation pass, nt/s' % (n // B),
            'value': n * Lt / best, 'unit': 'nt/s', 'seconds': best, 'sequences': n, 'host_input_bytes': int(D.nbytes)}


def bench_e2e_scan(ctx, model, K, M, Lt):
    """The scan through the reference-facing calls with HOST buffers: `motifHitProbs(one-hot NumPy array)` returns the (N,K,1,Lh)
    NumPy array (host->device copy of the one-hot input, packing, scan, device->host copy of the probabilities inside the timed
    region); `motifHitSummary` the three (N,K) reductions."""
    import numpy as np
    import torch
    import crbm_b200
    Lh = Lt - M + 1
    n = max(256, min(20000, int(1.5e9 // (4 * K * Lh))))
    rs = np.random.RandomState(3 + ctx.rank)
    codes = rs.randint(0, 4, size=(n, Lt))
    D = np.zeros((n, 1, 4, Lt), dtype=np.float32)
    for a in range(4):
        D[:, 0, a, :] = codes == a
    rec = {}
    for name, fn in (('motifHitProbs', model.motifHitProbs), ('motifHitSummary', model.motifHitSummary)):
        fn(D[:64])
        best, out_bytes = None, 0
        for _ in range(3):
            ctx.barrier()
            t0 = time.perf_counter()
            res = fn(D)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            out_bytes = res.nbytes if hasattr(res, 'nbytes') else sum(v.nbytes for v in res.values())
            del res
        ms = ctx.max_over_ranks(best * 1e3)
        rec[name] = {'value': n * Lt * ctx.world / (ms * 1e-3), 'unit': 'nt/s', 'ms': ms, 'sequences_per_gpu': n,
                     'host_input_bytes': int(D.nbytes), 'h2d_bytes': int(n * 4 * crbm_b200.packedRowWords(Lt)), 'd2h_bytes': int(out_bytes)}
    rec['call'] = ('host one-hot fp32 (N,1,4,L) in, NumPy result out; wall clock around the call (it returns host data), best of 3; the '
                   'one-hot array is packed to 2-bit rows by the host threads (crbm_pack_onehot_host), the rows are what crosses PCIe')
    return rec


def main():
    # exactly one line on stdout: libraries that print there (NCCL's version banner) are sent to stderr
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-scan', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--only-headline', action='store_true', help='skip the c3 / c4 / c5 / eval sub-records')
    ap.add_argument('--steps-per-graph', type=int, default=None, help='training steps per CUDA graph (default 8; a run of <= 64 steps is one graph)')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        args.steps = args.steps if args.steps is not None else 40
        args.warmup = args.warmup if args.warmup is not None else 3
        run_reference(args, wl, rank, world)
        return
    args.steps = args.steps if args.steps is not None else wl['default_steps']
    args.warmup = args.warmup if args.warmup is not None else 64
    args.warmup = max(args.warmup, 3)

    import torch
    import crbm_b200  # noqa: F401
    from crbm_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()
    ctx = Ctx(rank, world, local_rank, dev, dist)
    # a short timed region (the driver's --steps 20) as ONE graph launch: otherwise 2-3 graph launches, each with its own launch
    # latency and an idle gap, are a fifth of a 1 ms measurement
    spg = args.steps_per_graph if args.steps_per_graph else (args.steps if args.steps <= 64 else 8)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    rec, model, data = train_workload(ctx, wl, args.steps, args.warmup, spg, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    K, M, B, Lt = wl['K'], wl['M'], wl['batch'], rec['seq_len']

    line = {'metric': METRIC, 'value': rec['value'], 'unit': 'nt/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': rec['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'batch_per_gpu': B, 'seq_len': Lt, 'global_batch': B * world,
                       'parallelism': 'dp%d' % world, 'steps_per_graph': spg, 'untimed_rehearsal_steps': rec['untimed_rehearsal_steps'],
                       'l2': 'every step reads a new minibatch from a %d-sequence packed dataset (%.0f MB); model state '
                             '(W, persistent chain) is on-chip/L2 resident by design; a 256 MB write flushes L2 right before the start event' % (wl['n_seq'], wl['n_seq'] * data.row_words * 4 / 1e6)},
            'gpu_launches': rec['gpu_launches']}
    if 'ms_per_step_by_rank' in rec:
        line['ms_per_step_by_rank'] = rec['ms_per_step_by_rank']

    # every rank runs every section (the sections synchronise the ranks); rank 0 keeps the records
    if rank == 0:
        kern = phase_times(model, data, B, dev)
        line['step_kernels_us'] = kern
        line['roofline'] = train_roofline(wl, Lt, rec, kern, args.workload + '_positive')
    if not args.no_scan:
        sc = bench_scan(ctx, model, K, M, Lt, args.workload + '_scan')
        if not args.no_e2e:
            sc['e2e'] = bench_e2e_scan(ctx, model, K, M, Lt)
        line['scan'] = sc
    if not args.no_e2e:
        line['e2e'] = bench_e2e_train(ctx, model, B, Lt, min(args.steps, 2000))
    if not args.only_headline:
        line['eval'] = bench_eval(ctx, model, K, M, Lt)
        line['ingest'] = bench_ingest(ctx, Lt)
    del model, data
    torch.cuda.empty_cache()

    if not args.only_headline:
        for key in ('c2', 'c3', 'c4'):
            if key == args.workload:
                continue
            w2 = WORKLOADS[key]
            r2, m2, d2 = train_workload(ctx, w2, w2['sub_steps'], 16, 8)
            # (at N > 1 every rank times its own phases, without the exchange, so that a slow GPU shows up beside the job time)
            k2 = phase_times(m2, d2, w2['batch'], dev) if (rank == 0 or world > 1) else None
            local = ctx.gather(k2['whole_step']) if world > 1 else None
            if rank == 0:
                r2['step_kernels_us'] = k2
                r2['roofline'] = train_roofline(w2, r2['seq_len'], r2, k2, key + '_positive')
                if local is not None:
                    r2['single_rank_step_us_by_rank'] = local
            r2['n_gpus'], r2['scaling'] = world, 'weak'
            if key == 'c3' and world > 1 and w2['batch'] % world == 0:
                # strong scaling: config 3's global minibatch of 128 sequences split over the ranks
                ws = dict(w2, batch=w2['batch'] // world, name=w2['name'].replace('batch 128/GPU', 'global batch 128 split over %d GPUs' % world))
                rs, ms_, ds_ = train_workload(ctx, ws, w2['sub_steps'], 16, 8)
                rs['n_gpus'], rs['scaling'] = world, 'strong'
                rs['note'] = 'same optimisation problem as config 3 on one GPU (global batch 128); value = whole-job nt/s'
                line['c3_strong'] = rs
                del ms_, ds_
                torch.cuda.empty_cache()
            if key == 'c3':
                # the pooled scan of config 3's model (getHitProbs with pooling 5)
                r2['scan'] = bench_scan(ctx, m2, w2['K'], w2['M'], r2['seq_len'], 'c3_scan')
            if key == 'c4':
                # config 5 shares config 4's motif bank: the 512 x 25 scan
                Lt5 = C5['L']
                c5 = {'workload': C5['name']}
                c5['dense'] = bench_scan(ctx, m2, C5['K'], C5['M'], Lt5, 'c5_scan')
                c5['job'] = bench_scan_job(ctx, m2, C5['K'], C5['M'], Lt5, C5['total_nt'])
                line['c5'] = c5
            line[key] = r2
            del m2, d2
            torch.cuda.empty_cache()

    if world == 1 and not args.only_headline and 'c3' in line:
        # strong scaling at a global minibatch of 1024 = 8 GPUs x 128; this is its one-GPU end (the same optimisation
        # problem as the 8-GPU weak-scaling record of config 3: compare ms_per_step). Last GPU section: an extra record must not
        # cost the line.
        try:
            w3 = WORKLOADS['c3']
            wb = dict(w3, batch=1024, name=w3['name'].replace('batch 128/GPU', 'global batch 1024 on one GPU'))
            rb, mb, db = train_workload(ctx, wb, 128, 16, 8)
            line['c3']['global_batch_1024_on_one_gpu'] = {k: rb[k] for k in ('workload', 'value', 'unit', 'steps', 'ms_per_step', 'batch_per_gpu')}
            del mb, db
        except Exception as ex:
            line['c3']['global_batch_1024_on_one_gpu'] = {'error': str(ex)[:300]}

    if world == 1 and not args.only_headline:
        try:
            line['fit'] = bench_fit(ctx, wl)
        except Exception as ex:                               # (an extra record must not cost the line)
            line['fit'] = {'error': str(ex)[:300]}

    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline_sample(wl)
        line['clocks'] = clocks
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
