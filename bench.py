#!/usr/bin/env python
"""Headline benchmark: CD-1 training throughput (nucleotides/s) of the cRBM on BASELINE.json's configs[1]
("synthetic 1M x 200 bp one-hot, 100 motifs x 15 bp, pooling=1, CD-1, batch 100 on 1 B200"), plus motif-scan nt/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]

One "step" = one minibatch update (`theano_trainingFct`, reference secomo/convRBM.py:373-438) on `batch` sequences per
GPU. Weak scaling: every rank trains `batch` sequences per step, one all-reduce of the statistics vector per step.
  value  : whole-job data nucleotides/s with the packed dataset resident in HBM (CUDA-graph replay of steps)
  e2e    : same metric through the reference-facing call `CRBM._trainingFct(one-hot NumPy batch)`: pinned-host ->
           device copy of the fp32 one-hot minibatch, packing, the step, and a device -> host read of the updated
           visible bias every step
  scan   : `motifHitProbs` kernel (getHitProbs) nt/s with fp32 probabilities written to HBM
  roofline / cpu_baseline: see DESIGN.md "Measurement".
`--impl reference` times the CPU restatement of the reference's Theano graph (oracle/crbm_torch_cpu.py; Theano itself is
not installable offline) on the host cores with the same metric/config.
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
               batch=100, L=200, n_seq=1000000, default_steps=10000),
    'c3': dict(name='C3: synthetic x 1 kb, K=128, M=20, pooling=5, CD-1, batch 128/GPU', K=128, M=20, P=5, cd_k=1,
               batch=128, L=1000, n_seq=262144, default_steps=2048),
    'c4': dict(name='C4: synthetic x 500 bp, K=512, M=25, pooling=1, CD-5, batch 128/GPU', K=512, M=25, P=1, cd_k=5,
               batch=128, L=500, n_seq=65536, default_steps=512),
}
METRIC = 'CD-1 train nt/s'
SCAN_TRAFFIC = 8209066664   # dram__bytes_read.sum + dram__bytes_write.sum of the scan launch (profiles/r01_ncu_metrics_h.md), padded rows, 107 526 x 200 bp, K=100


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'MEASURED_PEAKS.json', float(p.get('sm_max_mhz', 1965.0))
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)', 1965.0


def tensor_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['bf16_tflops']), 'MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)'
    except Exception:
        return 1590.0, 'fallback (B200_PROFILING.md)'


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
    ap.add_argument('--steps-per-graph', type=int, default=8)
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

    import numpy as np
    import torch
    import crbm_b200
    from crbm_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    K, M, P, B, L = wl['K'], wl['M'], wl['P'], wl['batch'], wl['L']
    np.random.seed(0)
    model = crbm_b200.CRBM(K, M, epochs=1, batchsize=B * world, pooling=P, cd_k=wl['cd_k'], seed=1234 + rank, device=dev)
    Lt = model._truncated_length(L)
    n_seq = wl['n_seq']
    data = crbm_b200.PackedSequences.random(n_seq, Lt, seed=1234 + rank, device=dev)
    data.is_local_shard = True      # weak scaling: every rank owns its own synthetic dataset

    def run_steps(count, first):
        model._train_full_batches(data, first % (n_seq // B), count, steps_per_graph=args.steps_per_graph)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # priming (untimed, before the W warm-up steps): enough steps for every CUDA graph the timed region replays to exist
    # (the 8-step graph, the single-step graph for the remainder, the data-parallel alignment steps)
    prime = 2 * args.steps_per_graph + 3
    run_steps(prime, 0)
    run_steps(args.warmup, prime)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(args.steps, prime + args.warmup)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if rank == 0 and world == 1:
        # short runs end before nvidia-smi's first 100 ms sample: keep the load on until a few samples exist
        t_end = time.time() + 1.5
        while len(sampler.rows) < 4 and time.time() < t_end:
            run_steps(args.steps_per_graph, 0)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nt_per_step = B * Lt * world
    value = args.steps * nt_per_step / (ms * 1e-3)
    launches_per_step = model._engine.get('launches_per_step', 0)   # counted while capturing the step graph

    line = {'metric': METRIC, 'value': value, 'unit': 'nt/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'batch_per_gpu': B, 'seq_len': Lt, 'global_batch': B * world,
                       'parallelism': 'dp%d' % world, 'steps_per_graph': args.steps_per_graph,
                       'l2': 'every step reads a new minibatch from a %d-sequence packed dataset (%.0f MB); model state '
                             '(W, tables, persistent chain) is on-chip/L2 resident by design' % (n_seq, n_seq * data.row_words * 4 / 1e6)},
            'gpu_launches': launches_per_step * args.steps}

    if rank == 0:
        hbm_peak, peak_src, sm_max = peaks()
        # --- per-kernel device time of one step (CUDA events around eager launches on the launching stream) ------------
        state = (model._t_W, model._t_bias, model._t_c, model._t_vW, model._t_vb, model._t_vc, model._chain())
        saved = [t.clone() for t in state]
        kern = time_step_kernels(model, data, B, dev)
        for t, v in zip(state, saved):              # the per-phase timing trains this rank alone: keep the replicas identical
            t.copy_(v)
        line['step_kernels_us'] = kern
        pos_us = kern['positive_phase_kernel']
        Lh = Lt - M + 1
        # Algorithmic work of the positive-phase launch (DESIGN.md "Measurement"): per hidden position and motif the
        # one-hot contraction over 4*M inputs, 2 flops each, for x and x' (up-pass) and for the P, P', Q statistics.
        alg_flops = 2.0 * B * Lh * K * 4 * M * 5
        alg_bytes = B * data.row_words * 4 + K * 4 * M * 4 + K * 4          # packed minibatch + W + b, read once
        adds = B * Lh * K * M * 5                                           # SURVEY 8d gather-add count of the same work
        tf = alg_flops / (pos_us * 1e-6) / 1e12 if pos_us else None
        tpeak = tensor_peak()
        pooled = P != 1
        line['roofline'] = {
            'bound': 'tensor', 'kernel': ('v2h_mma_kernel<KS=%d, %s%s> (positive-phase launch%s)' % ((M + 3) // 4, 'STAT_P|STAT_Q' if M <= 16 else 'STAT_P> + <STAT_Q', ', POOL' if pooled else '', '' if M <= 16 else 'es')),
            'achieved': tf, 'peak': tpeak[0], 'unit': 'TFLOP/s', 'frac': (tf / tpeak[0]) if tf else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full (profiles/r01_ncu_metrics_h.md)
            'traffic': 62720 if (args.workload == 'c2' and not pooled) else None,
            'peak_source': tpeak[1], 'launch_us': pos_us, 'algorithmic_flops_per_launch': alg_flops,
            'note': 'dense-equivalent flops of the one-hot contractions (x, x\', and the P, P\', Q statistics); the fp16 hi/lo / '
                    'bf16x2 operand splits and tile padding are overhead, not counted. At 100 x 200 bp per step the launch is '
                    'latency/issue bound (ncu: profiles/), not throughput bound.',
            'hbm_frac': (alg_bytes / (pos_us * 1e-6) / 1e9 / hbm_peak) if pos_us else None,
            'gather_adds_per_launch': adds,
            'fp32_add_frac': (adds / (pos_us * 1e-6) / (148 * 128 * sm_max * 1e6)) if pos_us else None}
        if not args.no_scan:
            line['scan'] = bench_scan(model, K, M, Lt, dev, hbm_peak, peak_src, world, rank)
        if not args.no_e2e:
            line['e2e'] = bench_e2e(model, B, Lt, world, dev, min(args.steps, 2000))
        if not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_sample(wl)
        line['clocks'] = clocks
        print(json.dumps(line), flush=True)
    else:
        # the other ranks scan their own shard (no communication on the data path) and feed their own minibatches
        if not args.no_scan:
            bench_scan(model, K, M, Lt, dev, 1.0, '', world, rank)
        if not args.no_e2e:
            bench_e2e(model, B, Lt, world, dev, min(args.steps, 2000))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def time_step_kernels(model, data, B, dev):
    """Device time of the phases of one training step. Each phase is captured 20x back to back in a CUDA graph and the
    graph is timed with CUDA events on the launching stream, so host launch latency is excluded. The positive-phase
    kernel is obtained as (accumulate) - (accumulate with an empty minibatch): the latter launches everything but it."""
    import torch
    from crbm_b200 import _lib
    lib = _lib.load()
    reps, rounds = 20, 5
    full = model._fill_args(data.words.data_ptr(), data.row_words, B, data.L)
    empty = model._fill_args(data.words.data_ptr(), data.row_words, 0, data.L)

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

    out = {}
    out['accumulate'] = timed(lambda: _lib.check(lib.crbm_train_accumulate(full, _lib.stream_handle())))
    out['accumulate_without_positive_phase'] = timed(lambda: _lib.check(lib.crbm_train_accumulate(empty, _lib.stream_handle())))
    out['apply'] = timed(lambda: _lib.check(lib.crbm_train_apply(full, _lib.stream_handle())))
    out['fused_step'] = timed(lambda: _lib.check(lib.crbm_train_step(full, _lib.stream_handle())))
    out['positive_phase_kernel'] = max(0.0, out['accumulate'] - out['accumulate_without_positive_phase'])
    return out


def bench_scan(model, K, M, Lt, dev, hbm_peak, peak_src, world=1, rank=0):
    """getHitProbs kernel (tcgen05 / TMEM, csrc/scan_tc.cu) with fp32 output materialised in HBM; input and output larger
    than L2. Timed in the layout CRBM.motifHitProbs uses on the device (rows padded to a multiple of 8 floats, so every store
    covers whole 32-byte sectors) and in the reference's contiguous (N,K,1,Lh) layout. Algorithmic bytes count Lh floats
    per row in both cases."""
    import torch
    import crbm_b200
    from crbm_b200 import _lib
    lib = _lib.load()
    Lh = Lt - M + 1
    n = max(1024, min(200000, int(8e9 // (4 * K * Lh))))
    seqs = crbm_b200.PackedSequences.random(n, Lt, seed=99 + rank, device=dev)      # this rank's shard of the sequences
    s = _lib.stream_handle()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:                               # whole job = all shards; time = slowest rank
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ms_by_pitch = {}
    pitch_fast = (Lh + 7) // 8 * 8
    for pitch in sorted({pitch_fast, Lh}, reverse=True):
        out = torch.empty((n, K, 1, pitch), dtype=torch.float32, device=dev)
        ms_by_pitch[pitch] = timed(lambda: _lib.check(lib.crbm_v2h_dense_strided(
            _lib.ptr(seqs.words), seqs.row_words, n, Lt, _lib.ptr(model._t_W), _lib.ptr(model._t_bias), K, M, model.pooling,
            _lib.STRAND_FWD, _lib.OUT_PROBABILITY, _lib.ptr(out), pitch, s), 'scan'))
        del out
    pitch = pitch_fast
    ms = ms_by_pitch[pitch]
    nt = n * Lt * world                          # whole job
    alg = n * (Lt / 4.0) + 4.0 * n * K * Lh      # per GPU (the roofline is per device)
    # fused reductions (max / argmax / mean over positions per sequence and motif): nothing but (N,K) results is written
    mx = torch.empty((n, K), dtype=torch.float32, device=dev)
    am = torch.empty((n, K), dtype=torch.int32, device=dev)
    mean = torch.empty((n, K), dtype=torch.float32, device=dev)
    fused = None
    if model.pooling == 1:                      # (the fused reductions of pooled probabilities are not built)
        ms_red = timed(lambda: _lib.check(lib.crbm_scan_reduce(
            _lib.ptr(seqs.words), seqs.row_words, n, Lt, _lib.ptr(model._t_W), _lib.ptr(model._t_bias), K, M, model.pooling,
            _lib.STRAND_FWD, _lib.OUT_PROBABILITY, _lib.ptr(mx), _lib.ptr(am), _lib.ptr(mean), s), 'scan_reduce'))
        fused = {'metric': 'motif-scan nt/s, max/argmax/mean over positions fused (no dense output)',
                 'value': nt / (ms_red * 1e-3), 'unit': 'nt/s', 'ms': ms_red}
    return {'fused_reductions': fused,
            'metric': 'motif-scan nt/s (getHitProbs, fp32 probabilities written)', 'value': nt / (ms * 1e-3), 'unit': 'nt/s',
            'sequences_per_gpu': n, 'n_gpus': world, 'ms': ms, 'output_GB_per_gpu': 4.0 * n * K * Lh / 1e9,
            'layout': 'output rows padded from %d to %d floats (32-byte sectors); contiguous (N,K,1,Lh) layout: %.3f ms = %.3e nt/s'
                      % (Lh, pitch, ms_by_pitch[Lh], nt / (ms_by_pitch[Lh] * 1e-3)),
            'sharding': 'sequences sharded across ranks, weights replicated, no collective on the data path',
            'roofline': {'bound': 'hbm', 'kernel': ('v2h_dense_tc_kernel<PROB> (tcgen05.mma, TMEM accumulators)' if model.pooling == 1 else
                                                 'v2h_dense_tc_kernel<PROB, POOL> (tcgen05.mma, TMEM accumulators, pool groups summed by warp shuffles)'),
                         'achieved': alg / (ms * 1e-3) / 1e9, 'peak': hbm_peak,
                         'unit': 'GB/s', 'frac': alg / (ms * 1e-3) / 1e9 / hbm_peak,
                         # ncu --set full on this launch (107 526 x 200 bp, K=100, padded rows): dram written + read
                         'traffic': SCAN_TRAFFIC if (n == 107526 and K == 100) else None, 'algorithmic_bytes': alg,
                         'peak_source': peak_src, 'bytes_per_nt': alg / (n * Lt)}}


def bench_e2e(model, B, Lt, world, dev, steps):
    """Through the reference-facing call with HOST buffers, every step: a pinned one-hot fp32 minibatch is copied host -> device
    inside `CRBM._trainingFct`, the step runs, and the updated visible bias c (16 bytes) is read device -> host.
    `value`: the read-back is an asynchronous copy into a ring of pinned slots (consumed after a stream synchronisation every
    `ring` steps), so the host keeps issuing steps while the device works — the way a training loop that logs a metric runs.
    `blocking_read_value`: the host blocks on every step's read-back before it issues the next step."""
    import numpy as np
    import torch
    nb, ring = 16, 64
    rs = np.random.RandomState(7)
    host = []
    for i in range(nb):
        codes = torch.from_numpy(rs.randint(0, 4, size=(B, Lt)))
        oh = torch.nn.functional.one_hot(codes, 4).permute(0, 2, 1).reshape(B, 1, 4, Lt).to(torch.float32).contiguous()
        host.append(oh.pin_memory())
    sink = torch.empty((ring, 4), dtype=torch.float32).pin_memory()
    c_dev = model._t_c.reshape(-1)

    def one(i, blocking):
        model._trainingFct(host[i % nb], local_shard=True)          # pinned host one-hot batch in, copied by the call
        sink[i % ring].copy_(c_dev, non_blocking=not blocking)      # device -> host read of the step's result
        if not blocking and i % ring == ring - 1:
            torch.cuda.current_stream().synchronize()               # the ring is consumed: its slots may be overwritten

    def run(blocking, n):
        for i in range(5):
            one(i, True)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            one(i, blocking)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return n * B * Lt * world / (ms * 1e-3)

    steps = max(ring, steps // ring * ring)
    blocking = run(True, min(steps, 1024))
    value = run(False, steps)
    return {'value': value, 'unit': 'nt/s', 'h2d_bytes_per_step': B * 4 * Lt * 4, 'd2h_bytes_per_step': 16, 'steps': steps,
            'blocking_read_value': blocking,
            'call': 'CRBM._trainingFct(one-hot fp32 batch in pinned host memory) + read back c every step (asynchronous copy into a '
                    'pinned ring, stream synchronised every %d steps; blocking_read_value: host blocks on every read-back)' % ring}


if __name__ == '__main__':
    main()
