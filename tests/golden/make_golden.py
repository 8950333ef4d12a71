#!/usr/bin/env python
"""Generates tests/golden/case_*.npz by running the UNMODIFIED reference (/root/reference/secomo/convRBM.py,
schulter/crbm 1.1.0) on `oracle/theano_shim` (Theano itself is not installable offline; see the shim's
docstring for what it does and does not reproduce). Build-container only — /root/reference is absent on the
GPU boxes, which is why the outputs are committed.

    python tests/golden/make_golden.py            # rewrites every fixture (deterministic)

Each fixture holds: hyper-parameters, W/b/c (fp32-representable), base codes of the input batch(es), the initial
fantasy particles, every uniform the reference's multinomial ops consumed (per train step / Gibbs iteration,
rows in the reference's own order), and the reference's outputs in float64 (`theano.config.floatX='float64'`):
per-step W,b,c, velocities and fantasy particles after `theano_trainingFct`, plus `theano_freeEnergy`,
`theano_fePerMotif`, `theano_getHitProbs`, `theano_evaluateData`, `theano_evaluateParams`, and the symbolic
builders `_bottomUpActivity/_bottomUpProbability/_topDownActivity/_topDownProbability`.
Uniforms are re-drawn until every multinomial decision has a margin > 2e-5 to its thresholds, so an fp32
implementation replaying them makes identical choices.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'theano_shim'))

import theano  # noqa: E402  (the shim)
import theano.tensor as T  # noqa: E402
from theano.sandbox import rng_mrg  # noqa: E402

theano.config.floatX = 'float64'

spec = importlib.util.spec_from_file_location('ref_convRBM', '/root/reference/secomo/convRBM.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

MARGIN = 2e-5

CASES = {
    # name: K, M, P, ds, cd_k, N(batch), L, steps, wscale, bias
    'small_ds_p1':  dict(K=10, M=5,  P=1, ds=True,  cd_k=1, N=6, L=60,  steps=3, wscale=0.6, bias=-2.5),
    'pool4_ds_cd2': dict(K=7,  M=6,  P=4, ds=True,  cd_k=2, N=5, L=53,  steps=3, wscale=0.5, bias=-2.0),
    'pool2_ss_cd2': dict(K=5,  M=4,  P=2, ds=False, cd_k=2, N=4, L=41,  steps=3, wscale=0.7, bias=-1.5),
    'k33_m15_ds':   dict(K=33, M=15, P=1, ds=True,  cd_k=1, N=4, L=200, steps=2, wscale=0.3, bias=-4.0),
    'k40_m20_p5':   dict(K=40, M=20, P=5, ds=True,  cd_k=1, N=3, L=119, steps=2, wscale=0.3, bias=-4.0),
    'shortbatch':   dict(K=6,  M=5,  P=1, ds=True,  cd_k=1, N=3, L=40,  steps=2, wscale=0.6, bias=-2.0, batchsize=5),
}


def onehot(codes):
    N, L = codes.shape
    D = np.zeros((N, 1, 4, L), dtype=np.float64)
    for a in range(4):
        D[:, 0, a, :] = (codes == a)
    return D


def build_model(cfg, W, b, c, fh, fhp):
    kw = dict(num_motifs=cfg['K'], motif_length=cfg['M'], epochs=1, doublestranded=cfg['ds'],
              batchsize=cfg.get('batchsize', cfg['N']), learning_rate=0.1, momentum=0.95, pooling=cfg['P'],
              cd_k=cfg['cd_k'], rho=0.01, lambda_rate=0.1)
    node0 = rng_mrg._node_counter[0]
    model = ref.CRBM(**kw)
    model.motifs.set_value(W)
    model.bias.set_value(b)
    model.c.set_value(c)
    model.fantasy_h.set_value(fh)
    if cfg['ds']:
        model.fantasy_h_prime.set_value(fhp)
    # node numbering follows graph construction order in convRBM.py:373-451, 469
    roles = {}
    n = node0
    roles[n] = 'pos_h'; n += 1
    if cfg['ds']:
        roles[n] = 'pos_hp'; n += 1
    for it in range(cfg['cd_k']):
        roles[n] = 'v%d' % it; n += 1
        roles[n] = 'h%d' % it; n += 1
        if cfg['ds']:
            roles[n] = 'hp%d' % it; n += 1
    roles[n] = 'sparsity_h'; n += 1
    roles[n] = 'eval_h'; n += 1
    return model, roles, kw


def run_case(name, cfg, seed):
    rs = np.random.RandomState(seed)
    K, M, P, N, L = cfg['K'], cfg['M'], cfg['P'], cfg['N'], cfg['L']
    bs = cfg.get('batchsize', N)
    W = (rs.randn(K, 1, 4, M) * cfg['wscale']).astype(np.float32).astype(np.float64)
    b = (cfg['bias'] + 0.3 * rs.randn(1, K)).astype(np.float32).astype(np.float64)
    c = (0.2 * rs.randn(1, 4)).astype(np.float32).astype(np.float64)
    codes = rs.randint(0, 4, size=(cfg['steps'], N, L)).astype(np.uint8)
    eval_codes = rs.randint(0, 4, size=(N, L)).astype(np.uint8)
    fh = (rs.rand(bs, K, 1, 200) < 0.03).astype(np.float64)
    fhp = (rs.rand(bs, K, 1, 200) < 0.03).astype(np.float64)
    if P > 1:  # at most one active unit per pooling group, like any state the sampler can produce
        for arr in (fh, fhp):
            g = arr.reshape(bs, K, 1, 200 // P, P)
            g[...] = g * (np.cumsum(g, axis=4) == 1)

    uniforms = {}   # (step, role) -> float32 array

    def source_for(step, roles):
        def src(node, rows):
            key = (step, roles[node])
            if key not in uniforms or uniforms[key].shape[0] != rows:
                uniforms[key] = rs.rand(rows).astype(np.float32)
            return uniforms[key].astype(np.float64)
        return src

    for attempt in range(50):
        rng_mrg.reset_hooks()
        model, roles, kw = build_model(cfg, W, b, c, fh, fhp)
        out = {}
        bad = 0
        for step in range(cfg['steps']):
            rng_mrg.uniform_source = source_for(step, roles)
            del rng_mrg.draw_log[:]
            model.theano_trainingFct(onehot(codes[step]))
            for d in rng_mrg.draw_log:
                key = (step, roles[d['node']])
                csum = np.cumsum(d['pvals'], axis=1)
                margin = np.min(np.abs(csum - d['uniforms'][:, None]), axis=1)
                rows = np.nonzero(margin <= MARGIN)[0]
                if rows.size:
                    bad += rows.size
                    uniforms[key][rows] = rs.rand(rows.size).astype(np.float32)
            if bad:
                break
            out['W_%d' % step] = model.motifs.get_value()
            out['b_%d' % step] = model.bias.get_value()
            out['c_%d' % step] = model.c.get_value()
            out['vW_%d' % step] = model.motif_velocity.get_value()
            out['vb_%d' % step] = model.bias_velocity.get_value()
            out['vc_%d' % step] = model.c_velocity.get_value()
            out['fh_%d' % step] = model.fantasy_h.get_value().astype(np.uint8)
            if cfg['ds']:
                out['fhp_%d' % step] = model.fantasy_h_prime.get_value().astype(np.uint8)
        if bad:
            continue
        # evaluation functions on the trained parameters
        De = onehot(eval_codes)
        Lt = int((L - M + 1) / P) * P + M - 1
        De = De[:, :, :, :Lt]
        rng_mrg.uniform_source = source_for('eval', roles)
        del rng_mrg.draw_log[:]
        mfe, nmh = model.theano_evaluateData(De)
        d = rng_mrg.draw_log[0]
        margin = np.min(np.abs(np.cumsum(d['pvals'], axis=1) - d['uniforms'][:, None]), axis=1)
        rows = np.nonzero(margin <= MARGIN)[0]
        if rows.size:
            uniforms[('eval', 'eval_h')][rows] = rs.rand(rows.size).astype(np.float32)
            continue
        break
    else:
        raise RuntimeError("could not find well-separated uniforms for %s" % name)

    out['eval_mfe'], out['eval_nmh'] = np.float64(mfe), np.float64(nmh)
    wn, ic, medic = model.theano_evaluateParams()
    out['eval_wnorm'], out['eval_ic'], out['eval_medic'] = np.float64(wn), np.float64(ic), np.float64(medic)
    out['free_energy'] = model.theano_freeEnergy(De)
    out['fe_per_motif'] = model.theano_fePerMotif(De)
    out['hit_probs'] = model.theano_getHitProbs(De)
    out['pfms'] = np.stack(model.getPFMs())

    # symbolic builders, as the reference tests call them (tests/testcrbm.py:163-171, 224-247)
    i1, i2 = T.tensor4(), T.tensor4()
    for flip in (False, True):
        f = theano.function([i1], [model._bottomUpActivity(i1, flip),
                                   model._bottomUpProbability(model._bottomUpActivity(i1, flip))])
        a, p = f(De)
        out['bu_act_flip%d' % flip], out['bu_prob_flip%d' % flip] = a, p
    Hd = (rs.rand(N, K, 1, Lt - M + 1) < 0.1).astype(np.float64)
    Hdp = (rs.rand(N, K, 1, Lt - M + 1) < 0.1).astype(np.float64)
    f = theano.function([i1, i2], [model._topDownActivity(i1, i2),
                                   model._topDownProbability(model._topDownActivity(i1, i2))])
    out['td_act_ds'], out['td_prob_ds'] = f(Hd, Hdp)
    f = theano.function([i1], [model._topDownActivity(i1, None),
                               model._topDownProbability(model._topDownActivity(i1, None))])
    out['td_act_ss'], out['td_prob_ss'] = f(Hd)
    # real-valued hidden input (the reference tests feed probabilities/samples alike)
    Hr = rs.rand(N, K, 1, Lt - M + 1).astype(np.float32).astype(np.float64)
    out['td_act_real'] = theano.function([i1], model._topDownActivity(i1, None))(Hr)

    out.update(dict(W_init=W, b_init=b, c_init=c, codes=codes, eval_codes=eval_codes,
                    fh_init=fh.astype(np.uint8), fhp_init=fhp.astype(np.uint8),
                    td_h=Hd.astype(np.uint8), td_hp=Hdp.astype(np.uint8), td_h_real=Hr))
    for (step, role), u in uniforms.items():
        out['u_%s_%s' % (step, role)] = u
    hp = dict(kw)
    hp.update(steps=cfg['steps'], N=N, L=L, L_trunc=Lt)
    out['hyper'] = np.array(repr(hp))
    path = os.path.join(HERE, 'case_%s.npz' % name)
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %d attempts, %.1f KB)" % (path, len(out), attempt + 1, os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    for i, (name, cfg) in enumerate(sorted(CASES.items())):
        run_case(name, cfg, seed=1000 + i)
