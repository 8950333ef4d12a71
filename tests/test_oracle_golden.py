"""The NumPy oracle against fixtures produced by the UNMODIFIED reference (convRBM.py on oracle/theano_shim).

This is what pins the oracle: every compiled function of the reference (convRBM.py:459-514) plus the symbolic
builders its tests call, replayed with the same injected uniforms. CPU-only.
"""
import os

import numpy as np
import pytest

from oracle import crbm_oracle as O
from tests._golden import CASE_NAMES, load_case, onehot, draw_for

TOL = dict(rtol=1e-9, atol=1e-11)   # oracle and reference both run in float64 here


def make_oracle(case):
    hp = case['hyper']
    o = O.OracleCRBM(case['W_init'], case['b_init'], case['c_init'], doublestranded=hp['doublestranded'],
                     batchsize=hp['batchsize'], learning_rate=hp['learning_rate'], momentum=hp['momentum'],
                     pooling=hp['pooling'], cd_k=hp['cd_k'], rho=hp['rho'], lambda_rate=hp['lambda_rate'])
    o.fantasy_h = case['fh_init'].astype(np.float64)
    o.fantasy_hp = case['fhp_init'].astype(np.float64) if hp['doublestranded'] else None
    return o


@pytest.mark.parametrize('name', CASE_NAMES)
def test_train_steps_match_reference(name):
    case = load_case(name)
    hp = case['hyper']
    o = make_oracle(case)
    for step in range(hp['steps']):
        o.train_step(onehot(case['codes'][step], np.float64), draw_for(case, step))
        np.testing.assert_allclose(o.W, case['W_%d' % step], **TOL)
        np.testing.assert_allclose(o.b, case['b_%d' % step], **TOL)
        np.testing.assert_allclose(o.c, case['c_%d' % step], **TOL)
        np.testing.assert_allclose(o.vW, case['vW_%d' % step], **TOL)
        np.testing.assert_allclose(o.vb, case['vb_%d' % step], **TOL)
        np.testing.assert_allclose(o.vc, case['vc_%d' % step], **TOL)
        np.testing.assert_array_equal(o.fantasy_h.astype(np.uint8), case['fh_%d' % step])
        if hp['doublestranded']:
            np.testing.assert_array_equal(o.fantasy_hp.astype(np.uint8), case['fhp_%d' % step])


@pytest.mark.parametrize('name', CASE_NAMES)
def test_evaluation_functions_match_reference(name):
    case = load_case(name)
    hp = case['hyper']
    last = hp['steps'] - 1
    W, b, c = case['W_%d' % last], case['b_%d' % last], case['c_%d' % last]
    P, ds = hp['pooling'], hp['doublestranded']
    D = onehot(case['eval_codes'], np.float64)[:, :, :, :hp['L_trunc']]
    np.testing.assert_allclose(O.free_energy(D, W, b, c, P, ds), case['free_energy'], **TOL)
    np.testing.assert_allclose(O.free_energy_per_motif(D, W, b, c, P, ds), case['fe_per_motif'], **TOL)
    np.testing.assert_allclose(O.hit_probs(D, W, b, P, ds), case['hit_probs'], **TOL)
    wn, ic, medic = O.evaluate_params(W)
    np.testing.assert_allclose([wn, ic, medic], [case['eval_wnorm'], case['eval_ic'], case['eval_medic']], **TOL)
    np.testing.assert_allclose(np.stack(O.get_pfms(W)), case['pfms'], **TOL)
    o = make_oracle(case)
    o.W, o.b, o.c = W, b, c
    mfe, nmh = o.evaluate_data(D, draw_for(case, 'eval'))
    np.testing.assert_allclose([mfe, nmh], [case['eval_mfe'], case['eval_nmh']], **TOL)
    for flip in (0, 1):
        x = O.bottom_up_activity(D, W, b, bool(flip))
        np.testing.assert_allclose(x, case['bu_act_flip%d' % flip], **TOL)
        np.testing.assert_allclose(O.bottom_up_probability(x, P), case['bu_prob_flip%d' % flip], **TOL)
    h, hp_ = case['td_h'].astype(np.float64), case['td_hp'].astype(np.float64)
    a = O.top_down_activity(h, hp_, W, c)
    np.testing.assert_allclose(a, case['td_act_ds'], **TOL)
    np.testing.assert_allclose(O.top_down_probability(a), case['td_prob_ds'], **TOL)
    a = O.top_down_activity(h, None, W, c)
    np.testing.assert_allclose(a, case['td_act_ss'], **TOL)
    np.testing.assert_allclose(O.top_down_probability(a), case['td_prob_ss'], **TOL)
    np.testing.assert_allclose(O.top_down_activity(case['td_h_real'], None, W, c), case['td_act_real'], **TOL)


@pytest.mark.skipif(not os.path.isdir('/root/reference/secomo'), reason="the reference tree only exists in the build container")
def test_reference_own_test_suite_passes_on_the_theano_shim():
    """Pins the shim the golden fixtures were generated through: the reference's UNMODIFIED tests/testcrbm.py and
    tests/testloadsamples.py (NumPy control loops for conv2d valid / full / filter_flip, sigmoid, 4-way softmax, one-hot
    sampling; constructor validation; save / load) run from /root/reference on oracle/theano_shim. Skipped where the
    reference tree is absent (GPU boxes); never reads it at import time."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(['bash', os.path.join(root, 'oracle', 'run_reference_tests.sh')], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert '17 passed' in r.stdout, r.stdout[-500:]
