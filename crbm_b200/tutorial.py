"""End-to-end demo on the bundled Oct4 ChIP-seq sample: the call sequence of the reference's secomo/tutorial.py:6-67
(BASELINE config "tutorial") with every plotting call replaced by the numbers that plot is drawn from.

    load_sample -> CRBM(10 motifs x 15 bp, 1 epoch).fit -> saveModel / loadModel -> getPFMs -> saveMotifs -> motifHitProbs

Sequence logos, positional density, t-SNE and violin plots need matplotlib / seaborn / weblogo / sklearn (out of
scope). What they visualise is written to `motif_hit_summary.npz` instead:
  * max / argmax / mean of the hit probabilities over positions per (sequence, motif) — computed inside the scan kernel
    (`CRBM.motifHitSummary`; the reductions of utils.py:155,245-247,306),
  * the mean hit probability per (motif, position) over the sequences (utils.py:116), reduced on the device
    (`CRBM.motifHitPositionMean`).
"""
import os

import numpy as np

from .convRBM import CRBM
from .sequences import load_sample
from .utils import saveMotifs


def tutorial(path, num_motifs=10, motif_length=15, epochs=1):
    """Runs the demo, writes `oct4_model_params.pkl`, `pfms/` and `motif_hit_summary.npz` under `path`, returns the results."""
    data = load_sample()                                         # (3997, 1, 4, 200) one-hot
    trained = CRBM(num_motifs=num_motifs, motif_length=motif_length, epochs=epochs)
    trained.fit(data)                                            # one epoch: a smoke run, not a converged model

    params = os.path.join(path, 'oct4_model_params.pkl')
    trained.saveModel(params)
    model = CRBM.loadModel(params)                               # round trip through the reference's pickle layout

    pfms = model.getPFMs()
    saveMotifs(model, path=os.path.join(path, 'pfms'))           # one JASPAR-format file per motif

    head = data[:100]
    matches = model.motifHitProbs(head)                          # (100, K, 1, 186) per-position match probabilities
    summary = model.motifHitSummary(data)                        # fused reductions over all 3997 sequences
    density = model.motifHitPositionMean(data)                   # (K, positions), reduced on the device over all sequences
    np.savez(os.path.join(path, 'motif_hit_summary.npz'), max=summary['max'], argmax=summary['argmax'], mean=summary['mean'],
             positional_density=density)
    return {'model': model, 'pfms': pfms, 'matches': matches, 'summary': summary, 'positional_density': density}
