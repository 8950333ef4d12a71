"""The end-to-end demo of the reference (secomo/tutorial.py:6-67), minus the plots.

Same call sequence on the bundled Oct4 ChIP-seq sample: load -> CRBM(10 motifs x 15 bp, 1 epoch) -> fit -> saveModel /
loadModel -> getPFMs -> saveMotifs -> motifHitProbs. The plotting calls of the reference (sequence logos, positional density,
t-SNE, violin plots: matplotlib / seaborn / weblogo / sklearn, SURVEY §2 out of scope) are replaced by the numbers they are drawn
from, computed inside the scan kernel and written to `motif_hit_summary.npz`:
  * max / argmax / mean of the hit probabilities over positions per sequence and motif (utils.py:155,245-247,306), and
  * the mean hit probability per motif and position over the sequences (`positionalDensityPlot`, utils.py:116).
"""
import os

import numpy as np

from .convRBM import CRBM
from .sequences import load_sample
from .utils import saveMotifs


def tutorial(path):
    # one-hot (3997, 1, 4, 200) array of the bundled sample
    onehot = load_sample()

    # Epoch = 1 is not sufficient to train meaningful features, however, it speeds up automatic testing.
    model = CRBM(num_motifs=10, motif_length=15, epochs=1)
    model.fit(onehot)

    # save parameters and hyper-parameters, re-instantiate the model from them
    model.saveModel(os.path.join(path, 'oct4_model_params.pkl'))
    model = CRBM.loadModel(os.path.join(path, 'oct4_model_params.pkl'))

    # position frequency matrices, stored (in 'jaspar' format) in <path>/pfms/
    pfms = model.getPFMs()
    saveMotifs(model, path=os.path.join(path, 'pfms'))

    # per-position motif match probabilities for the first 100 sequences
    matches = model.motifHitProbs(onehot[:100])

    # what the reference's plots show, as numbers
    summary = model.motifHitSummary(onehot)
    density = matches.mean(axis=(0, 2))                      # (motifs, positions), utils.py:116
    np.savez(os.path.join(path, 'motif_hit_summary.npz'), max=summary['max'], argmax=summary['argmax'], mean=summary['mean'],
             positional_density=density)
    return {'model': model, 'pfms': pfms, 'matches': matches, 'summary': summary, 'positional_density': density}
