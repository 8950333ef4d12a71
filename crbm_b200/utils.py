"""Model artefacts of `secomo/utils.py` that need no plotting stack.

`saveMotifs` (reference secomo/utils.py:16-47) writes every position frequency matrix of a model (`CRBM.getPFMs()`,
convRBM.py:640-655) to `<path>/<name><i>.pfm`. The reference builds a `Bio.motifs.jaspar.Motif` and formats it with
`Bio.motifs.jaspar.write`; Biopython is not a dependency here, so the two text layouts of that writer are produced directly:

    jaspar:   >mot1 mot1                      pfm:     0.25   0.10 ...
              A [  0.25   0.10 ... ]                   0.25   0.40 ...
              C [  0.25   0.40 ... ]                   ...
              ...

(one row per nucleotide in the order A, C, G, T, values as `%6.2f`). The plotting helpers of secomo/utils.py (sequence logos,
positional density, t-SNE, violin plots) consume only `getPFMs()` / `motifHitProbs()` outputs and are out of scope.
"""
import os

_LETTERS = 'ACGT'


def formatMotif(pfm, motif_id, fformat='jaspar'):
    """Text of one (4, M) position frequency matrix in the JASPAR-family layout `fformat` ('jaspar' or 'pfm')."""
    rows = [' '.join('{0:6.2f}'.format(float(v)) for v in pfm[a]) for a in range(4)]
    if fformat == 'jaspar':
        return '>{0} {1}\n'.format(motif_id, motif_id) + ''.join('{0} [{1}]\n'.format(_LETTERS[a], rows[a]) for a in range(4))
    if fformat == 'pfm':
        return ''.join(row + '\n' for row in rows)
    raise ValueError("Unknown JASPAR format %s" % fformat)


def saveMotifs(model, path, name="mot", fformat='jaspar'):
    """Save the weight matrices of `model` as PFMs, one file `<name><i>.pfm` per motif (i = 0 .. K-1; the motif inside
    file i is called `<name><i+1>`, as in the reference)."""
    pfms = model.getPFMs()
    for i, pfm in enumerate(pfms):
        text = formatMotif(pfm, name + str(i + 1), fformat)
        if not os.path.exists(path):
            os.makedirs(path)
        with open(os.path.join(path, "{}{:d}.{}".format(name, i, "pfm")), "w") as f:
            f.write(text)
