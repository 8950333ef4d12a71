"""crbm_b200 — B200-native implementation of SECOMO's cRBM training / motif-scanning hot path.

Drop-in for the hot-path names of `secomo` (reference secomo/__init__.py:1-9): CRBM, seqToOneHot,
readSeqsFromFasta, load_sample, splitTrainingTest, saveMotifs. The plotting helpers of secomo.utils are out of scope.
"""
from .convRBM import CRBM
from .sequences import (seqToOneHot, readSeqsFromFasta, load_sample, splitTrainingTest, PackedSequences,
                        sample_fasta_path, seqToCodes, codesToOneHot, packCodesHost, unpackCodesHost, packedRowWords)
from .utils import saveMotifs, formatMotif

__version__ = '0.1.0'
__all__ = ['CRBM', 'seqToOneHot', 'readSeqsFromFasta', 'load_sample', 'splitTrainingTest', 'PackedSequences',
           'sample_fasta_path', 'seqToCodes', 'codesToOneHot', 'saveMotifs', 'formatMotif', 'packCodesHost', 'unpackCodesHost',
           'packedRowWords']
