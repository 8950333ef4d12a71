#!/usr/bin/env python
"""Derives crbm_b200/data/oct4_2bit.npz from the reference's sample dataset (/root/reference/secomo/data/oct4.fa,
3997 Oct4 ChIP-seq peaks x 200 bp; BASELINE config "tutorial"). Build-container only.

The fixture stores the record names, the bases as 2 bits each (A0 C1 G2 T3 — the packed layout the kernels use)
and a 1-bit lower-case mask, so `crbm_b200.sequences.sample_fasta_path()` can re-materialise the exact FASTA.
"""
import os
import numpy as np

SRC = '/root/reference/secomo/data/oct4.fa'
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'crbm_b200', 'data', 'oct4_2bit.npz')
CODE = {'A': 0, 'C': 1, 'G': 2, 'T': 3}

names, seqs = [], []
with open(SRC) as f:
    for line in f:
        line = line.rstrip('\n')
        if line.startswith('>'):
            names.append(line[1:])
            seqs.append('')
        elif line:
            seqs[-1] += line
L = len(seqs[0])
assert all(len(s) == L for s in seqs) and not any('N' in s.upper() for s in seqs)
codes = np.array([[CODE[ch.upper()] for ch in s] for s in seqs], dtype=np.uint8)
lower = np.array([[ch.islower() for ch in s] for s in seqs], dtype=np.uint8)
packed = (codes.reshape(len(seqs), L // 4, 4) << np.array([0, 2, 4, 6], dtype=np.uint8)).sum(axis=2).astype(np.uint8)
np.savez_compressed(DST, names=np.array(names), packed=packed, lower=np.packbits(lower, axis=1), length=np.int64(L))
print(DST, os.path.getsize(DST), codes.shape)
