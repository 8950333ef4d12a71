"""Sequence input for crbm_b200: the `secomo.sequences` interface without Biopython.

Mirrors reference secomo/sequences.py — `readSeqsFromFasta` (:33-52, records containing N/n are skipped with the
same message), `seqToOneHot` (:101-117 -> (N,1,4,L) float32, A0 C1 G2 T3 case-insensitive, KeyError for any other
letter as the reference's table lookup raises), `load_sample` (:120-137) and `splitTrainingTest` (:54-98).
Biopython is not available here, so records are a minimal `SeqRecord`/`Seq` pair with the attributes callers use
(`.id`, `.seq`, `len()`, indexing, `str()`).

`PackedSequences` is the B200-native container the kernels actually read: 2 bits per base, resident on the GPU.
`CRBM.fit` / `motifHitProbs` / `freeEnergy` accept it in place of the 16-byte-per-base one-hot array.
"""
import os
import re
import tempfile

import numpy as np

L = {'A': 0, 'a': 0, 'C': 1, 'c': 1, 'G': 2, 'g': 2, 'T': 3, 't': 3}

_CODE_LUT = np.full(256, 255, dtype=np.uint8)
for _ch, _v in L.items():
    _CODE_LUT[ord(_ch)] = _v


class _Alphabet(object):
    letters = 'GATC'          # Bio.Alphabet.IUPAC.unambiguous_dna.letters


class Seq(str):
    """A DNA string with the `.alphabet.letters` attribute the reference's `_getOneHotSeq` reads."""
    alphabet = _Alphabet()


class SeqRecord(object):
    def __init__(self, seq, id='<unknown id>', description=''):
        self.seq = seq if isinstance(seq, Seq) else Seq(seq)
        self.id = self.name = id
        self.description = description or id

    def __len__(self):
        return len(self.seq)

    def __repr__(self):
        return "SeqRecord(id=%r, len=%d)" % (self.id, len(self.seq))


def _parse_fasta(handle):
    name, desc, chunks = None, None, []
    for line in handle:
        line = line.strip()
        if not line:
            continue
        if line.startswith('>'):
            if name is not None:
                yield SeqRecord(''.join(chunks), id=name, description=desc)
            desc = line[1:]
            name = desc.split()[0] if desc.split() else ''
            chunks = []
        else:
            chunks.append(line.replace(' ', '').replace('\t', ''))     # (Biopython's reader drops blanks inside a record too)
    if name is not None:
        yield SeqRecord(''.join(chunks), id=name, description=desc)


def readSeqsFromFasta(filename):
    """Read sequences from a multi-fasta file; records containing N are skipped (reference sequences.py:33-52)."""
    seqs = []
    with open(filename) as handle:
        for faseq in _parse_fasta(handle):
            if re.search(r'N', str(faseq.seq), re.I):
                print("skip sequence containing N")
                continue
            seqs.append(faseq)
    return seqs


def seqToCodes(seqs):
    """List of records / strings of equal length -> (N,L) uint8 base codes. KeyError on a non-ACGT letter."""
    rows = []
    for s in seqs:
        text = str(s.seq) if hasattr(s, 'seq') else str(s)
        codes = _CODE_LUT[np.frombuffer(text.encode('latin-1'), dtype=np.uint8)]
        bad = np.nonzero(codes == 255)[0]
        if bad.size:
            raise KeyError(text[int(bad[0])])
        rows.append(codes)
    if not rows:
        return np.zeros((0, 0), dtype=np.uint8)
    if len({r.shape[0] for r in rows}) != 1:
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    return np.stack(rows)


# ---- streaming ingest ---------------------------------------------------------------------------------------------
_N_CODE = 4                                             # 'N' / 'n': the record is skipped (reference sequences.py:47-50)
_STREAM_LUT = _CODE_LUT.copy()
_STREAM_LUT[ord('N')] = _STREAM_LUT[ord('n')] = _N_CODE


def _iter_fasta_codes_native(lib, filename, chunk_bytes, quiet):
    """iterFastaCodes on the one-pass C parser of the library (crbm_fasta_parse_block_host, csrc/api.cu)."""
    import ctypes
    carry = b''
    L0 = ctypes.c_int(0)
    nrec, nskip, bad = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_int(0)
    with open(filename, 'rb') as f:
        while True:
            block = f.read(chunk_bytes)
            last = not block
            buf = carry + block
            if not buf:
                return
            if not last:
                cut = buf.rfind(b'\n>')                      # the last record of the block may be incomplete: keep it for the next one
                if cut < 0:
                    carry = buf
                    continue
                buf, carry = buf[:cut + 1], buf[cut + 1:]
            else:
                carry = b''
            # every record needs at least 3 bytes ('>', newline, one base): an upper bound on the records of the block
            cap = len(buf) // 3 + 1
            width = L0.value if L0.value > 0 else len(buf)
            if L0.value > 0:
                cap = min(cap, len(buf) // (L0.value + 2) + 1)
            out = np.empty(cap * width if L0.value > 0 else len(buf), dtype=np.uint8)
            rc = lib.crbm_fasta_parse_block_host(buf, len(buf), out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(L0),
                                                 ctypes.byref(nrec), ctypes.byref(nskip), ctypes.byref(bad))
            if rc == -11:
                raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
            if rc == -12:
                raise KeyError(chr(bad.value))
            if rc != 0:
                raise RuntimeError("crbm_fasta_parse_block_host failed (%d)" % rc)
            if not quiet:
                for _ in range(int(nskip.value)):
                    print("skip sequence containing N")
            if nrec.value:
                yield out[:nrec.value * L0.value].reshape(int(nrec.value), L0.value)
            if last:
                return


def iterFastaCodes(filename, chunk_bytes=32 << 20, quiet=False, native=None):
    """Streams a multi-FASTA file as (n, L) uint8 base-code arrays (A0 C1 G2 T3, case-insensitive), one per `chunk_bytes` of file,
    without ever holding the file or a per-record Python object: a block of bytes is split into records with NumPy (header lines
    and line ends removed by masks, record lengths by a segmented count), so multi-line records and CRLF files are handled.
    Semantics of the reference reader (secomo/sequences.py:33-52, 20-31): records containing N/n are skipped ("skip sequence
    containing N" is printed once per skipped record unless `quiet`), any other non-ACGT letter raises KeyError, and all records
    must have the same length (the reference's one-hot array cannot be built otherwise)."""
    lib = None
    if native is None or native:                             # native=False: the NumPy parser below (same results, tested against it)
        try:
            from . import _lib
            lib = _lib.load()
        except Exception:                                    # library not built: host-side parsing needs no GPU, NumPy does the same job
            if native:
                raise
    if lib is not None:
        for codes in _iter_fasta_codes_native(lib, filename, chunk_bytes, quiet):
            yield codes
        return
    carry = b''
    L0 = None
    with open(filename, 'rb') as f:
        while True:
            block = f.read(chunk_bytes)
            last = not block
            buf = carry + block
            if not buf:
                return
            if not last:
                cut = buf.rfind(b'>')                       # the last record of the block may be incomplete: keep it for the next one
                if cut <= 0:
                    carry = buf
                    continue
                buf, carry = buf[:cut], buf[cut:]
            else:
                carry = b''
            a = np.frombuffer(buf, dtype=np.uint8)
            nl = a == 10
            start = np.concatenate(([True], nl[:-1]))        # first byte of every line
            hdr_start = np.flatnonzero(start & (a == ord('>')))
            if hdr_start.size == 0:
                if last:
                    return
                continue
            nl_pos = np.flatnonzero(nl)
            # end of every header line (its newline, or the end of the buffer)
            idx = np.searchsorted(nl_pos, hdr_start)
            if nl_pos.size:
                hdr_end = np.where(idx < nl_pos.size, nl_pos[np.minimum(idx, nl_pos.size - 1)], a.size - 1)
            else:
                hdr_end = np.full(hdr_start.shape, a.size - 1)
            is_hdr = np.zeros(a.size + 1, dtype=np.int8)
            np.add.at(is_hdr, hdr_start, 1)
            np.add.at(is_hdr, hdr_end + 1, -1)
            in_hdr = np.cumsum(is_hdr[:-1]) > 0
            keep = ~(in_hdr | nl | (a == 13) | (a == 32) | (a == 9))
            keep[:hdr_start[0]] = False                      # anything before the first header is not a record
            rec_of_byte = np.cumsum(start & (a == ord('>'))) - 1
            lengths = np.bincount(rec_of_byte[keep], minlength=hdr_start.size)
            raw = a[keep]
            seq = _STREAM_LUT[raw]
            lengths = lengths[lengths > 0]                   # header without sequence lines: an empty record, dropped
            if lengths.size:
                starts = np.concatenate(([0], np.cumsum(lengths)[:-1]))
                # a record containing N is skipped whatever else it holds (the reference drops it before it looks at anything else)
                has_n = np.add.reduceat((seq == _N_CODE).astype(np.int64), starts) > 0
                if not quiet:
                    for _ in range(int(has_n.sum())):
                        print("skip sequence containing N")
                good = ~has_n
                if good.any():
                    bad_rec = (np.add.reduceat((seq == 255).astype(np.int64), starts) > 0)[good]
                    glen = lengths[good]
                    if L0 is None:
                        L0 = int(glen[0])
                    off = bad_rec | (glen != L0)
                    if off.any():                           # the first offending record in file order
                        k = int(np.argmax(off))
                        if bad_rec[k]:
                            g = int(np.flatnonzero(good)[k])
                            rec = slice(int(starts[g]), int(starts[g] + lengths[g]))
                            raise KeyError(chr(int(raw[rec][np.argmax(seq[rec] == 255)])))
                        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
                    codes = seq[np.repeat(good, lengths)].reshape(-1, L0)
                    if codes.shape[0]:
                        yield np.ascontiguousarray(codes)
            if last:
                return


def iterFastaPacked(filename, chunk_bytes=32 << 20, quiet=False, threads=0, get_buffer=None):
    """Streams a multi-FASTA file as rows of 2-bit bases in the device layout (packedRowWords(L) uint32 words per record): yields
    (rows, L) with rows an (n, row_words) uint32 array per block of about `chunk_bytes` of file. Same record semantics as
    iterFastaCodes (reference secomo/sequences.py:33-52, 20-31: N records skipped, case folded, KeyError on other letters, equal
    lengths). The file is memory-mapped and cut into blocks at record starts; a block is parsed and packed by the host threads of
    crbm_fasta_pack_block_host (csrc/ingest.cu) straight into the array `get_buffer(rows_needed, row_words)` returns (default: a
    fresh NumPy array; PackedSequences.from_fasta hands out page-locked staging buffers). Needs the built library."""
    import ctypes
    import mmap
    from . import _lib
    lib = _lib.load()
    size = os.path.getsize(filename)
    if size == 0:
        return
    if get_buffer is None:
        get_buffer = lambda rows, rw: np.empty((rows, rw), dtype=np.uint32)
    L0 = ctypes.c_int(0)
    nrec, nskip, bad = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_int(0)
    chunk_bytes = max(1, int(chunk_bytes))
    with open(filename, 'rb') as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    view = np.frombuffer(mm, dtype=np.uint8)
    try:
        addr = view.ctypes.data
        lo = 0
        while lo < size:
            hi = min(size, lo + chunk_bytes)
            if hi < size:                                   # a block ends where a record starts
                cut = mm.rfind(b'\n>', lo, hi + 1)
                if cut < 0:
                    cut = mm.find(b'\n>', hi)
                hi = size if cut < 0 else cut + 1
            n = hi - lo
            rows, rw = None, 0
            if L0.value > 0:
                rw = packedRowWords(L0.value)
                rows = get_buffer(n // (L0.value + 2) + 1, rw)
            for attempt in range(2):
                rc = lib.crbm_fasta_pack_block_host(ctypes.c_void_p(addr + lo), n, None if rows is None else rows.ctypes.data_as(ctypes.c_void_p),
                                                    rw, 0 if rows is None else rows.shape[0], ctypes.byref(L0), ctypes.byref(nrec),
                                                    ctypes.byref(nskip), ctypes.byref(bad), int(threads))
                if rc != -13:
                    break
                rw = packedRowWords(L0.value)                # the block may need more rows than its size suggests (or L was unknown)
                rows = get_buffer(int(nrec.value), rw)
            if rc == -11:
                raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
            if rc == -12:
                raise KeyError(chr(bad.value))
            if rc != 0:
                raise RuntimeError("crbm_fasta_pack_block_host failed (%d): %s" % (rc, lib.crbm_last_error().decode()))
            if not quiet:
                for _ in range(int(nskip.value)):
                    print("skip sequence containing N")
            if nrec.value:
                yield rows[:int(nrec.value)], int(L0.value)
            lo = hi
    finally:
        del view
        mm.close()


def codesToOneHot(codes):
    codes = np.asarray(codes)
    N, Ln = codes.shape
    out = np.zeros((N, 1, 4, Ln), dtype=np.float32)
    n_idx, i_idx = np.meshgrid(np.arange(N), np.arange(Ln), indexing='ij')
    out[n_idx, 0, codes, i_idx] = 1
    return out


def seqToOneHot(seqs):
    """Converts a set of DNA sequences to one-hot encoding, (N,1,4,L) float32 (reference sequences.py:101-117)."""
    return codesToOneHot(seqToCodes(seqs))


_SAMPLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'oct4_2bit.npz')


def _sample_records():
    z = np.load(_SAMPLE)
    names, packed, length = z['names'], z['packed'], int(z['length'])
    lower = np.unpackbits(z['lower'], axis=1)[:, :length].astype(bool)
    codes = ((packed[:, :, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(packed.shape[0], -1)[:, :length]
    upper = np.frombuffer(b'ACGT', dtype=np.uint8)[codes]
    text = np.where(lower, upper + 32, upper).astype(np.uint8)
    return [(str(n), text[i].tobytes().decode('ascii')) for i, n in enumerate(names)]


def sample_fasta_path(directory=None):
    """Materialises the bundled Oct4 ChIP-seq sample (3997 x 200 bp; the reference's secomo/data/oct4.fa) as FASTA."""
    directory = directory or os.path.join(tempfile.gettempdir(), 'crbm_b200_data')
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, 'oct4.fa')
    if not os.path.exists(path):
        tmp = path + '.%d.tmp' % os.getpid()
        with open(tmp, 'w') as f:
            for name, seq in _sample_records():
                f.write('>%s\n%s\n' % (name, seq))
        os.replace(tmp, path)
    return path


def load_sample():
    """Sample Oct4 ChIP-seq peaks (H1-hESC, ENCODE) in one-hot encoding, (3997,1,4,200) (reference :120-137)."""
    return seqToOneHot(readSeqsFromFasta(sample_fasta_path()))


def splitTrainingTest(filename, train_test_ratio, num_top_regions=None, randomize=True):
    """Splits a fasta file into <prefix>_train.fa / <prefix>_test.fa (reference sequences.py:54-98)."""
    seqs = readSeqsFromFasta(filename)
    if num_top_regions:
        seqs = seqs[:num_top_regions]
    if randomize:
        idx_permut = list(np.random.permutation(len(seqs)))
    else:
        idx_permut = list(range(len(seqs)))
    ntest = int(len(seqs) * train_test_ratio)
    itest, itrain = idx_permut[:ntest], idx_permut[ntest:]
    prefix = ".".join(filename.split(".")[:-1])
    for path, idx in ((prefix + "_train.fa", itrain), (prefix + "_test.fa", itest)):
        with open(path, 'w') as f:
            for i in idx:
                f.write('>%s\n%s\n' % (seqs[i].description, str(seqs[i].seq)))


# ---- the packed row layout on the host (include/crbm_b200.h; csrc/common.cuh packed_row_words) ---------------------------------
# A row = 1 zero pad word, ceil(L/16) data words (base i at bits 2*(i%16) of word 1 + i/16, A0 C1 G2 T3), >= 7 zero pad words, the
# stride rounded up to a multiple of 4 words. Packing on the host lets a FASTA collection be streamed to the GPU at 0.25 byte per
# base (64x less than the fp32 one-hot array of the reference) without ever building the one-hot array.
_FRONT_PAD_WORDS, _BACK_PAD_WORDS = 1, 7


def packedRowWords(L):
    return 4 * ((_FRONT_PAD_WORDS + (int(L) + 15) // 16 + _BACK_PAD_WORDS + 3) // 4)


def packCodesHost(codes):
    """(N,L) uint8 base codes in {0,1,2,3} -> (N, packedRowWords(L)) uint32 rows in the device layout (NumPy, no GPU)."""
    codes = np.asarray(codes)
    if codes.ndim != 2:
        raise ValueError("expected base codes of shape (N,L), got %r" % (codes.shape,))
    if codes.size and int(codes.max()) > 3:
        raise ValueError("base codes must be in {0,1,2,3}")
    N, Ln = codes.shape
    nw = (Ln + 15) // 16
    padded = np.zeros((N, nw * 16), dtype=np.uint32)
    padded[:, :Ln] = codes
    words = (padded.reshape(N, nw, 16) << (2 * np.arange(16, dtype=np.uint32))).sum(axis=2, dtype=np.uint32)
    out = np.zeros((N, packedRowWords(Ln)), dtype=np.uint32)
    out[:, _FRONT_PAD_WORDS:_FRONT_PAD_WORDS + nw] = words
    return out


def packOneHotHost(data, out=None, threads=0):
    """(N,1,4,L) float32 one-hot in host memory -> ((N, packedRowWords(L)) uint32 rows in the device layout, number of positions
    that are not one-hot), on the host threads of the library (crbm_pack_onehot_host; no GPU). `data` may be a view whose positions
    are contiguous (fit()'s truncation `data[:, :, :, :n]`); `out` an array to fill (a page-locked staging buffer)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    if data.ndim != 4 or data.shape[1] != 1 or data.shape[2] != 4:
        raise ValueError("expected one-hot data of shape (N,1,4,L), got %r" % (data.shape,))
    N, Ln = data.shape[0], data.shape[3]
    if data.dtype != np.float32 or (Ln > 1 and data.strides[3] != 4) or data.strides[0] % 4 or data.strides[2] % 4 \
            or data.strides[0] < 0 or data.strides[2] < 0:
        data = np.ascontiguousarray(data, dtype=np.float32)
    rw = packedRowWords(Ln)
    if out is None:
        out = np.empty((N, rw), dtype=np.uint32)
    if out.shape[0] < N or out.shape[1] != rw or out.dtype != np.uint32 or not out.flags.c_contiguous:
        raise ValueError("packOneHotHost: `out` must be a C-contiguous (>= %d, %d) uint32 array" % (N, rw))
    bad = ctypes.c_longlong(0)
    if N:
        _lib.check(lib.crbm_pack_onehot_host(data.ctypes.data_as(ctypes.c_void_p), N, Ln, data.strides[0] // 4, data.strides[2] // 4,
                                             out.ctypes.data_as(ctypes.c_void_p), rw, ctypes.byref(bad), int(threads)), 'pack_onehot_host')
    return out[:N], int(bad.value)


def unpackCodesHost(words, L):
    """Inverse of packCodesHost: (N, row_words) uint32 rows -> (N,L) uint8 base codes."""
    words = np.asarray(words, dtype=np.uint32)
    nw = (int(L) + 15) // 16
    data = words[:, _FRONT_PAD_WORDS:_FRONT_PAD_WORDS + nw]
    codes = (data[:, :, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3
    return codes.reshape(words.shape[0], nw * 16)[:, :int(L)].astype(np.uint8)


def ctypes_stream(stream):
    import ctypes
    return ctypes.c_void_p(stream.cuda_stream)


class PackedSequences(object):
    """N sequences of L bases, 2 bits per base, resident in GPU memory in the row layout of include/crbm_b200.h."""

    def __init__(self, words, N, L):
        self.words, self.N, self.L = words, int(N), int(L)
        self.row_words = int(words.shape[1])

    @property
    def shape(self):
        return (self.N, 1, 4, self.L)

    def __len__(self):
        return self.N

    @staticmethod
    def _alloc(N, L, device):
        import torch
        from . import _lib
        rw = _lib.load().crbm_packed_row_words(int(L))
        return torch.zeros((int(N), rw), dtype=torch.int32, device=device), rw

    @classmethod
    def from_onehot(cls, data, device='cuda', chunk_rows=None, threads=0):
        """(N,1,4,L) one-hot (NumPy or torch) -> packed and validated (non-one-hot input is rejected).

        A NumPy array (what seqToOneHot returns, 16 bytes per base) is packed where it lies by the host threads of the library
        (packOneHotHost -> crbm_pack_onehot_host) into two page-locked staging buffers in turn, and each chunk of rows goes to the
        device as one asynchronous copy of 0.25 byte per base while the next chunk is packed. A torch tensor (on any device) is
        packed by crbm_pack_onehot on the GPU."""
        import torch
        from . import _lib
        lib = _lib.load()
        if isinstance(data, np.ndarray):
            if data.ndim != 4 or data.shape[1] != 1 or data.shape[2] != 4:
                raise ValueError("expected one-hot data of shape (N,1,4,L), got %r" % (data.shape,))
            N, Ln = data.shape[0], data.shape[3]
            dev = torch.device(device)
            rw = lib.crbm_packed_row_words(int(Ln))
            words = torch.empty((int(N), rw), dtype=torch.int32, device=dev)
            chunk_rows = max(1, min(max(N, 1), chunk_rows or (256 << 20) // (16 * Ln)))
            main = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(main)
            nbuf = min(2, -(-N // chunk_rows))
            stage = [torch.empty((chunk_rows, rw), dtype=torch.int32, pin_memory=True) for _ in range(nbuf)]
            done = [None] * nbuf
            total_bad = 0
            for i, s in enumerate(range(0, N, chunk_rows)):
                e, j = min(N, s + chunk_rows), i % nbuf
                if done[j] is not None:
                    done[j].synchronize()                       # the copy that read this staging buffer has finished
                _, bad = packOneHotHost(data[s:e], out=stage[j].numpy().view(np.uint32), threads=threads)
                total_bad += bad
                with torch.cuda.stream(side):
                    words[s:e].copy_(stage[j][:e - s], non_blocking=True)
                    done[j] = torch.cuda.Event()
                    done[j].record(side)
            main.wait_stream(side)
            if N:
                words.record_stream(side)
            if total_bad:
                raise ValueError("input is not one-hot at %d positions (expected exactly one 1 among A,C,G,T)" % total_bad)
            return cls(words, N, Ln)
        N, Ln = data.shape[0], data.shape[3]
        words, rw = cls._alloc(N, Ln, device)
        bad = torch.zeros(1, dtype=torch.int32, device=device)
        chunk_rows = chunk_rows or max(1, (64 << 20) // (16 * Ln))
        total_bad = 0
        for s in range(0, N, chunk_rows):
            e = min(N, s + chunk_rows)
            dev = data[s:e].to(device=device, dtype=torch.float32).contiguous()
            _lib.check(lib.crbm_pack_onehot(_lib.ptr(dev), e - s, Ln, _lib.ptr(words[s:]), rw, _lib.ptr(bad),
                                            _lib.stream_handle()), 'pack_onehot')
            total_bad += int(bad.item())
            bad.zero_()
        if total_bad:
            raise ValueError("input is not one-hot at %d positions (expected exactly one 1 among A,C,G,T)" % total_bad)
        return cls(words, N, Ln)

    @classmethod
    def from_codes(cls, codes, device='cuda'):
        """(N,L) uint8 base codes in {0,1,2,3} -> packed."""
        import torch
        from . import _lib
        lib = _lib.load()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        N, Ln = codes.shape
        words, rw = cls._alloc(N, Ln, device)
        bad = torch.zeros(1, dtype=torch.int32, device=device)
        dev = torch.from_numpy(codes).to(device)
        _lib.check(lib.crbm_pack_codes(_lib.ptr(dev), N, Ln, _lib.ptr(words), rw, _lib.ptr(bad), _lib.stream_handle()),
                   'pack_codes')
        if int(bad.item()):
            raise ValueError("base codes must be in {0,1,2,3}")
        return cls(words, N, Ln)

    @classmethod
    def from_packed_host(cls, words, L, device='cuda'):
        """Rows packed on the host (packCodesHost) -> device; 0.25 byte per base crosses PCIe."""
        import torch
        words = np.ascontiguousarray(words, dtype=np.uint32)
        if words.ndim != 2 or words.shape[1] != packedRowWords(L):
            raise ValueError("expected packed rows of %d words for L=%d, got %r" % (packedRowWords(L), L, words.shape))
        return cls(torch.from_numpy(words.view(np.int32)).to(device), words.shape[0], L)

    @classmethod
    def from_fasta(cls, filename, device='cuda', chunk_bytes=32 << 20, quiet=True, threads=0):
        """FASTA file -> packed sequences on the device, streamed (semantics of readSeqsFromFasta + seqToOneHot,
        reference sequences.py:33-52,101-117: N records skipped, case folded, KeyError on other letters).

        The file is memory-mapped and cut into blocks of about `chunk_bytes` at record starts (iterFastaPacked). The host threads of
        crbm_fasta_pack_block_host (`threads`; 0 = one per core) parse a block and write its rows, already packed to 2 bits per base,
        into one of two page-locked staging buffers; one asynchronous host->device copy per block (0.25 byte per base, on a side
        stream) puts them in place while the host parses the next block. No packing kernel runs, and neither the file nor a one-hot
        array is ever held in host memory; the device tensor is sized from the file size (an upper bound on the record count)."""
        import torch
        dev = torch.device(device)
        size = os.path.getsize(filename)
        main = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        stage, done = [None, None], [None, None]
        state = {'j': 0}

        def get_buffer(rows, rw):
            j = state['j']
            if done[j] is not None:
                done[j].synchronize()                           # the copy that read this staging buffer two blocks ago has finished
                done[j] = None
            if stage[j] is None or stage[j].shape[0] < rows or stage[j].shape[1] != rw:
                stage[j] = torch.empty((max(rows, 1), rw), dtype=torch.int32, pin_memory=True)
            return stage[j].numpy().view(np.uint32)

        words, Ln, n_done = None, None, 0
        for rows, L in iterFastaPacked(filename, chunk_bytes, quiet, threads, get_buffer):
            n, j = rows.shape[0], state['j']
            if words is None:
                Ln = L
                words = torch.empty((size // (Ln + 2) + 1, rows.shape[1]), dtype=torch.int32, device=dev)
                side.wait_stream(main)
            if n_done + n > words.shape[0]:                  # (cannot happen: a record takes at least L + 2 bytes of file)
                grown = torch.empty((max(2 * words.shape[0], n_done + n), words.shape[1]), dtype=torch.int32, device=dev)
                side.synchronize()
                grown[:n_done] = words[:n_done]
                words = grown
                side.wait_stream(main)
            with torch.cuda.stream(side):
                words[n_done:n_done + n].copy_(stage[j][:n], non_blocking=True)
                done[j] = torch.cuda.Event()
                done[j].record(side)
            n_done += n
            state['j'] = j ^ 1
        if words is None:
            raise ValueError("no usable record in %s" % filename)
        main.wait_stream(side)
        words.record_stream(side)
        return cls(words[:n_done], n_done, Ln)

    @classmethod
    def random(cls, N, L, seed=1234, device='cuda'):
        """Synthetic i.i.d. uniform bases generated directly in packed form on the device."""
        from . import _lib
        lib = _lib.load()
        words, rw = cls._alloc(N, L, device)
        _lib.check(lib.crbm_random_bases(_lib.ptr(words), rw, int(N), int(L), int(seed), _lib.stream_handle()),
                   'random_bases')
        return cls(words, N, L)

    def to_onehot(self):
        """Back to the reference's (N,1,4,L) float32 array (NumPy)."""
        import torch
        from . import _lib
        lib = _lib.load()
        out = torch.empty((self.N, 1, 4, self.L), dtype=torch.float32, device=self.words.device)
        _lib.check(lib.crbm_unpack_onehot(_lib.ptr(self.words), self.row_words, self.N, self.L, _lib.ptr(out),
                                          _lib.stream_handle()), 'unpack_onehot')
        return out.cpu().numpy()

    def rows(self, start, end):
        return PackedSequences(self.words[start:end], end - start, self.L)

    def truncated(self, new_L):
        """View of the first new_L bases of every sequence (re-packed: the padding bits must be zero)."""
        if new_L == self.L:
            return self
        import torch
        from . import _lib
        rw = _lib.load().crbm_packed_row_words(int(new_L))
        words = torch.zeros((self.N, rw), dtype=torch.int32, device=self.words.device)
        nw = 1 + (new_L + 15) // 16
        words[:, :nw] = self.words[:, :nw]
        rem = new_L % 16
        if rem:
            mask = (1 << (2 * rem)) - 1
            words[:, nw - 1] &= int(np.int32(np.uint32(mask)))
        return PackedSequences(words, self.N, new_L)
