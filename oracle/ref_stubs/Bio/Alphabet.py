class _Alphabet(object):
    def __init__(self, letters):
        self.letters = letters


class IUPAC(object):
    unambiguous_dna = _Alphabet("GATC")
