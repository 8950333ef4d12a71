class _Seq(object):
    def __init__(self, data, alphabet):
        self._data = data
        self.alphabet = alphabet

    def __len__(self):
        return len(self._data)

    def __getitem__(self, i):
        return self._data[i]

    def __str__(self):
        return self._data


class _Record(object):
    def __init__(self, name, seq):
        self.id = self.name = self.description = name
        self.seq = seq


def parse(handle, fmt, alphabet=None):
    name, chunks = None, []
    for line in handle:
        line = line.strip()
        if not line:
            continue
        if line.startswith('>'):
            if name is not None:
                yield _Record(name, _Seq(''.join(chunks), alphabet))
            name, chunks = line[1:], []
        else:
            chunks.append(line)
    if name is not None:
        yield _Record(name, _Seq(''.join(chunks), alphabet))


def write(records, filename, fmt):
    with open(filename, 'w') as f:
        for r in records:
            f.write('>%s\n%s\n' % (r.id, str(r.seq)))
