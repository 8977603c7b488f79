"""Line-aligned bitext iterator with the behaviour of the reference's scripts/data_iterator.py:11-80
(word -> id through a pickled dictionary, UNK = 1, ids >= n_words -> 1, batches of `batch_size` pairs,
StopIteration at the end of the data followed by an automatic rewind).  Host-side, python 3."""
import gzip
import pickle as pkl


def fopen(filename, mode='r'):
    if filename.endswith('.gz'):
        return gzip.open(filename, mode + 't' if 't' not in mode and 'b' not in mode else mode)
    return open(filename, mode)


def load_dictionary(path):
    with open(path, 'rb') as f:
        try:
            return pkl.load(f)
        except UnicodeDecodeError:      # python-2 pickle produced by the reference's build_dictionary.py
            f.seek(0)
            return pkl.load(f, encoding='latin1')


class TextIterator(object):
    """`bucket_batches = k > 0` (not in the reference, default off) reads k batches at a time, sorts the pairs by
    source length and cuts the batches from the sorted pool: same pairs per epoch, far less padding per batch (every
    padded source position costs a full encoder step on the GPU).  With 0 the order is the file order, as the reference."""

    def __init__(self, source, target, dict, batch_size=128, n_words=-1, bucket_batches=0):
        self.source = fopen(source, 'r')
        self.target = fopen(target, 'r')
        self.dict = load_dictionary(dict)
        self.batch_size = batch_size
        self.n_words = n_words
        self.end_of_data = False
        self.bucket_batches = int(bucket_batches)
        self._pool = []

    def __iter__(self):
        return self

    def reset(self):
        self.source.seek(0)
        self.target.seek(0)

    def _ids(self, line):
        ids = [self.dict.get(w, 1) for w in line.strip().split()]
        if self.n_words > 0:
            ids = [w if w < self.n_words else 1 for w in ids]
        return ids

    def __next__(self):
        if self.bucket_batches > 0:
            return self._next_bucketed()
        if self.end_of_data:
            self.end_of_data = False
            self.reset()
            raise StopIteration
        source, target = [], []
        while len(source) < self.batch_size:
            ss = self.source.readline()
            tt = self.target.readline() if ss != '' else ''
            if ss == '' or tt == '':
                self.end_of_data = True
                break
            source.append(self._ids(ss))
            target.append(self._ids(tt))
        if len(source) <= 0:
            self.end_of_data = False
            self.reset()
            raise StopIteration
        return source, target

    def _next_bucketed(self):
        if not self._pool:
            if self.end_of_data:
                self.end_of_data = False
                self.reset()
                raise StopIteration
            pairs = []
            while len(pairs) < self.batch_size * self.bucket_batches:
                ss = self.source.readline()
                tt = self.target.readline() if ss != '' else ''
                if ss == '' or tt == '':
                    self.end_of_data = True
                    break
                pairs.append((self._ids(ss), self._ids(tt)))
            if not pairs:
                self.end_of_data = False
                self.reset()
                raise StopIteration
            pairs.sort(key=lambda p: (len(p[0]), len(p[1])))         # stable: ties keep the file order
            self._pool = [pairs[i:i + self.batch_size] for i in range(0, len(pairs), self.batch_size)]
        batch = self._pool.pop(0)
        return [p[0] for p in batch], [p[1] for p in batch]

    next = __next__
