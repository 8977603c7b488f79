"""Python-3 twin of the reference's data/build_dictionary.py:9-35: word frequencies in first-seen order, sorted by
ascending count (numpy.argsort) and reversed, ids from 2 (`eos` = 0, `UNK` = 1), pickled next to the text file as
`<file>.pkl` (protocol 2, readable by the python-2 reference as well).

    python -m nats_b200.build_dictionary corpus.txt [more.txt ...]
"""
import pickle as pkl
import sys
from collections import OrderedDict

import numpy


def build(filename):
    word_freqs = OrderedDict()
    with open(filename, 'r') as f:
        for line in f:
            for w in line.strip().split(' '):
                if w not in word_freqs:
                    word_freqs[w] = 0
                word_freqs[w] += 1
    words = list(word_freqs.keys())
    freqs = list(word_freqs.values())
    sorted_idx = numpy.argsort(freqs)
    sorted_words = [words[ii] for ii in sorted_idx[::-1]]
    worddict = OrderedDict()
    worddict['eos'] = 0
    worddict['UNK'] = 1
    for ii, ww in enumerate(sorted_words):
        worddict[ww] = ii + 2
    return worddict


def main(argv=None):
    for filename in (sys.argv[1:] if argv is None else argv):
        print('Processing', filename)
        worddict = build(filename)
        with open('%s.pkl' % filename, 'wb') as f:
            pkl.dump(worddict, f, protocol=2)
        print('Done')


if __name__ == '__main__':
    main()
