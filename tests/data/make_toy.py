"""Builds tests/data/toy/: a small, self-contained cut of the reference's toy corpus (data/toy_{train,validation,test}_
{input,output}.txt: CNN-style article -> highlight pairs) for the config-1 plumbing run (BASELINE.json configs[0]) on
the GPU box, where /root/reference does not exist.  Sources are cut to their first 120 tokens, targets to 30, 128 / 16 /
16 pairs; the dictionary is built from the cut training source with the python-3 twin of data/build_dictionary.py.

    python tests/data/make_toy.py          (from the repo root, with /root/reference present)
"""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nats_b200 import build_dictionary  # noqa: E402

SRC = '/root/reference/data'
DST = os.path.join(ROOT, 'tests', 'data', 'toy')


def cut(name_in, name_out, n_pairs, n_tok):
    with open(os.path.join(SRC, name_in)) as f:
        lines = [' '.join(l.split()[:n_tok]) for l in f][:n_pairs]
    with open(os.path.join(DST, name_out), 'w') as f:
        f.write('\n'.join(lines) + '\n')


def main():
    os.makedirs(DST, exist_ok=True)
    for split, n in (('train', 128), ('validation', 16), ('test', 16)):
        cut('toy_%s_input.txt' % split, '%s_input.txt' % split, n, 120)
        cut('toy_%s_output.txt' % split, '%s_output.txt' % split, n, 30)
    path = os.path.join(DST, 'train_input.txt')
    d = build_dictionary.build(path)
    with open(path + '.pkl', 'wb') as f:
        pickle.dump(d, f, protocol=2)
    print('toy corpus written to', DST, '- dictionary of', len(d), 'words')


if __name__ == '__main__':
    main()
