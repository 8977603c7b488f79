"""nats_b200.gen with a pool of worker processes (-p 2) on a tiny random-weight model: exercises the spawn path."""
import os, pickle, sys, tempfile
from collections import OrderedDict
import numpy as np
sys.path.insert(0, '.')
if __name__ == '__main__':
    from nats_b200 import nats as N, gen
    d = tempfile.mkdtemp()
    words = ['w%02d' % i for i in range(20)]
    wd = OrderedDict([('eos', 0), ('UNK', 1)] + [(w, i + 2) for i, w in enumerate(words)])
    pickle.dump(wd, open(os.path.join(d, 'dict.pkl'), 'wb'), protocol=2)
    rng = np.random.RandomState(0)
    with open(os.path.join(d, 'src.txt'), 'w') as f:
        for _ in range(6):
            f.write(' '.join(rng.choice(words, size=rng.randint(4, 9))) + '\n')
    opts = dict(dim_word=8, dim=16, dim_att=6, n_words=24, encoder='gru', decoder='gru_cond', use_dropout=False)
    np.random.seed(3)
    P = N.init_params(opts)
    np.savez(os.path.join(d, 'model.npz'), history_errs=[], **P)
    pickle.dump(opts, open(os.path.join(d, 'model.npz.pkl'), 'wb'), protocol=2)
    out1, out2 = os.path.join(d, 'o1.txt'), os.path.join(d, 'o2.txt')
    gen.main(os.path.join(d, 'model.npz'), os.path.join(d, 'dict.pkl'), os.path.join(d, 'src.txt'), out1, k=3, normalize=True, n_process=1)
    gen.main(os.path.join(d, 'model.npz'), os.path.join(d, 'dict.pkl'), os.path.join(d, 'src.txt'), out2, k=3, normalize=True, n_process=2)
    a, b = open(out1).read(), open(out2).read()
    assert a == b and len(a.split('\n')) >= 6, (a, b)
    print('gen pool check ok:', len(a.split('\n')), 'lines, identical with 1 and 2 workers')
