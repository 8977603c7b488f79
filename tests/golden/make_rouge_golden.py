"""Generates tests/golden/rouge_cases.json by running the REFERENCE's own scripts/ROUGE.pl (perl is in the build
container) on seeded pseudo-sentences, including the corner cases of its tokeniser (leading / repeated blanks, empty
lines, fewer system lines than reference lines).  Run from the repo root with /root/reference present:

    python tests/golden/make_rouge_golden.py
"""
import json
import os
import random
import subprocess
import tempfile

REF = '/root/reference/scripts/ROUGE.pl'
HERE = os.path.dirname(os.path.abspath(__file__))


def sentences(rng, n, vocab, lo, hi):
    return [' '.join(rng.choice(vocab) for _ in range(rng.randint(lo, hi))) for _ in range(n)]


def corrupt(rng, line, vocab):
    toks = line.split()
    out = []
    for t in toks:
        r = rng.random()
        if r < 0.15:
            continue
        out.append(rng.choice(vocab) if r < 0.35 else t)
        if rng.random() < 0.1:
            out.append(rng.choice(vocab))
    return ' '.join(out)


def main():
    rng = random.Random(20260921)
    vocab = ['w%d' % i for i in range(40)] + ['the', 'a', 'of', '.', ',']
    cases = []
    for ci in range(6):
        refs = sentences(rng, 12, vocab, 3, 30)
        sys_ = [corrupt(rng, r, vocab) for r in refs]
        if ci == 1:
            sys_[2] = ''                       # empty system line
            refs[5] = ''                       # empty reference line
        if ci == 2:
            sys_[0] = '  ' + sys_[0]           # leading blanks -> an empty first token in perl's split
            refs[1] = refs[1].replace(' ', '   ', 2)
            sys_[3] = sys_[3] + '   '
        if ci == 3:
            sys_ = sys_[:7]                    # fewer system lines
        if ci == 4:
            sys_ = list(refs)                  # identical
        if ci == 5:
            sys_ = [' '.join(reversed(r.split())) for r in refs]
        case = {'ref': refs, 'sys': sys_, 'scores': {}}
        with tempfile.TemporaryDirectory() as d:
            rp, sp = os.path.join(d, 'ref.txt'), os.path.join(d, 'sys.txt')
            open(rp, 'w').write('\n'.join(refs) + '\n')
            open(sp, 'w').write('\n'.join(sys_) + '\n')
            for n, metric in ((1, 'N'), (2, 'N'), (3, 'N'), (1, 'L')):
                out = subprocess.run(['perl', REF, str(n), metric, rp, sp], capture_output=True, text=True, check=True).stdout
                case['scores']['%d%s' % (n, metric)] = out
        cases.append(case)
    with open(os.path.join(HERE, 'rouge_cases.json'), 'w') as f:
        json.dump(cases, f, indent=0)
    print('wrote', len(cases), 'cases')


if __name__ == '__main__':
    main()
