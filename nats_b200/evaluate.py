"""Evaluation chain of the reference's scripts/test.sh:18-26 -- gen -> replace_unk -> ROUGE -- as python-3 modules.

  replace_unk(input, origin, new)   restates scripts/replace_unk.py:25-48: the generated file holds "word [pos]" pairs
                                    (gen.py:88-98); UNK is replaced by the source word at the aligned position, <EOS> is
                                    dropped, everything else is copied.
  rouge_n / rouge_l / rouge_file    restate scripts/ROUGE.pl (ROUGE-N with clipped n-gram hits :72-138,181-232 and
                                    ROUGE-L from the longest common subsequence :141-179), INCLUDING its rounding: every
                                    per-line recall / precision / F is cut to 5 decimals (sprintf "%7.5f") before the
                                    average, which is again cut to 5 decimals and printed with 3.  alpha = 0.5 (:6).
The perl script itself runs in the build container (perl is present): tests/golden/rouge_cases.json was produced by it
(tests/golden/make_rouge_golden.py) and pins this restatement; on the GPU box only this module is available.

    python -m nats_b200.evaluate rouge 1 N reference.txt generated.txt
    python -m nats_b200.evaluate replace_unk source.txt generated_with_positions.txt final.txt
"""
import re
import sys

ALPHA = 0.5


def _r5(v):
    """sprintf("%7.5f", v) read back as a number, as the perl script does at every stage"""
    return float('%7.5f' % v)


def _split(text):
    """perl split(/\\s+/, $text): a leading blank yields an empty first field, trailing empty fields are dropped"""
    if text == '':
        return []
    toks = re.split(r'\s+', text)
    while toks and toks[-1] == '':
        toks.pop()
    return toks


def _ngrams(text, n):
    toks = _split(text)
    grams = {}
    count = 0
    for i in range(len(toks) - n + 1):
        g = ' '.join(toks[i:i + n])
        grams[g] = grams.get(g, 0) + 1
        count += 1
    return grams, count


def _prf(hit, n_model, n_peer):
    r = _r5(hit / n_model) if n_model != 0 else _r5(0)
    p = _r5(hit / n_peer) if n_peer != 0 else _r5(0)
    den = (1 - ALPHA) * p + ALPHA * r
    f = _r5((p * r) / den) if den > 0 else _r5(0)
    return r, p, f


def rouge_n(model_line, peer_line, n):
    """(recall, precision, F) of one reference / system line pair (ROUGE.pl computeNGramScore)"""
    mg, mc = _ngrams(model_line, n)
    pg, pc = _ngrams(peer_line, n)
    hit = sum(min(c, pg[g]) for g, c in mg.items() if g in pg)
    return _prf(hit, mc, pc)


def _lcs(a, b):
    m, n = len(a), len(b)
    prev = [0] * (n + 1)
    for i in range(1, m + 1):
        cur = [0] * (n + 1)
        ai = a[i - 1]
        for j in range(1, n + 1):
            if ai == b[j - 1]:
                cur[j] = prev[j - 1] + 1
            else:
                cur[j] = prev[j] if prev[j] >= cur[j - 1] else cur[j - 1]
        prev = cur
    return prev[n]


def rouge_l(model_line, peer_line):
    """(recall, precision, F) from the longest common subsequence (ROUGE.pl computeLCSScore / lcs_inner)"""
    a, b = _split(model_line), _split(peer_line)
    if len(a) == 0:                    # lcs_inner returns an empty list: all three counts undefined -> scores 0
        return _r5(0), _r5(0), _r5(0)
    return _prf(_lcs(a, b), len(a), len(b))


def rouge_lines(model_lines, peer_lines, n, metric):
    rs, ps, fs = [], [], []
    for ml, pl in zip(model_lines, peer_lines):            # stops at the shorter file, as the perl while-loop does
        ml, pl = ml.rstrip('\n'), pl.rstrip('\n')
        r, p, f = rouge_n(ml, pl, n) if metric == 'N' else rouge_l(ml, pl)
        rs.append(r); ps.append(p); fs.append(f)
    k = len(rs)
    if k == 0:
        raise ZeroDivisionError('no line pairs')           # perl: Illegal division by zero
    return _r5(sum(rs) / k), _r5(sum(ps) / k), _r5(sum(fs) / k)


def rouge_file(n, metric, model_path, peer_path):
    with open(model_path) as fm, open(peer_path) as fp:
        return rouge_lines(fm.readlines(), fp.readlines(), int(n), metric)


def format_report(n, metric, scores):
    head = 'ROUGE-%s\n' % n if metric == 'N' else 'ROUGE-L\n'
    return head + 'Ave_R | Ave_P | Ave_F\n' + '%.3f\t%.3f\t%.3f\n\n' % scores


def replace_unk(corpus, summary, new_summary, extractive=0, remove_eos=1):
    """scripts/replace_unk.py:25-48"""
    with open(corpus, 'r') as f:
        all_words = [line.strip().split() for line in f]
    with open(new_summary, 'w') as fo, open(summary, 'r') as f:
        for line, words in zip(f, all_words):
            wp = line.strip().split()
            ys = wp[::2]
            pos = [int(re.sub(r'\[|\]', '', p)) for p in wp[1::2]]
            out = []
            for a, b in zip(ys, pos):
                if remove_eos and a == '<EOS>':
                    continue
                if not extractive:
                    if a == 'UNK' and b < len(words):
                        if words[b] == '<EOS>':
                            continue
                        out.append(words[b])
                    else:
                        out.append(a)
                else:
                    out.append(a)
            # the python-2 `print >>fo, w,` idiom separates the words by one blank and ends the line with a newline
            fo.write(' '.join(out) + '\n')


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == 'rouge' and len(argv) == 5:
        n, metric = argv[1], argv[2]
        sys.stdout.write(format_report(n, metric, rouge_file(n, metric, argv[3], argv[4])))
        return 0
    if argv and argv[0] == 'replace_unk' and len(argv) == 4:
        replace_unk(argv[1], argv[2], argv[3])
        return 0
    sys.stderr.write(__doc__)
    return 2


if __name__ == '__main__':
    sys.exit(main())
