"""Self-critical sequence training reward: ``get_self_critical_reward`` of the reference's
models/modelutils.py:200-238, which scores the sampled and the greedy captions of a batch against the ground-truth
captions with the CIDEr and BLEU scorers vendored under its pycocoevalcap/ (cider/cider_scorer.py, bleu/bleu_scorer.py)
and returns their difference, repeated over the sequence length.

Host code like the reference's (n-gram counting over a few hundred short integer sequences); the scorers here work on
the token ids directly — the reference joins the words to a string and splits it again, which is the same n-grams as
long as the word map is one-to-one.  Quirks kept:
  * ``array_to_str`` tests ``end_encode in out`` with an integer against a list of words, so it never stops at <end>:
    everything that is not <start> / <pad> counts, <end> included (modelutils.py:190-198);
  * CIDEr's "length" of a sentence is its number of BIGRAMS (``if n == 1`` on a zero-based n-gram order,
    cider_scorer.py:120-121); the document frequencies come from the batch's own references, each image counted twice
    (once for the sampled, once for the greedy caption), and the idf uses log(number of scored captions).
"""
import math
from collections import Counter

import numpy as np


def array_to_ids(arr, skip):
    """modelutils.py:190-198 on ids: every token that is not <start> / <pad>, in order (the <end> test never fires)."""
    return [int(w) for w in arr if int(w) not in skip]


def _ngrams(words, n=4):
    c = Counter()
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            c[tuple(words[i:i + k])] += 1
    return c


def cider_scores(gts, res, n=4, sigma=6.0):
    """pycocoevalcap/cider: gts = list (one per scored caption) of lists of reference id-sequences, res = list of
    hypothesis id-sequences.  -> (mean, per-caption scores)."""
    refs = [[_ngrams(r, n) for r in rs] for rs in gts]
    hyps = [_ngrams(h, n) for h in res]
    df = Counter()
    for rs in refs:
        for g in set(g for r in rs for g in r):
            df[g] += 1
    log_n = np.log(float(len(refs)))

    def vec(cnt):
        v = [dict() for _ in range(n)]
        norm = [0.0] * n
        length = 0
        for g, tf in cnt.items():
            k = len(g) - 1
            w = float(tf) * (log_n - np.log(max(1.0, df.get(g, 0.0))))
            v[k][g] = w
            norm[k] += w * w
            if k == 1:
                length += tf
        return v, [np.sqrt(x) for x in norm], length

    scores = []
    for h, rs in zip(hyps, refs):
        vh, nh, lh = vec(h)
        score = np.zeros(n)
        for r in rs:
            vr, nr, lr = vec(r)
            delta = float(lh - lr)
            val = np.zeros(n)
            for k in range(n):
                for g, w in vh[k].items():
                    wr = vr[k].get(g, 0.0)
                    val[k] += min(w, wr) * wr
                if nh[k] != 0 and nr[k] != 0:
                    val[k] /= nh[k] * nr[k]
                val[k] *= np.e ** (-(delta ** 2) / (2 * sigma ** 2))
            score += val
        scores.append(np.mean(score) / len(rs) * 10.0)
    scores = np.array(scores)
    return float(np.mean(scores)), scores


def bleu_scores(gts, res, n=4):
    """pycocoevalcap/bleu with option 'closest': -> (corpus BLEU-1..n, per-caption lists of BLEU-1..n)."""
    small, tiny = 1e-9, 1e-15
    per = [[] for _ in range(n)]
    tot_guess, tot_correct = [0] * n, [0] * n
    tot_test, tot_ref = 0, 0
    for h, rs in zip(res, gts):
        maxc = {}
        for r in rs:
            for g, c in _ngrams(r, n).items():
                maxc[g] = max(maxc.get(g, 0), c)
        testlen = len(h)
        reflen = min((abs(len(r) - testlen), len(r)) for r in rs)[1]
        guess = [max(0, testlen - k + 1) for k in range(1, n + 1)]
        correct = [0] * n
        for g, c in _ngrams(h, n).items():
            correct[len(g) - 1] += min(maxc.get(g, 0), c)
        tot_test += testlen
        tot_ref += reflen
        b = 1.0
        for k in range(n):
            tot_guess[k] += guess[k]
            tot_correct[k] += correct[k]
            b *= (float(correct[k]) + tiny) / (float(guess[k]) + small)
            per[k].append(b ** (1.0 / (k + 1)))
        ratio = (testlen + tiny) / (reflen + small)
        if ratio < 1:
            for k in range(n):
                per[k][-1] *= math.exp(1 - 1 / ratio)
    corpus, b = [], 1.0
    for k in range(n):
        b *= float(tot_correct[k] + tiny) / (tot_guess[k] + small)
        corpus.append(b ** (1.0 / (k + 1)))
    ratio = (tot_test + tiny) / (tot_ref + small)
    if ratio < 1:
        corpus = [c * math.exp(1 - 1 / ratio) for c in corpus]
    return corpus, per


def self_critical_reward(greedy_res, data_gts, gen_result, word_map, cider_reward_weight, bleu_reward_weight):
    """modelutils.py:200-238.  greedy_res / gen_result: (B, L) integer arrays (tensors are read once to the host),
    data_gts: per image a list of reference id-sequences.  -> float64 (B, L): score(sampled) - score(greedy), repeated
    along the sequence."""
    to_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    gen, greedy = to_np(gen_result), to_np(greedy_res)
    B = gen.shape[0]
    skip = {word_map['<start>'], word_map['<pad>']}
    res = [array_to_ids(gen[i], skip) for i in range(B)] + [array_to_ids(greedy[i], skip) for i in range(B)]
    refs = [[array_to_ids(to_np(c), skip) for c in data_gts[i]] for i in range(len(data_gts))]
    gts = [refs[i % B] for i in range(2 * B)]
    scores = 0
    if cider_reward_weight > 0:
        scores = scores + cider_reward_weight * cider_scores(gts, res)[1]
    if bleu_reward_weight > 0:
        scores = scores + bleu_reward_weight * np.array(bleu_scores(gts, res)[1][3])
    scores = scores[:B] - scores[B:]
    return np.repeat(scores[:, np.newaxis], gen.shape[1], 1)
