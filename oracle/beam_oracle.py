"""CPU restatement of the reference's CTC prefix beam search, ha/beam.py:71-137 (TEST INFRASTRUCTURE ONLY).

Same control flow as the reference (in-place per-beam update ha/beam.py:99-111, first-match prefix lookup
ha/beam.py:104, un-merged duplicates, candidates = beams unchanged then (beam, class) extensions ha/beam.py:121-126),
in numpy float64 or float32, with the one thing the reference leaves open made explicit: among equal scores the
lower candidate index wins (torch.topk does not specify it).  ext_blank = 0.0 is the reference's log-domain function
(ha/beam.py:124: torch.zeros), ext_blank = -inf its probability-domain twin (ha/beam.py:56: zeros as probabilities).
Pinned by tests/golden/beam_*.npz and graves_*.npz (tests/test_oracle_golden.py), which oracle/make_beam_golden.py
wrote by running the unmodified reference.
"""
import numpy as np


def ctc_beam_search(lp, beam_size=3, ext_blank=0.0, dtype=np.float64):
    lp = np.asarray(lp, dtype=dtype)
    T, K = lp.shape
    seqs = [[]]
    seq = np.zeros(1, dtype)
    blank = np.zeros(1, dtype)
    label = np.full(1, -np.inf, dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        for t in range(T):
            e = lp[t]
            nb = len(seqs)
            ext = np.empty((nb, K), dtype)
            for s, sq in enumerate(seqs):
                if sq:
                    label[s] = label[s] + e[sq[-1]]
                    try:
                        q = seqs.index(sq[:-1])
                    except ValueError:
                        pass
                    else:
                        label[s] = np.logaddexp(label[s], e[sq[-1]] + blank[q])
                blank[s] = seq[s] + e[0]
                last = sq[-1] if sq else 0
                base = np.full(K, seq[s], dtype)
                base[last] = blank[s]
                ext[s] = e + base
            cb = np.concatenate([blank, np.full(nb * K, ext_blank, dtype)])
            cl = np.concatenate([label, ext.reshape(-1)])
            cs = np.logaddexp(cb, cl)
            order = np.lexsort((np.arange(len(cs)), -cs))[:min(beam_size, len(cs))]
            new = []
            for i in order:
                if i < nb:
                    new.append(seqs[i])
                else:
                    s, k = divmod(i - nb, K)
                    new.append(seqs[s] + [int(k)])
            seqs, seq, blank, label = new, cs[order].copy(), cb[order].copy(), cl[order].copy()
    return seqs, seq
