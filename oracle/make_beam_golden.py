"""Generate tests/golden/beam_*.npz by running the UNMODIFIED reference prefix beam search
(ha/beam.py:71-137 ctc_beam_search_decode_logits) on seeded log-probs.

Run in the build container only (it needs /root/reference):
    python oracle/make_beam_golden.py
The reference is evaluated in float64 and in float32 (its default); a case is kept only when both agree on every
hypothesis and consecutive scores are at least 2e-6 apart (scores are O(0.1): ~100 float32 ulps), so that the stored hypotheses do not hinge on a tie
(torch.topk's order among equal scores is unspecified).  That bounds T: the reference gives every extension
candidate a blank score of 0.0 = log 1 (ha/beam.py:124), so after a dozen frames all surviving scores collapse
onto the same float.  Scores stored are the float64 ones.

The "graves_*" cases pin the other mode (extension blank score -inf) with the reference's probability-domain twin
ctc_beam_search_decode_probs (ha/beam.py:4-68), run unmodified on exp(log-probs) with the module global `device` it
reads (ha/beam.py:46) defined; it computes in float32 probabilities, so T is kept short of underflow.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import ha.beam  # noqa: E402
from ha.beam import ctc_beam_search_decode_logits, ctc_beam_search_decode_probs  # noqa: E402

ha.beam.device = "cpu"    # the global ctc_beam_search_decode_probs reads (ha/beam.py:46)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name, T, K, beam sizes, scale of the logits, number of utterances
CASES = [
    ("beam_small", 12, 6, (1, 2, 3, 5), 1.0, 4),
    ("beam_short", 6, 20, (3, 8, 16), 1.0, 3),
    ("beam_peaky", 10, 12, (3, 4, 12), 4.0, 3),
    ("beam_blanky", 10, 9, (3, 7), 1.0, 3),          # blank-dominated frames, as a trained CTC model emits
    ("graves_small", 12, 6, (1, 2, 3, 5), 1.0, 4),
    ("graves_medium", 20, 20, (3, 8, 16), 1.0, 3),
    ("graves_peaky", 24, 12, (3, 4, 12), 4.0, 3),
    ("graves_blanky", 30, 9, (3, 7), 1.0, 3),
    ("graves_wide", 10, 300, (3, 16), 1.0, 2),
]
MAX_TRIES = 200


def run(name, lp32, b):
    """-> (seqs, scores as float64 log) or None when the case hinges on a tie"""
    if name.startswith("graves"):
        s, v = ctc_beam_search_decode_probs(lp32.exp(), beam_size=b)
        v = v.double()
        keep = int((v > 0).sum())
        if keep < len(v) or (len(v) > 1 and float((v[:-1] / v[1:]).min()) < 1 + 1e-3):
            return None
        return s, v.log()
    s64, v64 = ctc_beam_search_decode_logits(lp32.double(), beam_size=b, dtype=torch.float64)
    s32, _ = ctc_beam_search_decode_logits(lp32, beam_size=b)
    if s64 != s32 or (len(v64) > 1 and float((v64[:-1] - v64[1:]).min()) < 2e-6):
        return None
    return s64, v64


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, T, K, beams, scale, n in CASES:
        d = {"beams": np.asarray(beams)}
        seed, kept = 0, 0
        while kept < n and seed < MAX_TRIES:
            seed += 1
            g = torch.Generator().manual_seed(1000 * len(name) + seed)
            x = torch.randn(T, K, generator=g) * scale
            if name.endswith("blanky"):
                x[:, 0] += 3.0 * (torch.rand(T, generator=g) < 0.7)
            lp32 = x.log_softmax(-1)
            res = {}
            for b in beams:
                r = run(name, lp32, b)
                if r is None:
                    break
                hyp = -np.ones((b, T), dtype=np.int64)
                for i, s in enumerate(r[0]):
                    hyp[i, :len(s)] = s
                res[b] = (hyp, np.asarray([len(s) for s in r[0]]), r[1].numpy())
            if len(res) < len(beams):
                continue
            d[f"lp_{kept}"] = lp32.numpy()
            for b, (hyp, hl, v) in res.items():
                d[f"hyp_{kept}_{b}"], d[f"len_{kept}_{b}"], d[f"score_{kept}_{b}"] = hyp, hl, v
            kept += 1
        d["n"] = kept
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "seeds tried", seed, "kept", kept, flush=True)


if __name__ == "__main__":
    main()
