"""Alignment paths: greedy argmax decode (ha/recognizer.py:48-59) and CTC Viterbi forced alignment."""
import torch

from . import ops


def greedy_decode(log_probs, input_lengths=None):
    """TemporalClassifier.decode's hot part (ha/recognizer.py:51-57) on the GPU.

    log_probs (N,T,C).  Returns (hypotheses, output_lengths, alignments, scores) where alignments
    and scores are torch.max(dim=-1) bit for bit (first index wins ties), hypotheses is the
    collapsed (unique_consecutive, blanks dropped) label sequence padded with -1 to (N,T) and
    output_lengths its lengths.  input_lengths=None reproduces the reference, which ignores lengths
    (ha/recognizer.py:51); pass them to stop the collapse at each utterance's end.
    """
    ali, sc, hyp, hl = ops.greedy_decode(log_probs, input_lengths)
    return hyp, hl, ali, sc


def greedy_decode_nested(log_probs, input_lengths=None):
    """Same, with hypotheses as the nested tensor the reference returns (ha/recognizer.py:52-55)."""
    hyp, hl, ali, sc = greedy_decode(log_probs, input_lengths)
    lens = hl.tolist()
    nested = torch.nested.nested_tensor([hyp[i, :n] for i, n in enumerate(lens)])
    return nested, hl, ali, sc, None


def ctc_beam_search_decode_logits(emit_logits, beam_size=3, input_lengths=None, graves=False):
    """ha/beam.py:71-137 on the GPU.  `emit_logits` (T,K), as the reference takes it, returns (top_seqs, seq_logits)
    with top_seqs a list of `beam_size` python lists, best first; a batch (N,T,K) returns (hyp (N,beam,T) padded
    with -1, hyp_len (N,beam), seq_logits (N,beam)) and honours `input_lengths`.

    graves=False reproduces the reference function as written, including the blank score 0.0 (= log 1) it gives
    every extension candidate (ha/beam.py:124), which makes every hypothesis grow by one symbol per frame once
    T is more than a few frames; graves=True gives extensions blank score -inf, i.e. the reference's
    probability-domain twin (ha/beam.py:4-68) in the log domain - the decoder one actually wants."""
    if emit_logits.dim() == 2:
        hyp, hl, sc = ops.ctc_beam_search(emit_logits[None], None, beam_size, not graves)
        lens = hl[0].tolist()
        nb = int((sc[0] > float("-inf")).sum()) if graves else len(lens)
        return [hyp[0, b, :n].tolist() for b, n in enumerate(lens)][:max(nb, 1)], sc[0, :max(nb, 1)]
    return ops.ctc_beam_search(emit_logits, input_lengths, beam_size, not graves)


def ctc_viterbi_align(log_probs, targets, input_lengths, target_lengths):
    """Best CTC alignment (max-semiring of ha/ctc.py:144-167; not in the reference).

    log_probs (T,N,C) -> (alignment (N,T) int64 class per frame, -1 beyond the input length;
    score (N,) log-prob of that path)."""
    return ops.ctc_viterbi(log_probs, targets, input_lengths, target_lengths)
