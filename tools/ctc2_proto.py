"""CPU prototype (numpy, float32) of the lane-normalised two-sided CTC sweep that csrc/ctc2.cuh
implements: validates the index maps (position groups, phantom pairs, the inject, the reversed beta
side, the Z correspondences) and the numeric scheme against the float64 oracle before any CUDA runs.

  python tools/ctc2_proto.py            # a handful of seeded cases, prints max errors

Not product code, not imported by anything.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402

f32 = np.float32
VOID_E = -(1 << 28)
TARGET = 32          # lane maximum normalised into [2^32, 2^33)
JP = 4


def pow2(d):
    """2^d as float32 with flush below -126 (exponent-field construction)."""
    d = np.asarray(d)
    return np.where(d + 127 <= 0, f32(0), np.ldexp(f32(1), np.clip(d, -126, 127)).astype(f32)).astype(f32)


class Side:
    """One sweep direction of one utterance: NL lanes of JP pairs (blank, label), lane exponent e."""

    def __init__(self, y, L, dirn):
        self.dir = dirn
        self.L = L
        self.G = (L + 3) // 4
        self.NL = self.G + 2
        NL = self.NL
        self.bm = np.zeros((NL, JP), f32)
        self.lm = np.zeros((NL, JP), f32)
        self.e = np.full((NL, JP), VOID_E, np.int64)         # one exponent per (blank, label) pair
        self.A = 4 * NL                       # a-space size
        # pair k of this side <-> label index a
        k = np.arange(self.A)
        self.a_of = k if dirn == 0 else (self.A - 1 - k)
        a_inj = 3 if dirn == 0 else L + 4
        k_inj = a_inj if dirn == 0 else self.A - 1 - a_inj
        self.lm[k_inj // 4, k_inj % 4] = 1.0
        self.e[k_inj // 4, k_inj % 4] = 0
        # allowed mask per pair
        al = np.zeros(self.A, bool)
        for kk in range(self.A):
            a = self.a_of[kk]
            p = a - 4
            if dirn == 0:
                if p == 0:
                    al[kk] = True
                elif 1 <= p < L:
                    al[kk] = (y[p] != y[p - 1]) and (y[p] != 0)
            else:
                if p == L - 1 or p == -1:     # p == -1: the last pair's phantom label only ever collects Z (T_n = 1)
                    al[kk] = True
                elif 0 <= p and p + 1 < L:
                    al[kk] = (y[p + 1] != y[p]) and (y[p + 1] != 0)
            if L == 0:
                # no labels: the first real pair is the lone blank's pair
                al[kk] = (a == 4) if dirn == 0 else (a == 3)
        self.al = al.reshape(NL, JP)

    def step(self, em):
        """em: float32 [A] emissions by a (0 for phantoms), em_blank scalar in em[-1] slot -> passed apart."""
        raise NotImplementedError


def lane_step(s, pb, pl):
    """One recursion step of side s.  pl[NL,JP]: label emission of each pair.  Returns (u, v, e_pre)."""
    NL = s.NL
    lm_flat = s.lm.reshape(-1); e_flat = s.e.reshape(-1)
    cm = np.concatenate([[f32(0)], lm_flat[:-1]]).astype(f32).reshape(NL, JP)      # label state of the pair below
    ce = np.concatenate([[VOID_E], e_flat[:-1]]).reshape(NL, JP)
    d = ce - s.e
    big = d > 30
    sh = np.where(big, d - 30, 0)
    s.bm = np.ldexp(s.bm, -np.minimum(sh, 300).astype(np.int32)).astype(f32)
    s.lm = np.ldexp(s.lm, -np.minimum(sh, 300).astype(np.int32)).astype(f32)
    s.e = s.e + sh
    d = np.where(big, 30, d)
    c = np.ldexp(cm, np.maximum(d, -300).astype(np.int32)).astype(f32)     # exponent-field shift: exact while representable
    u = (s.bm + c).astype(f32)
    v = (s.lm + np.where(s.al, u, s.bm)).astype(f32)
    e_pre = s.e.copy()
    nb = (u * f32(pb)).astype(f32)
    nl = (v * pl).astype(f32)
    mx = np.maximum(nb, nl)
    ex = (mx.view(np.int32) >> 23).astype(np.int64)
    delta = np.minimum((TARGET + 127) - ex, 120)
    f = pow2(delta)
    zero = mx == 0
    s.bm = np.where(zero, f32(0), nb * f).astype(f32)
    s.lm = np.where(zero, f32(0), nl * f).astype(f32)
    s.e = np.where(zero, VOID_E, s.e - delta)
    return u, v, e_pre


def emission(x, l2):
    """p = 2^(x log2e - l2) the way the kernel forms it (integer part split off before rounding)."""
    LOG2E = f32(1.4426950408889634)
    l2 = f32(l2)
    l2q = f32(np.rint(l2 * f32(1024)) / f32(1024))
    dl = f32(l2 - l2q)
    t = f32(np.float64(x) * np.float64(LOG2E) - np.float64(l2q))            # fma
    k = np.rint(t)
    fr = f32(f32(np.float64(x) * np.float64(LOG2E) - np.float64(f32(l2q + k))) - dl)
    p = f32(np.exp2(np.float64(fr)))
    if k < -125:
        return f32(1.1754943508222875e-38)
    return f32(np.ldexp(p, int(k)))


def run_utt(x, y, L, Tn, from_logits=True, gout=1.0):
    """x [T,V] float32 logits of one utterance -> (loss, grad [T,V]) via the prototype scheme."""
    T, V = x.shape
    LOG2E = f32(1.4426950408889634)
    if from_logits:
        m = x.max(axis=1)
        l2 = (m * LOG2E + np.log2(np.exp2((x - m[:, None]).astype(np.float64) * np.float64(LOG2E)).sum(axis=1))).astype(f32)
    else:
        l2 = np.zeros(T, f32)
    sides = [Side(y, L, 0), Side(y, L, 1)]
    A = sides[0].A
    NL = sides[0].NL

    def em_row(t):
        em = np.zeros(A, f32)
        for p in range(L):
            em[p + 4] = emission(x[t, y[p]], l2[t])
        return emission(x[t, 0], l2[t]), em

    tm = Tn // 2
    steps1 = [tm, Tn - tm]
    stored_l = np.zeros((Tn, A), f32)        # post-emission label values by a
    stored_e = np.zeros((Tn, A), np.int64)   # exponent of the pair, by a
    bound = []
    for s in sides:
        for i in range(steps1[s.dir]):
            t = (Tn - 1 - i) if s.dir else i
            pb, em = em_row(t)
            pl = em[s.a_of].reshape(NL, JP)
            lane_step(s, pb, pl)
            stored_l[t, s.a_of] = s.lm.reshape(-1)
            stored_e[t, s.a_of] = s.e.reshape(-1)
        bB = np.zeros(A, f32); lB = np.zeros(A, f32); eB = np.zeros(A, np.int64)
        bB[s.a_of] = s.bm.reshape(-1); lB[s.a_of] = s.lm.reshape(-1)
        eB[s.a_of] = s.e.reshape(-1)
        bound.append((bB, lB, eB))

    # Z in a safe way: collect (mantissa, exponent) terms and sum relative to the max exponent
    def z_terms(d):
        import copy
        s = copy.deepcopy(sides[d])
        u, v, e_pre = lane_step(s, f32(1), np.ones((NL, JP), f32))
        oB, oL, oE = bound[1 - d]
        a = s.a_of.reshape(NL, JP)
        ab = a - 1 if d == 0 else a + 1
        terms = []
        for ln in range(NL):
            for j in range(JP):
                if 0 <= ab[ln, j] < A:
                    m = float(u[ln, j]) * float(oB[ab[ln, j]])
                    if m > 0:
                        terms.append((m, int(e_pre[ln, j] + oE[ab[ln, j]])))
                m = float(v[ln, j]) * float(oL[a[ln, j]])
                if m > 0:
                    terms.append((m, int(e_pre[ln, j] + oE[a[ln, j]])))
        if not terms:
            return None
        pm = max(np.log2(m) + e for m, e in terms)
        pmi = int(np.floor(pm))
        sm = sum(m * 2.0 ** (e - pmi) for m, e in terms if e - pmi > -1000)
        return pmi, sm

    z0 = z_terms(0); z1 = z_terms(1)
    assert z0 is not None and z1 is not None
    log2z0 = z0[0] + np.log2(z0[1]); log2z1 = z1[0] + np.log2(z1[1])
    assert abs(log2z0 - log2z1) < 1e-4 * max(1.0, abs(log2z0)), (log2z0, log2z1)
    pm, sm = z0
    ex = int(np.floor(np.log2(sm)))
    rZ = f32(1.0 / (sm / 2.0 ** ex)); eZ = pm + ex
    loss = -(log2z0) * np.log(2.0)

    # phase 2
    occ = np.zeros((Tn, A), f32)
    for s in sides:
        for i in range(steps1[s.dir], Tn):
            t = (Tn - 1 - i) if s.dir else i
            pb, em = em_row(t)
            pl = em[s.a_of].reshape(NL, JP)
            u, v, e_pre = lane_step(s, pb, pl)
            a = s.a_of.reshape(NL, JP)
            ob = stored_l[t][a]
            eb = stored_e[t][a]
            xexp = e_pre + eb - eZ
            sc = (rZ * pow2(np.clip(xexp, -127, 90))).astype(f32)
            g = ((v * ob).astype(f32) * sc).astype(f32)
            occ[t, a.reshape(-1)] = g.reshape(-1)
    grad = np.zeros((T, V), np.float64)
    for t in range(Tn):
        if from_logits:
            grad[t] = np.exp2(x[t].astype(np.float64) * 1.4426950408889634 - l2[t])
        bs = 0.0
        for p in range(L):
            grad[t, y[p]] -= occ[t, p + 4]
            bs += occ[t, p + 4]
        grad[t, 0] -= 1.0 - bs
    return loss, grad * gout


def main():
    rng = np.random.default_rng(0)
    cases = [(30, 8, 5, 1.0), (64, 32, 9, 1.0), (50, 16, 12, 1.0), (40, 16, 0, 1.0), (1, 8, 1, 1.0), (2, 8, 1, 1.0),
             (33, 12, 16, 1.0), (90, 40, 17, 3.0), (120, 24, 40, 5.0), (7, 6, 3, 1.0), (25, 5, 11, 1.0)]
    for (T, V, L, scale) in cases:
        x = (rng.standard_normal((T, 1, V)) * scale).astype(f32)
        y = rng.integers(1, V, size=(1, max(L, 1)))
        if L >= 4:
            y[0, 2] = y[0, 1]       # a repeat
        if L >= 6:
            y[0, 4] = 0             # label 0 inside the target (the dst != 0 quirk)
        ol, og = oracle.ctc(x, y, np.array([T]), np.array([L]))
        loss, grad = run_utt(x[:, 0, :], y[0], L, T)
        if not np.isfinite(ol[0]):
            print(f"T={T} V={V} L={L}: infeasible in the oracle, skipped")
            continue
        print(f"T={T:4d} V={V:3d} L={L:3d} x{scale}: loss {loss:.6f} vs {ol[0]:.6f} rel {abs(loss / ol[0] - 1):.2e}  "
              f"grad max err {np.abs(grad - og[:, 0, :]).max():.2e}")


if __name__ == "__main__":
    main()
