"""CPU check of the fused star-CTC lane arithmetic (csrc/star2_math.h) against the float64 oracle.

Builds tools/star2_host_check.cpp (the shipped lane functions compiled for the host, stepped lane by lane) and
compares loss / gradient with oracle.star on random cases: repeats, label 0, L = 0, L = S, T = 1, peaky logits,
log-prob input, penalties.  Test infrastructure: runs here without a GPU."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

SO = "/tmp/libstar2_host.so"


def build():
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", SO, os.path.join(ROOT, "tools", "star2_host_check.cpp")])
    lib = ctypes.CDLL(SO)
    lib.star2_host.restype = ctypes.c_int
    return lib


def run(lib, x, y, L, Tn, pen, from_logits, gout=1.0, J=4):
    T, V = x.shape
    S = len(y)
    x = np.ascontiguousarray(x, np.float32)
    y32 = np.ascontiguousarray(y, np.int32)
    loss = ctypes.c_float()
    grad = np.zeros((T, V), np.float32)
    lib.star2_host(x.ctypes.data_as(ctypes.c_void_p), int(T), int(V), y32.ctypes.data_as(ctypes.c_void_p), int(S), int(L), int(Tn),
                   ctypes.c_float(pen), int(from_logits), ctypes.c_float(gout), ctypes.byref(loss),
                   grad.ctypes.data_as(ctypes.c_void_p), int(J))
    return loss.value, grad


def main():
    lib = build()
    rng = np.random.default_rng(0)
    worst_l = worst_g = 0.0
    cases = []
    for T, V, S, L, scale in [(1, 8, 3, 0, 1), (1, 8, 3, 1, 1), (2, 8, 3, 1, 1), (5, 8, 4, 4, 1), (30, 16, 7, 7, 1), (30, 16, 9, 5, 1),
                              (64, 32, 12, 12, 3), (100, 12, 40, 37, 1), (200, 64, 50, 50, 1), (300, 20, 9, 0, 1),
                              (400, 48, 130, 129, 2), (150, 8, 20, 20, 6), (500, 512, 200, 200, 1), (3000, 8, 4, 4, 1)]:
        for rep in range(3):
            cases.append((T, V, S, L, scale, rep))
    for (T, V, S, L, scale, rep) in cases:
        x = (rng.standard_normal((T, V)) * scale).astype(np.float32)
        y = rng.integers(0 if rep == 2 else 1, min(V, 4 if rep == 1 else V), S)       # rep 1: many repeats; rep 2: label 0 too
        Tn = T if rep != 1 else max(1, T - rng.integers(0, max(1, T // 3)))
        pen = [-0.5, 0.0, -2.0][rep]
        fl = rep != 2 or T > 100
        xin = x if fl else (x - np.log(np.exp(x.astype(np.float64)).sum(-1, keepdims=True))).astype(np.float32)
        lo, go = oracle.star(xin[:, None, :], y[None], np.array([Tn]), np.array([L]), star_penalty=pen, from_logits=fl)
        J = (4, 2, 1)[len(str(T * 7 + S + L)) % 3] if T < 400 else (4, 2, 1)[rep]
        lh, gh = run(lib, xin, y, L, Tn, pen, fl, J=J)
        if not np.isfinite(lo[0]):
            ok = not np.isfinite(lh)
            print(f"T={T} V={V} S={S} L={L} Tn={Tn}: infeasible, host {'agrees' if ok else 'DISAGREES'}")
            continue
        dl = abs(lh / lo[0] - 1) if lo[0] != 0 else abs(lh)
        dg = np.abs(gh - go[:, 0, :]).max()
        worst_l = max(worst_l, dl); worst_g = max(worst_g, dg)
        flag = "" if (dl < 1e-4 and dg < 1e-5) else "   <-- FAIL"
        print(f"T={T} V={V} S={S} L={L} Tn={Tn} x{scale} pen={pen} logits={int(fl)} J={J}: loss {lo[0]:.6f} rel {dl:.1e}  grad abs {dg:.1e}{flag}")
    print(f"worst: loss rel {worst_l:.2e}, grad abs {worst_g:.2e}")
    return 0 if (worst_l < 1e-4 and worst_g < 1e-5) else 1


if __name__ == "__main__":
    sys.exit(main())
