"""The C ABI on its own: no torch anywhere on the call path.  Buffers come from cudaMalloc (cuda-python), the
library is bound with ctypes from the declarations of include/ha_b200.h -- what a cgo / JNI / N-API binding of
INTEGRATION.md section 3 would do -- and the results are checked against the oracle."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rt():
    return pytest.importorskip("cuda.bindings.runtime")


class _Dev:
    def __init__(self, rt, nbytes):
        self.rt = rt
        err, self.ptr = rt.cudaMalloc(max(nbytes, 16))
        assert int(err) == 0, err

    def put(self, a):
        a = np.ascontiguousarray(a)
        (err,) = self.rt.cudaMemcpy(self.ptr, a.ctypes.data, a.nbytes, self.rt.cudaMemcpyKind.cudaMemcpyHostToDevice)
        assert int(err) == 0, err
        return self

    def get(self, shape, dtype):
        out = np.empty(shape, dtype)
        (err,) = self.rt.cudaMemcpy(out.ctypes.data, self.ptr, out.nbytes, self.rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        assert int(err) == 0, err
        return out

    def free(self):
        self.rt.cudaFree(self.ptr)


def test_ctc_through_the_c_abi_only():
    rt = _rt()
    from haloop_b200 import _lib              # path + signature table only (ctypes); no tensor library involved
    from oracle import oracle
    L = _lib.lib()
    rng = np.random.default_rng(3)
    T, N, V, S = 90, 5, 40, 17
    x = rng.standard_normal((T, N, V)).astype(np.float32)
    tg = rng.integers(1, V, (N, S)).astype(np.int64)
    il = np.array([90, 80, 61, 90, 45], np.int64); tl = np.array([17, 9, 1, 0, 12], np.int64)
    go = np.linspace(0.5, 1.5, N).astype(np.float32)
    nbytes = L.ha_ctc_workspace_bytes(T, N, V, S)
    assert nbytes > 0
    bufs = dict(x=_Dev(rt, x.nbytes).put(x), tg=_Dev(rt, tg.nbytes).put(tg), il=_Dev(rt, il.nbytes).put(il),
                tl=_Dev(rt, tl.nbytes).put(tl), go=_Dev(rt, go.nbytes).put(go), loss=_Dev(rt, 4 * N),
                gx=_Dev(rt, x.nbytes), ws=_Dev(rt, nbytes))
    try:
        rc = L.ha_ctc_fwd(bufs["x"].ptr, N * V, V, T, N, V, bufs["tg"].ptr, S, S, 1, bufs["il"].ptr, bufs["tl"].ptr, 1,
                          1, bufs["loss"].ptr, bufs["ws"].ptr, nbytes, None)           # NULL = the default stream
        assert rc == 0, L.ha_b200_last_error()
        rc = L.ha_ctc_bwd(bufs["x"].ptr, N * V, V, T, N, V, S, bufs["go"].ptr, 1, bufs["gx"].ptr, N * V, V,
                          bufs["ws"].ptr, nbytes, None)
        assert rc == 0, L.ha_b200_last_error()
        (err,) = rt.cudaDeviceSynchronize()
        assert int(err) == 0, err
        loss = bufs["loss"].get((N,), np.float32)
        gx = bufs["gx"].get((T, N, V), np.float32)
    finally:
        for b in bufs.values():
            b.free()
    ol, og = oracle.ctc(x, tg, il, tl, grad_out=go)
    np.testing.assert_allclose(loss, ol, rtol=1e-4)
    assert np.abs(gx - og).max() < 1e-5
    # error conventions: a too-small workspace is refused with a code and a message, nothing is launched
    rc = L.ha_ctc_fwd(1, N * V, V, T, N, V, 1, S, S, 1, 1, 1, 1, 1, 1, 16, 8, None)
    assert rc != 0 and L.ha_b200_last_error()
