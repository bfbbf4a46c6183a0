"""ctypes access to the oracle's compiled pieces.  TEST INFRASTRUCTURE ONLY.

  port  = oracle/liboracle_port[_omp].so   (plain-C restatement, oracle/port.c)
  ref   = oracle/_ref/libfeature_ref[_omp].so (the reference's own Feature.hpp functions,
          built by oracle/build_ref.sh from /root/reference; prebuilt file travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)


def build(quiet=True):
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


def _fp(a):
    return a.ctypes.data_as(_f)


class Port:
    def __init__(self, omp=False):
        path = os.path.join(HERE, "liboracle_port_omp.so" if omp else "liboracle_port.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.port_simd_dot.restype = C.c_float
        L.port_simd_dot.argtypes = [_f, _f, C.c_long]
        L.port_gallery_scores.restype = None
        L.port_gallery_scores.argtypes = [_f, _f, C.c_long, C.c_long, _f, _f]
        L.port_gallery_argmax.restype = C.c_int
        L.port_gallery_argmax.argtypes = [_f, _f, C.c_long, C.c_long, C.c_float, _f, _f, _f]
        L.port_gallery_topk.restype = C.c_int
        L.port_gallery_topk.argtypes = [_f, _f, C.c_long, C.c_long, C.c_float, _f, C.c_int, _f, _f, _i]

    def simd_dot(self, x, y):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        return np.float32(self.lib.port_simd_dot(_fp(x), _fp(y), len(x)))

    def scores(self, G, gnorm, q):
        G = np.ascontiguousarray(G, np.float32); gnorm = np.ascontiguousarray(gnorm, np.float32)
        q = np.ascontiguousarray(q, np.float32)
        out = np.empty(G.shape[0], np.float32)
        self.lib.port_gallery_scores(_fp(G), _fp(gnorm), G.shape[0], G.shape[1], _fp(q), _fp(out))
        return out

    def argmax(self, G, gnorm, q, sim_th):
        G = np.ascontiguousarray(G, np.float32); gnorm = np.ascontiguousarray(gnorm, np.float32)
        q = np.ascontiguousarray(q, np.float32)
        scratch = np.empty(G.shape[0], np.float32); ms = C.c_float(0)
        r = self.lib.port_gallery_argmax(_fp(G), _fp(gnorm), G.shape[0], G.shape[1], sim_th, _fp(q),
                                         _fp(scratch), C.byref(ms))
        return r, np.float32(ms.value)

    def topk(self, G, gnorm, q, K, sim_th):
        G = np.ascontiguousarray(G, np.float32); gnorm = np.ascontiguousarray(gnorm, np.float32)
        q = np.ascontiguousarray(q, np.float32)
        scratch = np.empty(G.shape[0], np.float32)
        s = np.zeros(K, np.float32); ix = np.full(K, -1, np.int32)
        st = self.lib.port_gallery_topk(_fp(G), _fp(gnorm), G.shape[0], G.shape[1], sim_th, _fp(q), K,
                                        _fp(scratch), _fp(s), ix.ctypes.data_as(_i))
        return st, s, ix


class Ref:
    """The reference's own compiled functions; `available` False when oracle/_ref is absent."""

    def __init__(self, omp=False):
        path = os.path.join(HERE, "_ref", "libfeature_ref_omp.so" if omp else "libfeature_ref.so")
        self.lib = _load(path)
        self.available = self.lib is not None
        if not self.available:
            return
        L = self.lib
        L.ref_simd_dot.restype = C.c_float
        L.ref_simd_dot.argtypes = [_f, _f, C.c_long]
        L.ref_compare_face_from_db.restype = C.c_int
        L.ref_compare_face_from_db.argtypes = [_f, _f, C.c_int, C.c_int, C.c_float, _f]
        L.ref_compare_face_db.restype = C.c_int
        L.ref_compare_face_db.argtypes = [_f, _f, _i, C.c_int, C.c_int, C.c_float, _f]
        L.ref_compare_face_person.restype = C.c_int
        L.ref_compare_face_person.argtypes = [_f, _f, C.c_int, C.c_int, C.c_float, _f, _f, C.c_int, C.c_float]

    def simd_dot(self, x, y):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        return np.float32(self.lib.ref_simd_dot(_fp(x), _fp(y), len(x)))

    def compare_face_from_db(self, G, gnorm, q, sim_th):
        G = np.ascontiguousarray(G, np.float32); gnorm = np.ascontiguousarray(gnorm, np.float32)
        q = np.ascontiguousarray(q, np.float32)
        return self.lib.ref_compare_face_from_db(_fp(G), _fp(gnorm), G.shape[0], G.shape[1], sim_th, _fp(q))

    def compare_face_db(self, G, gnorm, pids, q, sim_th):
        G = np.ascontiguousarray(G, np.float32); gnorm = np.ascontiguousarray(gnorm, np.float32)
        q = np.ascontiguousarray(q, np.float32); pids = np.ascontiguousarray(pids, np.int32)
        return self.lib.ref_compare_face_db(_fp(G), _fp(gnorm), pids.ctypes.data_as(_i), G.shape[0],
                                            G.shape[1], sim_th, _fp(q))

    def compare_face_person(self, G, gnorm, q, sim_th, maxsim, qnorm=None):
        G = np.ascontiguousarray(G, np.float32); gnorm = np.ascontiguousarray(gnorm, np.float32)
        q = np.ascontiguousarray(q, np.float32)
        ms = C.c_float(maxsim)
        hit = self.lib.ref_compare_face_person(_fp(G), _fp(gnorm), G.shape[0], G.shape[1], sim_th, _fp(q),
                                               C.byref(ms), 0 if qnorm is None else 1,
                                               0.0 if qnorm is None else float(qnorm))
        return bool(hit), np.float32(ms.value)
