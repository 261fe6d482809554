"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/a3d_oracle.c (the fast C
checker).  Only tests/, smoke() and bench.py's CPU legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liba3d_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "a3d_oracle.c")):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(SO)
        _lib.a3do_pitch_words.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack(masks: np.ndarray, thresh=0.5, nonzero=False) -> np.ndarray:
    lib = load()
    m = np.ascontiguousarray(masks, dtype=np.float32)
    n, H, W = m.shape
    bits = np.empty((n, H, lib.a3do_pitch_words(W)), dtype=np.uint32)
    lib.a3do_pack_f32(_p(m), C.c_int64(n), H, W, C.c_float(thresh), int(nonzero), _p(bits))
    return bits


def project(kinv, f, cx, cy, H, W, src_bits, normal, offset, pivot, mode, xform) -> np.ndarray:
    lib = load()
    kinv = np.ascontiguousarray(kinv, dtype=np.float64).reshape(9)
    normal = np.ascontiguousarray(normal, dtype=np.float32)
    pivot = np.ascontiguousarray(pivot, dtype=np.float32)
    xform = np.ascontiguousarray(xform, dtype=np.float32).reshape(-1, 12)
    src_bits = np.ascontiguousarray(src_bits, dtype=np.uint32)
    A = xform.shape[0]
    out = np.empty((A, H, lib.a3do_pitch_words(W)), dtype=np.uint32)
    lib.a3do_project(_p(kinv), C.c_float(f), C.c_float(cx), C.c_float(cy), H, W, _p(src_bits), _p(normal),
                     C.c_float(offset), _p(pivot), int(mode), _p(xform), A, _p(out))
    return out


def score(H, W, tgt_bits, proj_bits):
    lib = load()
    tgt_bits = np.ascontiguousarray(tgt_bits, dtype=np.uint32)
    proj_bits = np.ascontiguousarray(proj_bits, dtype=np.uint32)
    T, A = tgt_bits.shape[0], proj_bits.shape[0]
    inter = np.empty((T, A), np.int32)
    uni = np.empty((T, A), np.int32)
    best = np.empty(T, np.int32)
    iou = np.empty(T, np.float32)
    lib.a3do_score(H, W, _p(tgt_bits), T, _p(proj_bits), A, _p(inter), _p(uni), _p(best), _p(iou))
    return inter, uni, best, iou


def unpack(bits: np.ndarray, W: int) -> np.ndarray:
    u8 = np.ascontiguousarray(bits).view(np.uint8).reshape(bits.shape[0], bits.shape[1], -1)
    return np.unpackbits(u8, axis=-1, bitorder="little")[..., :W].astype(bool)
