"""oracle/ref.py — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libstts_ref.so: the UNMODIFIED reference NN objects
(/root/reference/src/{nn_op,modules,models}) compiled in place by oracle/Makefile, driven at the
phoneme-ID level by oracle/ref_driver.cpp.  Only tests/, bench.py's cpu_baseline / --impl reference
arm and __graft_entry__.smoke() may import this module; the product (summertts_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libstts_ref.so")


class _Result(C.Structure):
    _fields_ = [
        ("T", C.c_int32), ("F", C.c_int32), ("S", C.c_int32), ("hidden", C.c_int32), ("inter", C.c_int32),
        ("xx", C.POINTER(C.c_float)), ("m", C.POINTER(C.c_float)), ("logw", C.POINTER(C.c_float)),
        ("w_ceil", C.POINTER(C.c_float)), ("z_p", C.POINTER(C.c_float)), ("z", C.POINTER(C.c_float)),
        ("o", C.POINTER(C.c_float)), ("pcm", C.POINTER(C.c_int16)), ("ms", C.c_double * 6),
    ]


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.sref_create.restype = C.c_void_p
        L.sref_create.argtypes = [C.c_void_p, C.c_int64]
        L.sref_nn_end.restype = C.c_int64
        L.sref_nn_end.argtypes = [C.c_void_p]
        L.sref_speaker_num.restype = C.c_int32
        L.sref_speaker_num.argtypes = [C.c_void_p]
        L.sref_lang_type.restype = C.c_int32
        L.sref_lang_type.argtypes = [C.c_void_p]
        L.sref_infer.restype = C.c_int
        L.sref_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int,
                                 C.POINTER(_Result)]
        L.sref_free_result.argtypes = [C.POINTER(_Result)]
        L.sref_destroy.argtypes = [C.c_void_p]
        L.sref_set_threads.argtypes = [C.c_int]
        L.sref_get_threads.restype = C.c_int
        L.sref_free.argtypes = [C.c_void_p]
        for name in ("sref_conv1d", "sref_conv1d_transposed", "sref_layer_norm", "sref_istft", "sref_pqmf",
                     "sref_wn", "sref_resblock1", "sref_mha", "sref_eltwise"):
            getattr(L, name).restype = C.POINTER(C.c_float)
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _take(ptr, shape):
    n = int(np.prod(shape))
    out = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].copy().reshape(shape)
    return out


@dataclass
class RefOutput:
    T: int
    F: int
    S: int
    pcm: np.ndarray
    ms: dict
    xx: np.ndarray | None = None
    m: np.ndarray | None = None
    logw: np.ndarray | None = None
    w_ceil: np.ndarray | None = None
    z_p: np.ndarray | None = None
    z: np.ndarray | None = None
    o: np.ndarray | None = None
    extra: dict = field(default_factory=dict)


class RefModel:
    """The reference's NN stack built from a .bin blob (SynthesizerTrn.cpp:101-167)."""

    def __init__(self, blob: np.ndarray):
        self._blob = _f32(blob)  # ctors copy weights, but keep alive anyway (emg_ is copied too)
        self._h = lib().sref_create(self._blob.ctypes.data, self._blob.size)
        if not self._h:
            raise RuntimeError("sref_create failed")
        self.nn_end = lib().sref_nn_end(self._h)
        self.speaker_num = lib().sref_speaker_num(self._h)
        self.lang_type = lib().sref_lang_type(self._h)

    def infer(self, ids, sid=0, length_scale=1.0, forced_w=None, dumps=True) -> RefOutput:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        fw = None if forced_w is None else _f32(forced_w)
        R = _Result()
        rc = lib().sref_infer(self._h, ids.ctypes.data, ids.size, int(sid), float(length_scale),
                              None if fw is None else fw.ctypes.data, 1 if dumps else 0, C.byref(R))
        if rc != 0:
            raise RuntimeError("sref_infer failed")
        try:
            ms = dict(zip(("enc", "dp", "expand", "flow", "dec", "total"), list(R.ms)))
            out = RefOutput(T=R.T, F=R.F, S=R.S, pcm=_take(R.pcm, (R.S,)), ms=ms)
            if dumps:
                out.xx = _take(R.xx, (R.T, R.hidden))
                out.m = _take(R.m, (R.T, R.inter))
                out.logw = _take(R.logw, (R.T,))
                out.w_ceil = _take(R.w_ceil, (R.T,))
                out.z_p = _take(R.z_p, (R.F, R.inter))
                out.z = _take(R.z, (R.F, R.inter))
                out.o = _take(R.o, (R.S,))
        finally:
            lib().sref_free_result(C.byref(R))
        return out

    def close(self):
        if self._h:
            lib().sref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_threads(n: int):
    lib().sref_set_threads(int(n))


def _call_tm(fn, *args):
    To, Co = C.c_int32(), C.c_int32()
    p = fn(*args, C.byref(To), C.byref(Co))
    y = _take(p, (To.value, Co.value))
    lib().sref_free(p)
    return y


def conv1d(rec, x, mode=0, pad=0, dil=1, sep=0):
    rec, x = _f32(rec), _f32(x)
    return _call_tm(lib().sref_conv1d, C.c_void_p(rec.ctypes.data), C.c_void_p(x.ctypes.data), C.c_int32(x.shape[0]),
                    C.c_int32(x.shape[1]), C.c_int32(mode), C.c_int32(pad), C.c_int32(dil), C.c_int32(sep))


def conv1d_transposed(rec, x, stride, pad):
    rec, x = _f32(rec), _f32(x)
    return _call_tm(lib().sref_conv1d_transposed, C.c_void_p(rec.ctypes.data), C.c_void_p(x.ctypes.data),
                    C.c_int32(x.shape[0]), C.c_int32(x.shape[1]), C.c_int32(stride), C.c_int32(pad))


def layer_norm(rec, x):
    rec, x = _f32(rec), _f32(x)
    return _call_tm(lib().sref_layer_norm, C.c_void_p(rec.ctypes.data), C.c_void_p(x.ctypes.data),
                    C.c_int32(x.shape[0]), C.c_int32(x.shape[1]))


def istft(mag, phase):
    mag, phase = _f32(mag), _f32(phase)
    return _call_tm(lib().sref_istft, C.c_void_p(mag.ctypes.data), C.c_void_p(phase.ctypes.data),
                    C.c_int32(mag.shape[0]), C.c_int32(mag.shape[1]))


def pqmf(x):
    x = _f32(x)
    return _call_tm(lib().sref_pqmf, C.c_void_p(x.ctypes.data), C.c_int32(x.shape[0]))


def wn(rec, x, g=None):
    rec, x = _f32(rec), _f32(x)
    gg = None if g is None else _f32(g)
    return _call_tm(lib().sref_wn, C.c_void_p(rec.ctypes.data), C.c_int32(0 if g is None else 1),
                    C.c_void_p(x.ctypes.data), C.c_int32(x.shape[0]), C.c_int32(x.shape[1]),
                    C.c_void_p(None if gg is None else gg.ctypes.data), C.c_int32(0 if gg is None else gg.size))


def resblock1(rec, x):
    rec, x = _f32(rec), _f32(x)
    return _call_tm(lib().sref_resblock1, C.c_void_p(rec.ctypes.data), C.c_void_p(x.ctypes.data),
                    C.c_int32(x.shape[0]), C.c_int32(x.shape[1]))


def mha(rec, x):
    rec, x = _f32(rec), _f32(x)
    return _call_tm(lib().sref_mha, C.c_void_p(rec.ctypes.data), C.c_void_p(x.ctypes.data),
                    C.c_int32(x.shape[0]), C.c_int32(x.shape[1]))


def eltwise(which, x):
    x = _f32(x).ravel()
    p = lib().sref_eltwise(C.c_int32(which), C.c_void_p(x.ctypes.data), C.c_int32(x.size))
    y = _take(p, (x.size,))
    lib().sref_free(p)
    return y


class RefG2p:
    """The reference's unmodified EnglishText2Id object + its GRU internals (oracle/ref_g2p.cpp)."""

    def __init__(self, section: np.ndarray):
        L = lib()
        if not hasattr(L, "_g2p_bound"):
            L.sref_g2p_create.restype = C.c_void_p
            L.sref_g2p_create.argtypes = [C.c_void_p, C.c_int64]
            for name in ("sref_g2p_consumed", "sref_g2p_hidden", "sref_g2p_phones"):
                getattr(L, name).restype = C.c_int32
                getattr(L, name).argtypes = [C.c_void_p]
            L.sref_g2p_destroy.argtypes = [C.c_void_p]
            L.sref_g2p_ipa_ids.restype = C.c_int32
            L.sref_g2p_ipa_ids.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int32]
            L.sref_g2p_word.restype = C.c_int32
            L.sref_g2p_word.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L._g2p_bound = True
        sec = _f32(section)
        self._h = L.sref_g2p_create(sec.ctypes.data, sec.size)
        self.consumed = L.sref_g2p_consumed(self._h)
        self.hidden = L.sref_g2p_hidden(self._h)
        self.phones = L.sref_g2p_phones(self._h)

    def ipa_ids(self, text: str) -> list[int]:
        """EnglishText2Id::getIPAId(text), unmodified (EnglishText2Id.cpp:456-609)."""
        out = np.zeros(65536, np.int32)
        n = lib().sref_g2p_ipa_ids(self._h, text.encode("utf-8"), out.ctypes.data, out.size)
        return out[:n].tolist()

    def word(self, word: bytes | str):
        """The out-of-vocabulary branch (:496-545) over the reference's gru / gru_cell: (preds, hidden, logits0)."""
        if isinstance(word, str):
            word = word.encode("utf-8")
        preds = np.zeros(20, np.int32)
        hid = np.zeros(self.hidden, np.float32)
        lg = np.zeros(self.phones, np.float32)
        n = lib().sref_g2p_word(self._h, word, preds.ctypes.data, hid.ctypes.data, lg.ctypes.data)
        return preds[:n].tolist(), hid, lg

    def close(self):
        if self._h:
            lib().sref_g2p_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
