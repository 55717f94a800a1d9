"""ctypes binding of oracle/_build/libzkp_oracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY — see oracle/zkp_oracle.py.  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzkp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc, no reference sources)."""
    if force or not os.path.exists(_SO):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, u32p, u64p, i32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_int32))
        vp = C.c_void_p
        L.ora_mimc_hash.argtypes = [C.c_uint64, u8p]
        L.ora_mimc_constant.argtypes = [C.c_uint32, u8p]
        L.ora_ntt.argtypes = [vp, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.ora_msm_g1.argtypes = [vp, vp, C.c_size_t, vp, C.c_int]
        L.ora_msm_g2.argtypes = [vp, vp, C.c_size_t, vp, C.c_int]
        L.ora_g1_gen_mul.argtypes = [vp, C.c_size_t, vp]
        L.ora_g2_gen_mul.argtypes = [vp, C.c_size_t, vp]
        L.ora_circuit_equality.restype = vp
        L.ora_circuit_equality.argtypes = [C.c_uint32]
        L.ora_circuit_membership.restype = vp
        L.ora_circuit_membership.argtypes = [C.c_uint32]
        L.ora_circuit_free.argtypes = [vp]
        L.ora_circuit_shape.argtypes = [vp, u64p]
        L.ora_circuit_matrix.argtypes = [vp, C.c_int, vp, vp, vp]
        L.ora_assign.argtypes = [vp, C.c_uint64, C.c_uint64, vp, C.c_uint32, vp, vp]
        L.ora_witness_map.argtypes = [vp, vp, vp, C.c_int]
        L.ora_pk_parse.restype = vp
        L.ora_pk_parse.argtypes = [vp, C.c_size_t]
        L.ora_pk_free.argtypes = [vp]
        L.ora_pk_size.restype = C.c_size_t
        L.ora_pk_size.argtypes = [vp]
        L.ora_vk_size.restype = C.c_size_t
        L.ora_vk_size.argtypes = [vp]
        L.ora_setup.argtypes = [vp, vp, vp, vp]
        L.ora_prove.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int]
        L.ora_prove_batch.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp, C.c_uint32, vp, vp, vp, vp, C.c_int]
        L.ora_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fr_array(vals: Sequence[int]) -> np.ndarray:
    """ints -> (n, 32) uint8 canonical LE."""
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).reshape(-1, 32).copy()


def fr_list(arr: np.ndarray):
    b = np.ascontiguousarray(arr, dtype=np.uint8).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def mimc_hash(v: int) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().ora_mimc_hash(v, out)
    return bytes(out)


def mimc_constant(i: int) -> int:
    out = (C.c_uint8 * 32)()
    assert lib().ora_mimc_constant(i, out) == 0
    return int.from_bytes(bytes(out), "little")


def ntt(data: np.ndarray, inverse=False, coset=False, threads=0) -> np.ndarray:
    a = np.ascontiguousarray(data, dtype=np.uint8).copy()
    n = a.size // 32
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().ora_ntt(_ptr(a), log_n, int(inverse), int(coset), threads)
    assert rc == 0, rc
    return a


def msm_g1(bases: np.ndarray, scalars: np.ndarray, threads=0) -> bytes:
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    n = min(bases.size // 64, scalars.size // 32)
    out = np.zeros(64, dtype=np.uint8)
    rc = lib().ora_msm_g1(_ptr(bases), _ptr(scalars), n, _ptr(out), threads)
    assert rc == 0, rc
    return out.tobytes()


def msm_g2(bases: np.ndarray, scalars: np.ndarray, threads=0) -> bytes:
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    n = min(bases.size // 128, scalars.size // 32)
    out = np.zeros(128, dtype=np.uint8)
    rc = lib().ora_msm_g2(_ptr(bases), _ptr(scalars), n, _ptr(out), threads)
    assert rc == 0, rc
    return out.tobytes()


def g1_gen_mul(scalars: np.ndarray) -> np.ndarray:
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    n = scalars.size // 32
    out = np.zeros((n, 64), dtype=np.uint8)
    lib().ora_g1_gen_mul(_ptr(scalars), n, _ptr(out))
    return out


def g2_gen_mul(scalars: np.ndarray) -> np.ndarray:
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    n = scalars.size // 32
    out = np.zeros((n, 128), dtype=np.uint8)
    lib().ora_g2_gen_mul(_ptr(scalars), n, _ptr(out))
    return out


class Circuit:
    """kind 'equality' (rounds, default 110 = the reference circuit) or 'membership' (slots)."""

    def __init__(self, kind: str, param: Optional[int] = None):
        L = lib()
        self.kind = kind
        if kind == "equality":
            self.param = 110 if param is None else param
            self.h = L.ora_circuit_equality(self.param)
        elif kind == "membership":
            self.param = 64 if param is None else param
            self.h = L.ora_circuit_membership(self.param)
        else:
            raise ValueError(kind)
        shp = (C.c_uint64 * 7)()
        L.ora_circuit_shape(self.h, shp)
        self.m, self.n_inst, self.n_wit, self.log_n = int(shp[0]), int(shp[1]), int(shp[2]), int(shp[3])
        self.nnz = (int(shp[4]), int(shp[5]), int(shp[6]))
        self.n_vars = self.n_inst + self.n_wit
        self.n = 1 << self.log_n

    def __del__(self):
        try:
            lib().ora_circuit_free(self.h)
        except Exception:
            pass

    def matrix(self, which: int):
        rowptr = np.zeros(self.m + 1, dtype=np.uint32)
        col = np.zeros(self.nnz[which], dtype=np.uint32)
        val = np.zeros((self.nnz[which], 32), dtype=np.uint8)
        lib().ora_circuit_matrix(self.h, which, _ptr(rowptr), _ptr(col), _ptr(val))
        return rowptr, col, val

    def assign(self, a: int, b: int = 0, set_: Optional[Sequence[int]] = None,
               commitment: Optional[bytes] = None) -> Optional[np.ndarray]:
        if commitment is None:
            commitment = mimc_hash(a)
        z = np.zeros((self.n_vars, 32), dtype=np.uint8)
        cm = np.frombuffer(commitment, dtype=np.uint8).copy()
        sa = None if set_ is None else np.asarray(list(set_), dtype=np.uint64)
        rc = lib().ora_assign(self.h, a, b, _ptr(sa), 0 if sa is None else len(sa), _ptr(cm), _ptr(z))
        return z if rc == 0 else None

    def witness_map(self, z: np.ndarray, threads=0) -> np.ndarray:
        z = np.ascontiguousarray(z, dtype=np.uint8)
        h = np.zeros((self.n, 32), dtype=np.uint8)
        assert lib().ora_witness_map(self.h, _ptr(z), _ptr(h), threads) == 0
        return h

    def setup(self, trapdoor: Sequence[int]):
        """trapdoor = (alpha, beta, gamma, delta, tau) -> (pk_bytes, vk_bytes)."""
        td = fr_array(trapdoor)
        pk = np.zeros(lib().ora_pk_size(self.h), dtype=np.uint8)
        vk = np.zeros(lib().ora_vk_size(self.h), dtype=np.uint8)
        rc = lib().ora_setup(self.h, _ptr(td), _ptr(pk), _ptr(vk))
        assert rc == 0, rc
        return pk.tobytes(), vk.tobytes()


class ProvingKey:
    def __init__(self, pk_bytes: bytes):
        buf = np.frombuffer(pk_bytes, dtype=np.uint8)
        self.h = lib().ora_pk_parse(_ptr(buf), len(pk_bytes))
        if not self.h:
            raise ValueError("pk parse failed")

    def __del__(self):
        try:
            lib().ora_pk_free(self.h)
        except Exception:
            pass


def prove(circ: Circuit, pk: ProvingKey, z: np.ndarray, r: int, s: int, threads=0) -> bytes:
    z = np.ascontiguousarray(z, dtype=np.uint8)
    out = np.zeros(256, dtype=np.uint8)
    rr, ss = fr_array([r]), fr_array([s])
    rc = lib().ora_prove(circ.h, pk.h, _ptr(z), _ptr(rr), _ptr(ss), _ptr(out), threads)
    assert rc == 0, rc
    return out.tobytes()


def prove_batch(circ: Circuit, pk: ProvingKey, a: np.ndarray, b: Optional[np.ndarray],
                sets: Optional[np.ndarray], set_len: Optional[np.ndarray],
                r: np.ndarray, s: np.ndarray, threads=0):
    """Proof-parallel CPU batch (the reference's process_batch shape). Returns (proofs[n,256], status[n])."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    n = len(a)
    b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    stride = 0
    if sets is not None:
        sets = np.ascontiguousarray(sets, dtype=np.uint64)
        stride = sets.shape[1]
        set_len = np.ascontiguousarray(set_len, dtype=np.uint32)
    r = np.ascontiguousarray(r, dtype=np.uint8)
    s = np.ascontiguousarray(s, dtype=np.uint8)
    out = np.zeros((n, 256), dtype=np.uint8)
    status = np.zeros(n, dtype=np.int32)
    rc = lib().ora_prove_batch(circ.h, pk.h, n, _ptr(a), _ptr(b), _ptr(sets), _ptr(set_len), stride,
                               _ptr(r), _ptr(s), _ptr(out), _ptr(status), threads)
    assert rc == 0
    return out, status


def num_threads() -> int:
    return lib().ora_num_threads()
