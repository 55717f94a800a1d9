"""Host-side mirror of the reference's SNARK backend (src/backend/snark.rs) over the C ABI.

Same names, argument meaning and failure convention as ``SnarkBackend`` (snark.rs:293-495) and the
``ZkpBackend`` trait (src/backend/mod.rs:5-8): proving returns ``b""`` on any failure
(snark.rs:345,350,361,366,371).  What differs is where the work happens: the calls the reference
makes into ark-groth16 (``Groth16::<Bn254>::prove``, snark.rs:364 and :442) are replaced by
``lzkp_prove_equality_batch`` / ``lzkp_prove_membership_batch`` on a proving key that was uploaded
once (the reference's OnceLock fill, snark.rs:295-339) and stays resident in HBM.

In production this layer is Rust (INTEGRATION.md shows the ``extern "C"`` shim); no Rust toolchain
exists in this image, so the mirror is Python.  It never touches ``oracle/``.
"""
from __future__ import annotations

import os
import secrets
import threading
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine
from .errors import ConfigError

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
MIMC_ROUNDS = 110          # snark.rs:182
MAX_SET_SIZE = 64          # snark.rs:503

_lock = threading.RLock()
_key_dir_override: Optional[str] = None
_setups: Dict[str, object] = {}      # prefix -> Setup | Exception text (OnceLock<Result<..>>)
# Test / deployment hook: a callable(prefix) -> (pk_bytes, vk_bytes) used when no key files exist.
_generator: Optional[Callable[[str], Tuple[bytes, bytes]]] = None
# validate None = the reference's behaviour: keys READ FROM FILES are validated (deserialize_uncompressed checks curve and
# subgroup membership, snark.rs:64-66), keys this process just generated are not
_pk_options = {"window_bits": 0, "table_budget_bytes": 0, "max_chunk": 0, "validate": None}


class OsRng:
    """Prover randomness r, s: the reference draws Fr::rand from OsRng (snark.rs:363,441)."""

    def next_fr(self) -> int:
        return secrets.randbelow(R_MOD)

    def scalars(self, n: int) -> np.ndarray:
        """n uniform canonical scalars as an (n, 32) byte array: OS entropy, 254-bit mask, rejection of values
        >= r (the sampling rule arkworks' Fr::rand applies), vectorised so that a 4096-proof batch does not spend
        longer drawing randomness on the host than proving on the device."""
        out = np.zeros((0, 4), np.uint64)
        r_words = np.array([(R_MOD >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], np.uint64)
        while out.shape[0] < n:
            m = max(64, int((n - out.shape[0]) * 1.5))
            w = np.frombuffer(os.urandom(32 * m), dtype="<u8").reshape(m, 4).copy()
            w[:, 3] &= np.uint64((1 << 62) - 1)
            lt = np.zeros(m, bool)
            eq = np.ones(m, bool)
            for i in (3, 2, 1, 0):                      # lexicographic compare, most significant word first
                lt |= eq & (w[:, i] < r_words[i])
                eq &= w[:, i] == r_words[i]
            out = np.concatenate([out, w[lt]])
        return out[:n].view(np.uint8).reshape(n, 32).copy()


class Setup:
    def __init__(self, pk: engine.ProvingKey, vk_bytes: bytes):
        self.pk = pk
        self.vk_bytes = vk_bytes
        self._vk: Optional[engine.VerifyingKey] = None

    @property
    def vk(self) -> engine.VerifyingKey:
        """Device verifying key with e(alpha, beta) precomputed once (the reference recomputes process_vk per call)."""
        if self._vk is None:
            self._vk = engine.VerifyingKey(self.vk_bytes)
        return self._vk


def configure(window_bits: int = 0, table_budget_bytes: int = 0, max_chunk: int = 0, validate: Optional[bool] = None,
              generator: Optional[Callable[[str], Tuple[bytes, bytes]]] = None) -> None:
    """Engine options used when a setup is first loaded (window size of the resident tables, ...)."""
    global _generator
    with _lock:
        _pk_options.update(window_bits=window_bits, table_budget_bytes=table_budget_bytes, max_chunk=max_chunk,
                           validate=validate)
        _generator = generator


def reset() -> None:
    """Drop loaded setups and the key-dir override (tests only; the reference has no equivalent)."""
    global _key_dir_override
    with _lock:
        for v in _setups.values():
            if isinstance(v, Setup):
                v.pk.close()
                if v._vk is not None:
                    v._vk.close()
        _setups.clear()
        _key_dir_override = None


def _get_key_dir() -> Optional[str]:               # snark.rs:23-30
    with _lock:
        if _key_dir_override is not None:
            return _key_dir_override
    return os.environ.get("LIBZKP_SNARK_KEY_DIR")


def _key_paths(prefix: str) -> Optional[Tuple[str, str]]:   # snark.rs:32-38
    d = _get_key_dir()
    if d is None:
        return None
    return os.path.join(d, f"{prefix}_pk.bin"), os.path.join(d, f"{prefix}_vk.bin")


def set_snark_key_dir(path: str) -> bool:          # snark.rs:141-171 (advanced::set_snark_key_dir returns true)
    global _key_dir_override
    if path == "":
        raise ConfigError("SNARK key directory cannot be empty")
    with _lock:
        if _setups:
            raise ConfigError("SNARK setup is already initialized; set LIBZKP_SNARK_KEY_DIR before first proof")
        if _key_dir_override is not None and _key_dir_override != path:
            raise ConfigError(f"SNARK key directory already set to {_key_dir_override}; new value {path} rejected")
        _key_dir_override = path
    return True


def is_snark_initialized() -> bool:                # snark.rs:173-175
    with _lock:
        return bool(_setups)


def _generate(prefix: str) -> Tuple[bytes, bytes]:
    if _generator is not None:
        return _generator(prefix)
    from . import setup as _setup               # device-side circuit_specific_setup
    return _setup.generate(prefix)


def _load_or_generate(prefix: str) -> Tuple[bytes, bytes, bool]:   # snark.rs:122-139; third item: read from files
    paths = _key_paths(prefix)
    if paths is not None:
        pk_path, vk_path = paths
        if os.path.exists(pk_path) and os.path.exists(vk_path):
            with open(pk_path, "rb") as f:
                pk = f.read()
            with open(vk_path, "rb") as f:
                vk = f.read()
            return pk, vk, True
        pk, vk = _generate(prefix)
        try:                                       # best effort, errors ignored (snark.rs:130-132)
            os.makedirs(os.path.dirname(pk_path) or ".", exist_ok=True)
            with open(pk_path, "wb") as f:
                f.write(pk)
            with open(vk_path, "wb") as f:
                f.write(vk)
        except OSError:
            pass
        return pk, vk, False
    pk, vk = _generate(prefix)
    return pk, vk, False


_CIRCUITS = {"equality_mimc": (engine.EQUALITY, MIMC_ROUNDS), "membership_mimc": (engine.MEMBERSHIP, MAX_SET_SIZE)}


def _get_setup(prefix: str):
    """OnceLock<Result<SnarkKeyPair, String>> (snark.rs:295-327): the outcome, good or bad, is kept."""
    with _lock:
        got = _setups.get(prefix)
        if got is None:
            try:
                pk_bytes, vk_bytes, from_file = _load_or_generate(prefix)
                o = _pk_options
                validate = from_file if o["validate"] is None else bool(o["validate"])
                pk = engine.ProvingKey(pk_bytes, validate=validate, window_bits=o["window_bits"],
                                       table_budget_bytes=o["table_budget_bytes"], max_chunk=o["max_chunk"])
                kind, param = _CIRCUITS[prefix]
                pk.circuit_builtin(kind, param)
                got = Setup(pk, vk_bytes)
            except Exception as e:                 # noqa: BLE001 - mirrors Result<_, String>
                got = f"setup failed: {e!r}"
            _setups[prefix] = got
        return got


def mimc_commitment(value: int) -> bytes:
    """fr_to_commitment(mimc_hash_native(value)) (snark.rs:201-221)."""
    return engine.commit_value_snark(value)


def _fr_from_commitment(b: bytes) -> Optional[int]:   # snark.rs:224-229: canonical only
    if len(b) != 32:
        return None
    v = int.from_bytes(b, "little")
    return v if v < R_MOD else None


def _scalars(rng, n: int) -> np.ndarray:
    if hasattr(rng, "scalars"):
        return rng.scalars(n)
    return np.frombuffer(b"".join(rng.next_fr().to_bytes(32, "little") for _ in range(n)), np.uint8).reshape(n, 32)


class SnarkBackend:
    """Static-method mirror of ``SnarkBackend`` (snark.rs:293) + the batched entry points."""

    @staticmethod
    def get_universal_setup():
        return _get_setup("equality_mimc")

    @staticmethod
    def get_membership_setup():
        return _get_setup("membership_mimc")

    # ---- single proofs (the reference's signatures; rng is the OsRng seam, SURVEY.md headline fact 4)
    @staticmethod
    def prove_equality_zk(a: int, b: int, hash_input: bytes, rng=None) -> bytes:      # snark.rs:343-374
        out = SnarkBackend.prove_equality_zk_batch([a], [b], [hash_input], rng)
        return out[0]

    @staticmethod
    def prove_membership_zk(value: int, set_: Sequence[int], commitment: bytes, rng=None) -> bytes:  # :405-452
        out = SnarkBackend.prove_membership_zk_batch([value], [list(set_)], [commitment], rng)
        return out[0]

    # ---- batched: one device call for many independent proofs of one circuit
    @staticmethod
    def prove_equality_zk_batch(a: Sequence[int], b: Sequence[int], hash_inputs: Optional[Sequence[bytes]] = None,
                                rng=None, return_commitments: bool = False):
        """n x prove_equality_zk in one device call.  hash_inputs=None lets the device compute the MiMC
        commitments (commit_value_snark) itself; with return_commitments the pair (proofs, commitments) is returned."""
        n = len(a)
        out: List[bytes] = [b""] * n
        cms_out: List[bytes] = [b""] * n
        done = lambda: (out, cms_out) if return_commitments else out
        live = [i for i in range(n) if a[i] == b[i]
                and (hash_inputs is None or _fr_from_commitment(bytes(hash_inputs[i])) is not None)]
        if not live:
            return done()
        setup = SnarkBackend.get_universal_setup()
        if not isinstance(setup, Setup):
            return done()
        rng = rng or OsRng()
        rs = _scalars(rng, 2 * len(live))           # r then s per proof, the order prove() draws them
        cm_in = None
        if hash_inputs is not None:
            cm_in = np.frombuffer(b"".join(bytes(hash_inputs[i]) for i in live), np.uint8).reshape(-1, 32)
        try:
            av = np.array([a[i] for i in live], np.uint64)
            proofs, cms, status = setup.pk.prove_equality_batch(av, av, rs[0::2], rs[1::2], cm_in)
        except Exception:                           # noqa: BLE001 - any backend error -> empty Vec
            return done()
        pb, cb = proofs.tobytes(), cms.tobytes()
        for k, i in enumerate(live):
            if status[k] == 0:
                out[i] = pb[256 * k:256 * k + 256]
                cms_out[i] = cb[32 * k:32 * k + 32]
        return done()

    @staticmethod
    def prove_membership_zk_batch(values: Sequence[int], sets: Sequence[Sequence[int]],
                                  commitments: Optional[Sequence[bytes]] = None, rng=None,
                                  return_commitments: bool = False):
        n = len(values)
        out: List[bytes] = [b""] * n
        cms_out: List[bytes] = [b""] * n
        done = lambda: (out, cms_out) if return_commitments else out
        live = [i for i in range(n)
                if 1 <= len(sets[i]) <= MAX_SET_SIZE
                and (commitments is None or _fr_from_commitment(bytes(commitments[i])) is not None)
                and values[i] in sets[i]]
        if not live:
            return done()
        setup = SnarkBackend.get_membership_setup()
        if not isinstance(setup, Setup):
            return done()
        rng = rng or OsRng()
        rs = _scalars(rng, 2 * len(live))
        sets_arr = np.zeros((len(live), MAX_SET_SIZE), np.uint64)
        lens = np.zeros(len(live), np.uint32)
        for k, i in enumerate(live):
            lens[k] = len(sets[i])
            sets_arr[k, :lens[k]] = np.array(sets[i], np.uint64)
        cm_in = None
        if commitments is not None:
            cm_in = np.frombuffer(b"".join(bytes(commitments[i]) for i in live), np.uint8).reshape(-1, 32)
        try:
            proofs, cms, status = setup.pk.prove_membership_batch(
                np.array([values[i] for i in live], np.uint64), sets_arr, lens, rs[0::2], rs[1::2], cm_in)
        except Exception:                           # noqa: BLE001
            return done()
        pb, cb = proofs.tobytes(), cms.tobytes()
        for k, i in enumerate(live):
            if status[k] == 0:
                out[i] = pb[256 * k:256 * k + 256]
                cms_out[i] = cb[32 * k:32 * k + 32]
        return done()

    # ---- final libzkp bytes straight from the device (envelope framing fused into the batch call)
    @staticmethod
    def prove_equality_enveloped_batch(a: Sequence[int], b: Sequence[int], rng=None) -> List[bytes]:
        """[prove_equality(a_i, b_i)] as 298-byte envelopes; b"" where the reference would fail."""
        n = len(a)
        out: List[bytes] = [b""] * n
        live = [i for i in range(n) if a[i] == b[i]]
        setup = SnarkBackend.get_universal_setup() if live else None
        if not live or not isinstance(setup, Setup):
            return out
        rs = _scalars(rng or OsRng(), 2 * len(live))
        try:
            av = np.array([a[i] for i in live], np.uint64)
            env, lens, status = setup.pk.prove_equality_enveloped(av, av, rs[0::2], rs[1::2])
        except Exception:                           # noqa: BLE001
            return out
        eb = env.tobytes()
        for k, i in enumerate(live):
            if status[k] == 0:
                out[i] = eb[298 * k:298 * k + int(lens[k])]
        return out

    @staticmethod
    def prove_membership_enveloped_batch(values: Sequence[int], sets: Sequence[Sequence[int]], rng=None) -> List[bytes]:
        n = len(values)
        out: List[bytes] = [b""] * n
        live = [i for i in range(n) if 1 <= len(sets[i]) <= MAX_SET_SIZE and values[i] in sets[i]]
        setup = SnarkBackend.get_membership_setup() if live else None
        if not live or not isinstance(setup, Setup):
            return out
        rs = _scalars(rng or OsRng(), 2 * len(live))
        sets_arr = np.zeros((len(live), MAX_SET_SIZE), np.uint64)
        lens_in = np.zeros(len(live), np.uint32)
        for k, i in enumerate(live):
            lens_in[k] = len(sets[i])
            sets_arr[k, :lens_in[k]] = np.array(sets[i], np.uint64)
        try:
            env, lens, status = setup.pk.prove_membership_enveloped(
                np.array([values[i] for i in live], np.uint64), sets_arr, lens_in, rs[0::2], rs[1::2])
        except Exception:                           # noqa: BLE001
            return out
        row = env.shape[1]
        eb = env.tobytes()
        for k, i in enumerate(live):
            if status[k] == 0:
                out[i] = eb[row * k:row * k + int(lens[k])]
        return out

    # ---- verification (snark.rs:377-401, 455-495), batched on the device
    @staticmethod
    def verify_equality_zk_batch(proofs: Sequence[bytes], hash_inputs: Sequence[bytes]) -> List[bool]:
        n = len(proofs)
        out = [False] * n
        live = [i for i in range(n) if len(proofs[i]) == 256 and _fr_from_commitment(bytes(hash_inputs[i])) is not None]
        setup = SnarkBackend.get_universal_setup() if live else None
        if not live or not isinstance(setup, Setup):
            return out
        try:
            ok = setup.vk.verify_batch(np.frombuffer(b"".join(bytes(proofs[i]) for i in live), np.uint8),
                                       np.frombuffer(b"".join(bytes(hash_inputs[i]) for i in live), np.uint8))
        except Exception:                           # noqa: BLE001
            return out
        for k, i in enumerate(live):
            out[i] = bool(ok[k])
        return out

    @staticmethod
    def verify_equality_zk(proof_data: bytes, hash_input: bytes) -> bool:
        return SnarkBackend.verify_equality_zk_batch([proof_data], [hash_input])[0]

    @staticmethod
    def verify_membership_zk_batch(proofs: Sequence[bytes], sets: Sequence[Sequence[int]],
                                   commitments: Sequence[bytes]) -> List[bool]:
        n = len(proofs)
        out = [False] * n
        live = [i for i in range(n) if len(proofs[i]) == 256 and 1 <= len(sets[i]) <= MAX_SET_SIZE
                and _fr_from_commitment(bytes(commitments[i])) is not None]
        setup = SnarkBackend.get_membership_setup() if live else None
        if not live or not isinstance(setup, Setup):
            return out
        # public inputs [commitment, set[0..64) zero-padded, is_real[0..64)] (snark.rs:482-492)
        x = np.zeros((len(live), 1 + 2 * MAX_SET_SIZE, 32), np.uint8)
        for k, i in enumerate(live):
            x[k, 0] = np.frombuffer(bytes(commitments[i]), np.uint8)
            L = len(sets[i])
            x[k, 1:1 + L, :8] = np.array(sets[i], np.uint64).astype("<u8").view(np.uint8).reshape(L, 8)
            x[k, 1 + MAX_SET_SIZE:1 + MAX_SET_SIZE + L, 0] = 1
        try:
            ok = setup.vk.verify_batch(np.frombuffer(b"".join(bytes(proofs[i]) for i in live), np.uint8), x)
        except Exception:                           # noqa: BLE001
            return out
        for k, i in enumerate(live):
            out[i] = bool(ok[k])
        return out

    @staticmethod
    def verify_membership_zk(proof_data: bytes, set_: Sequence[int], commitment: bytes) -> bool:
        return SnarkBackend.verify_membership_zk_batch([proof_data], [list(set_)], [commitment])[0]

    @staticmethod
    def verify(proof: bytes, data: bytes) -> bool:          # ZkpBackend::verify (snark.rs:608-610)
        return SnarkBackend.verify_equality_zk(proof, data)

    # ---- ZkpBackend trait (src/backend/mod.rs:5-8; impl snark.rs:587-611)
    @staticmethod
    def prove(data: bytes) -> bytes:
        if len(data) != 48:
            return b""
        a = int.from_bytes(data[0:8], "little")
        b = int.from_bytes(data[8:16], "little")
        return SnarkBackend.prove_equality_zk(a, b, bytes(data[16:48]))
