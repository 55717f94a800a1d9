"""libzkp_b200 — B200-native Groth16/BN254 proving engine behind libzkp's SNARK backend.

The module-level functions carry the names of the reference's PyO3 module ``libzkp``
(src/python_api.rs:110-164) for the Groth16 prover path: ``prove_equality``, ``prove_membership``,
``snark_commit_value``, the batch API and the key-dir switches.  Everything computes on the GPU
through the C ABI of include/lzkp_b200.h; there is no CPU fallback.
"""
from .errors import (BackendError, ConfigError, CryptoError, EngineError, InvalidInput, InvalidProofFormat,
                     ProofGenerationFailed, ZkpError)
from .proof import (Proof, commit_value_snark, prove_equality, prove_membership, verify_equality,
                    verify_equality_with_commitment, verify_membership)
from .batch import (batch_add_equality_proof, batch_add_membership_proof, clear_batch, create_proof_batch,
                    get_batch_status, process_batch)
from .snark import SnarkBackend, set_snark_key_dir
from .snark import is_snark_initialized as is_snark_setup_initialized

snark_commit_value = commit_value_snark        # python_api.rs:33


def verify_proofs_parallel(proofs):
    """verify_proofs_parallel (utils/performance.rs:251-267) for the two Groth16 kinds: a list of (proof_bytes, type)
    with type "equality" / "membership"; proofs of one kind are verified in one device call.  Like the reference's
    generic dispatcher it checks each proof against the commitment / set embedded in its own envelope."""
    from . import proof as _p
    out = [False] * len(proofs)
    eq, mb = [], []
    for i, (pb, kind) in enumerate(proofs):
        try:
            env = Proof.from_bytes(bytes(pb))
        except Exception:                           # noqa: BLE001
            continue
        if kind == "equality":
            eq.append((i, (pb, env.commitment)))
        elif kind == "membership":
            emb = _p._embedded_set(env.proof)
            if emb is not None:
                mb.append((i, (pb, emb[0])))
    for (i, _), ok in zip(eq, _p.verify_equality_many([x for _, x in eq])):
        out[i] = ok
    for (i, _), ok in zip(mb, _p.verify_membership_many([x for _, x in mb])):
        out[i] = ok
    return out

__all__ = [
    "prove_equality", "prove_membership", "verify_equality", "verify_equality_with_commitment", "verify_membership",
    "verify_proofs_parallel", "snark_commit_value", "commit_value_snark", "create_proof_batch",
    "batch_add_equality_proof", "batch_add_membership_proof", "process_batch", "get_batch_status", "clear_batch",
    "set_snark_key_dir", "is_snark_setup_initialized", "SnarkBackend", "Proof", "ZkpError", "InvalidInput",
    "InvalidProofFormat", "ConfigError", "ProofGenerationFailed", "BackendError", "CryptoError", "EngineError",
]
