"""libzkp_b200 — B200-native Groth16/BN254 proving engine behind libzkp's SNARK backend.

The module-level functions carry the names of the reference's PyO3 module ``libzkp``
(src/python_api.rs:110-164) for the Groth16 prover path: ``prove_equality``, ``prove_membership``,
``snark_commit_value``, the batch API and the key-dir switches.  Everything computes on the GPU
through the C ABI of include/lzkp_b200.h; there is no CPU fallback.
"""
from .errors import (BackendError, ConfigError, CryptoError, EngineError, InvalidInput, InvalidProofFormat,
                     ProofGenerationFailed, ZkpError)
from .proof import Proof, commit_value_snark, prove_equality, prove_membership
from .batch import (batch_add_equality_proof, batch_add_membership_proof, clear_batch, create_proof_batch,
                    get_batch_status, process_batch)
from .snark import SnarkBackend, set_snark_key_dir
from .snark import is_snark_initialized as is_snark_setup_initialized

snark_commit_value = commit_value_snark        # python_api.rs:33

__all__ = [
    "prove_equality", "prove_membership", "snark_commit_value", "commit_value_snark", "create_proof_batch",
    "batch_add_equality_proof", "batch_add_membership_proof", "process_batch", "get_batch_status", "clear_batch",
    "set_snark_key_dir", "is_snark_setup_initialized", "SnarkBackend", "Proof", "ZkpError", "InvalidInput",
    "InvalidProofFormat", "ConfigError", "ProofGenerationFailed", "BackendError", "CryptoError", "EngineError",
]
