"""Error types of the host-side mirror.

Mirrors ``ZkpError`` (reference: src/utils/error_handling.rs:8-18) and the Python exception
each variant becomes through PyO3 (src/utils/error_handling.rs:39-49): ``InvalidInput`` ->
``ValueError``; ``InvalidProofFormat`` / ``ConfigError`` -> ``TypeError``; everything else ->
``RuntimeError`` carrying the ``Display`` text (":21-33").
"""


class ZkpError(Exception):
    prefix = "Error"

    def __init__(self, msg: str):
        self.msg = msg
        super().__init__(msg)


class InvalidInput(ZkpError, ValueError):
    prefix = "Invalid input"


class InvalidProofFormat(ZkpError, TypeError):
    prefix = "Invalid proof format"


class ConfigError(ZkpError, TypeError):
    prefix = "Configuration error"


class _Runtime(ZkpError, RuntimeError):
    def __init__(self, msg: str):
        ZkpError.__init__(self, f"{self.prefix}: {msg}")
        self.msg = msg


class ProofGenerationFailed(_Runtime):
    prefix = "Proof generation failed"


class BackendError(_Runtime):
    prefix = "Backend error"


class CryptoError(_Runtime):
    prefix = "Cryptographic error"


class EngineError(_Runtime):
    """A non-zero return code from the C ABI (include/lzkp_b200.h); carries lzkp_last_error()."""
    prefix = "Backend error"

    def __init__(self, code: int, msg: str):
        super().__init__(f"lzkp error {code}: {msg}")
        self.code = code
