"""ExactoError (error.rs:4-31) and the status codes of the C ABI."""
from __future__ import annotations


class ExactoError(Exception):
    """Mirror of the reference's error enum; ``kind`` is the variant name and ``str()``
    renders like the reference's ``Display`` (error.rs:6-30)."""

    KINDS = {
        1: ("InvalidParam", "invalid parameter: {}"),
        2: ("DimensionMismatch", "dimension mismatch: {}"),
        3: ("ModulusMismatch", "modulus mismatch"),
        4: ("InvalidRingDegree", "{}"),
        5: ("DecryptionError", "decryption error: noise budget exhausted"),
        6: ("DecompositionError", "decomposition error: {}"),
        7: ("LatticeError", "lattice error: {}"),
        8: ("MissingKey", "key not available: {}"),
        9: ("NotImplemented", "not yet implemented: {}"),
        100: ("Cuda", "CUDA error: {}"),
    }

    def __init__(self, code: int, detail: str = ""):
        kind, fmt = self.KINDS.get(code, ("Unknown", "{}"))
        super().__init__(fmt.format(detail) if "{}" in fmt else fmt)
        self.code = code
        self.kind = kind
        self.detail = detail


def InvalidParam(msg): return ExactoError(1, msg)
def DimensionMismatch(expected, got): return ExactoError(2, f"expected {expected}, got {got}")
def ModulusMismatch(): return ExactoError(3)
def InvalidRingDegree(n): return ExactoError(4, f"ring degree must be a power of 2, got {n}")
def NotImplementedErr(msg): return ExactoError(9, msg)
