"""zinc_b200 -- B200 (sm_100a) implementation of the commit path of zinc's Zip polynomial commitment.

Only what the path needs: csrc/ (CUDA kernels + the C ABI of include/zipgpu.h, built into libzipgpu.so),
and the host-side mirror of the reference's commit API (zip.py, transcript.py) over that C ABI.
"""
from .pcs_transcript import PcsStream  # noqa: F401
from .transcript import KeccakTranscript, MockTranscript  # noqa: F401
from .zip import (  # noqa: F401
    Context,
    DefaultLinearCodeSpec,
    DenseMultilinearExtension,
    Error,
    InvalidPcsParam,
    MerkleProof,
    MultiContext,
    MerkleTree,
    MultilinearZip,
    MultilinearZipCommitment,
    MultilinearZipData,
    MultilinearZipParams,
    RaaCode,
    RandomFieldZipTypes,
    ResidentZipData,
    SparseMatrixZ,
    ZipLinearCode,
    ZipTypes,
    default_context,
    shuffle_seeded_indices,
)
