"""gokalman_b200 -- B200-native batched Kalman-filter engine with gokalman's API surface.

Hand-written sm_100a CUDA kernels behind a C-ABI (include/gokalman_b200.h); this package is the
host-side mirror of the reference's Go API.  No CPU fallback: importing works anywhere, but every
filter call needs libgokalman_b200.so and a B200.
"""
from ._lib import GkbError, load, LIB_PATH  # noqa: F401
from .api import (  # noqa: F401
    AsSymDense, DenseIdentity, HouseholderTransf, Identity, IsNil, ScaledDenseIdentity, ScaledIdentity, Sign,
    AWGN, BatchGroundTruth, BatchKF, BatchNoise, ErrorEstimate, Estimate, HybridKF, NewBatchGroundTruth, NewBatchKF, Information, MonteCarloRuns, NewAWGN, NewChiSquare, NewHybridKF,
    NewInformation, NewInformationFromState, NewMonteCarloRuns, NewNoiseless, NewPurePredictorVanilla, NewSRIF,
    NewSquareRoot, NewVanilla, Noiseless, ReplayNoise, SRIF, SquareRoot, VanLoan, VanLoanBatch, Vanilla,
)
