"""scan_rs_b200 -- B200-native normalize -> PCA hot path of scan-rs behind a C ABI.

The package mirrors the reference's module layout for this path only:
  sqz            AdaptiveMat (device-resident counts), LowRankOffset (normalized matrix)
  normalization  Normalization, normalize, normalize_with_size_factor, ...
  dim_red        BkSvd, RandSvd, Irlba, svd_bk, svd_rand, irlba
  snoop          NoOpSnoop, AtomicSnoop
  mtx            load_mtx (gz MatrixMarket -> device matrix)
  nn             knn, find_nn on the PCA scores
  synth          synthetic workloads (test / bench utility)
All compute runs in scan_rs_b200/libscanb200.so (hand-written sm_100a CUDA + cuSOLVER/cuBLAS for the
small dense steps); there is no CPU fallback."""
from ._lib import CancellationError, ScanB200Error, LIB_PATH  # noqa: F401
from .sqz import AdaptiveMat, Context, LowRankOffset  # noqa: F401
from .normalization import (LogBase, Normalization, binom_deviance_resid, binom_pearson_resid,  # noqa: F401
                            log1p_normalize_fixed_point, log_normalize, log_normalize_with_size_factor,
                            normalize, normalize_with_size_factor)
from .dim_red import BkSvd, Irlba, RandSvd, irlba, irlba_start, omega, pinned_outputs, svd_bk, svd_rand, variance_explained  # noqa: F401
ScanError = ScanB200Error
from .snoop import AtomicSnoop, NoOpSnoop  # noqa: F401
from .mtx import load_mtx  # noqa: F401
from .nn import find_nn, knn  # noqa: F401
from .multi import MultiContext  # noqa: F401
