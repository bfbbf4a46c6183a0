"""B200-native (sm_100a) embedding-metric hot path: triplet-loss head, EFM/MFM activation, cosine
verification and gallery scoring -- Python mirror of the reference's operator interface over libefmtl.so.

There is no CPU fallback: importing the package loads (or builds) the CUDA library and fails loudly otherwise.
"""
from . import _lib
from ._lib import EfmError, last_error
from .activation import EFM, efm, efm3, mfm2
from .cosine import (calc_counts, calc_lines, cosine_allpairs, cosine_dist, cosine_pairs, cosine_pairs_indexed,
                     read_csv, write_cosine_csv)
from .features import (acc, extract_features, format_feature_rows, format_label_rows, write_feature_csv,
                       write_label_csv)
from . import batches, formats, verification
from .batches import TripletPairIter, define_pos, pk_batches
from .gallery import (Compare_Face_DB, Compare_Face_From_DB, Compare_Face_Person, Gallery, merge_topk, simd_dot)
from .loss import (BatchHardStep, BatchHardTripletLoss, TripletLoss, batch_hard_backward, batch_hard_grad_rows, batch_hard_triplet,
                   sample_negatives, triplet_loss)
from .norm import frob_normalize, l2_normalize, norm, row_norms

__all__ = [
    "EFM", "efm", "efm3", "mfm2", "TripletLoss", "triplet_loss", "BatchHardTripletLoss", "BatchHardStep", "batch_hard_triplet",
    "batch_hard_backward", "batch_hard_grad_rows", "sample_negatives", "norm", "l2_normalize", "frob_normalize", "row_norms", "cosine_dist",
    "cosine_pairs", "cosine_pairs_indexed", "cosine_allpairs", "write_cosine_csv", "read_csv", "calc_counts",
    "calc_lines", "Gallery", "merge_topk", "simd_dot", "Compare_Face_From_DB", "Compare_Face_Person",
    "Compare_Face_DB", "extract_features", "acc", "format_feature_rows", "format_label_rows", "write_feature_csv",
    "write_label_csv", "formats", "verification", "batches", "TripletPairIter", "define_pos", "pk_batches", "EfmError",
    "last_error",
]
