"""CPU oracle for the embedding-metric hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package never imports it; the product path
fails loudly when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * simd_dot / Compare_Face_* (Feature.hpp:273-392,763-804): PINNED against the
    reference's own C++ compiled from /root/reference into oracle/_ref/ (see Makefile)
    and against the committed golden vectors in tests/golden/ generated from it.
  * negative selection (a2), draw_cos_dis_real.py statistics (a5) and the facenet.py LFW functions: PINNED against the
    reference's own Python code run in the build container (tests/golden/make_golden_py.py -> tests/golden/ref_python.npz).
  * TripletLoss / norm / cosine_dist / EFM (MXNet arithmetic): the reference holds no
    tests or golden vectors and MXNet is not installable here -> "parity unpinned";
    anchored on the reference call sites and hand-computed known-answer vectors.
  * Batch-hard mining: no reference implementation exists -> "parity unpinned".
"""
from .oracle import *  # noqa: F401,F403
