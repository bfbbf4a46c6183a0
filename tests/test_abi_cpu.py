"""CPU tests of the drop-in boundary: libefmtl.so loads without a GPU and exports every symbol that
include/efmtl.h declares; argument validation (which runs before any CUDA call) follows the documented error codes."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from improving_face_recognition_performance_using_triplet_loss_b200 import _lib
    return _lib


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "efmtl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(efm_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(L):
    syms = declared_symbols()
    assert len(syms) >= 30
    raw = ctypes.CDLL(L.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), "libefmtl.so does not export %s" % s
    assert set(syms) == set(L.SIGNATURES), "ctypes table and header disagree: %s" % (set(syms) ^ set(L.SIGNATURES))


def test_abi_version_and_error_string(L):
    assert L.lib.efm_abi_version() == 2
    rc = L.lib.efm_act_fwd(7, None, None, 1, 3, 1, 0, None)            # bad mode: rejected before any CUDA call
    assert rc == -1 and "mode" in L.last_error()
    rc = L.lib.efm_act_fwd(3, None, None, 1, 4, 1, 0, None)            # C % 3 != 0
    assert rc == -1 and "divisible" in L.last_error()
    assert L.lib.efm_act_fwd(3, None, None, 0, 3, 1, 0, None) == 0      # empty batch is a no-op
    rc = L.lib.efm_gallery_topk(None, None, 10, 8, None, 1, 99, 0.0, 0, 0, 0, None, None, None, 0, None)
    assert rc == -1 and "K" in L.last_error()
    rc = L.lib.efm_bh_triplet_fwd_bwd(None, None, 8, 4, 8, 16, 1, 0.2, None, 0, None, None, None, None, None, 0, None)
    assert rc == -1 and "bad shape" in L.last_error()


def test_workspace_sizes_are_consistent(L):
    a = L.lib.efm_bh_workspace_bytes(8192, 8192, 256, L.BF16, L.BH_NORMALIZE)
    b = L.lib.efm_bh_workspace_bytes(8192, 65536, 256, L.BF16, L.BH_NORMALIZE)
    assert 0 < a < b and a % 256 == 0
    assert a >= 8192 * 256 * 2                                          # holds the bf16 mining rows
    assert L.lib.efm_bh_workspace_bytes(8192, 8192, 256, L.F32, 0) < a  # fp32 path without normalise mines E in place
    assert L.lib.efm_gallery_topk_workspace_bytes(1000, 512, 4, 5, L.F32) >= 4 * 1000 * 4
    assert L.lib.efm_cosine_allpairs_workspace_bytes(100, 200, 342, L.BF16) >= (100 + 200) * 384 * 2


def test_operators_refuse_cpu_tensors():
    import improving_face_recognition_performance_using_triplet_loss_b200 as P
    x = torch.zeros(4, 6)
    for fn in (P.efm3, P.mfm2, P.l2_normalize, P.norm, P.frob_normalize):
        with pytest.raises(ValueError, match="CUDA"):
            fn(x)
    with pytest.raises(ValueError, match="CUDA"):
        P.TripletLoss(0.2)(x, x, x)
    with pytest.raises(ValueError, match="CUDA"):
        P.batch_hard_triplet(x, torch.zeros(4, dtype=torch.int32), 0.2)
    with pytest.raises(ValueError, match="CUDA"):
        P.Gallery(x)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "improving_face_recognition_performance_using_triplet_loss_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_host_side_formats_and_sampler():
    import random

    import numpy as np

    import improving_face_recognition_performance_using_triplet_loss_b200 as P
    import oracle as O
    lab = torch.tensor([0, 0, 1, 1, 2, 2, 0, 1] * 2, dtype=torch.float32)
    got = P.sample_negatives(lab, 8, "train", random.Random(3)).numpy()
    assert np.array_equal(got, O.sample_negatives(lab.numpy(), 8, "train", random.Random(3)))
    fc = torch.tensor([[0.6, 0.8, 0.0]])
    assert P.format_feature_rows(fc) == "0.6000000238418579,0.800000011920929,0.0,\n"
    assert P.format_label_rows(torch.tensor([12.0])) == "12.0\n"
