"""On-disk formats of the reference: byte-level checks of the sqlite register DB and the
slice_celeb1m.py text files.  Pure host I/O -- no GPU needed; the parts of registration that compute (stored norm,
duplicate check) are GPU tests in test_gpu_parity.py."""
import os
import sqlite3
import struct

import numpy as np
import pytest

import oracle as O


@pytest.fixture(scope="module")
def F():
    from improving_face_recognition_performance_using_triplet_loss_b200 import formats
    return formats


def test_reg_face_schema_and_blob_layout(F, tmp_path):
    db = str(tmp_path / "reg.db")
    F.create_register_db(db)
    con = sqlite3.connect(db)
    sql = con.execute("SELECT sql FROM sqlite_master WHERE name='reg_face'").fetchone()[0]
    con.close()
    assert sql == F.CREATE_REG_FACE.rstrip(";")             # the reference's CREATE TABLE text (Feature.hpp:437)
    rng = np.random.default_rng(0)
    G = rng.standard_normal((5, 342)).astype(np.float32)
    norms = np.array([np.sqrt(np.float32(O.simd_dot(g, g))) for g in G], dtype=np.float32)   # sqrt(simd_dot(fv, fv)), Feature.hpp:476
    names = [F.registered_name("/data/celeb/%s/img_%d.jpg" % (n, i)) for i, n in enumerate(["ann", "bob", "ann", "cy", "dee"])]
    assert names == ["ann", "bob", "ann", "cy", "dee"]
    F.insert_registered(db, names, G, norms)
    con = sqlite3.connect(db)
    rows = con.execute("SELECT IDS, NAME, FEATEXE, SIMD_DOT FROM reg_face ORDER BY IDS").fetchall()
    con.close()
    assert [r[0] for r in rows] == [1, 2, 3, 4, 5]          # AUTOINCREMENT ids start at 1
    for r, g, n in zip(rows, G, norms):
        assert len(r[2]) == 342 * 4 and struct.unpack("<342f", r[2]) == tuple(g.tolist())   # native little-endian fp32
        assert isinstance(r[3], float) and np.float32(r[3]) == n                            # double holding the fp32 norm
    ids, names2, G2, gn2 = F.select_registered(db, 342)
    assert ids.tolist() == [1, 2, 3, 4, 5] and names2 == names
    assert np.array_equal(G2, G) and np.array_equal(gn2, norms)
    # fvSize smaller than the blob: the reference memcpy's the first fvSize floats (Feature.hpp:411)
    _, _, G3, _ = F.select_registered(db, 100)
    assert np.array_equal(G3, G[:, :100])
    with pytest.raises(ValueError):
        F.select_registered(db, 343)
    with pytest.raises(ValueError):
        F.registered_name("img.jpg")


def _reference_read_fv(path):
    """The reference's own parser, restated (slice_celeb1m.py:17-32), to cross-check ours."""
    fv_id = {}
    with open(path, "r") as f:
        for line in f:
            fv = []
            line = line.split(']')
            line[0] = line[0][1:]
            line[1] = line[1][2:-1]
            for ele in line[0].split(","):
                fv.append(float(ele))
            fv_id.setdefault(line[1], []).append(fv)
    return fv_id


def test_feature_vector_txt_and_split(F, tmp_path):
    rng = np.random.default_rng(1)
    ids = ["m.0%d" % (i % 7) for i in range(40)]
    rows = rng.standard_normal((40, 8)).astype(np.float32)
    p = str(tmp_path / "feature_vector.txt")
    with open(p, "w") as f:
        for fv, ident in zip(rows, ids):
            f.write(F.format_fv_line(fv, ident))
    got = F.read_fv(p)
    ref = _reference_read_fv(p)
    assert list(got.keys()) == list(ref.keys()) and all(got[k] == ref[k] for k in ref)
    assert got["m.03"][0] == [float(v) for v in rows[3]]
    tr_i, tr_y, te_i, te_y = F.split_train_test(got, 0.5)
    # 7 identities, limit 3.5 -> identities number 0..3 train, 4..6 test (slice_celeb1m.py:63-72)
    assert sorted(set(tr_y)) == ["m.00", "m.01", "m.02", "m.03"] and sorted(set(te_y)) == ["m.04", "m.05", "m.06"]
    assert len(tr_i) == len(tr_y) and len(te_i) == len(te_y) and len(tr_i) + len(te_i) == 40
    F.write_img_csv(str(tmp_path / "train_img.csv"), tr_i)
    F.write_id_txt(str(tmp_path / "train_id.txt"), tr_y)
    lines = open(str(tmp_path / "train_img.csv")).read().splitlines()
    assert len(lines) == len(tr_i) and not lines[0].endswith(",")
    assert lines[0] == ",".join("{}".format(v) for v in tr_i[0])          # '{},' ... '{}\n' (slice_celeb1m.py:37-41)
    assert np.allclose(np.loadtxt(str(tmp_path / "train_img.csv"), delimiter=",", dtype=np.float64), np.array(tr_i))
    assert open(str(tmp_path / "train_id.txt")).read().splitlines() == tr_y
