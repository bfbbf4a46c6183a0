"""On-disk artefacts of the reference, byte-compatible: the callers and data formats either side of the
hot path, so that galleries / training sets written by the reference can be consumed and vice versa.

* sqlite register DB `reg_face(IDS INTEGER PRIMARY KEY AUTOINCREMENT, NAME TEXT, FEATEXE BLOB, SIMD_DOT REAL)`
  (Feature.hpp:431-449 `Create_Register_DB`, :393-419 `Select_Registed_Data`, :453-489 `Register_DB_File`):
  FEATEXE = the D fp32 values of the feature vector in native (little-endian) byte order, SIMD_DOT = sqrt(simd_dot(fv, fv)),
  NAME = the second-to-last component of the image path (the person's directory).
* `feature_vector.txt` lines `[f, f, ...]<2 chars>id` and the `train_img.csv` / `train_id.txt` / `test_img.csv` /
  `test_id.txt` split of feature_extraction/c_version/slice_celeb1m.py:17-80 (the CSVIter inputs of
  pre-trained_efm_v3.py:155).

Host-side I/O only (sqlite3 / text); everything numeric that the reference computes at registration time (the stored
norm, the duplicate check through Compare_Face_From_DB) goes through the CUDA library -- there is no CPU compute path here.
"""
from __future__ import annotations

import sqlite3
from collections import OrderedDict

import numpy as np

CREATE_REG_FACE = ("CREATE TABLE `reg_face` ( `IDS` INTEGER PRIMARY KEY AUTOINCREMENT,  `NAME`  TEXT,  `FEATEXE` BLOB, "
                   "`SIMD_DOT` REAL);")                      # Feature.hpp:437, verbatim schema
INSERT_REG_FACE = "INSERT INTO `reg_face` (`NAME`,`FEATEXE`,`SIMD_DOT`) values (?,?,?);"      # Feature.hpp:473
SELECT_REG_FACE = "SELECT * FROM  `reg_face`;"               # Feature.hpp:399


def create_register_db(path: str):
    """Create_Register_DB (Feature.hpp:431-449): a fresh DB file with the `reg_face` table."""
    con = sqlite3.connect(path)
    try:
        con.execute(CREATE_REG_FACE)
        con.commit()
    finally:
        con.close()


def registered_name(image_path: str) -> str:
    """Register_DB_File names a row after the image's directory: split on '/', take element size-2 (Feature.hpp:474-475)."""
    parts = image_path.split("/")
    if len(parts) < 2:
        raise ValueError("image path %r has no directory component (the reference indexes [size - 2])" % image_path)
    return parts[-2]


def insert_registered(path: str, names, features, simd_norms):
    """Append rows (NAME, FEATEXE, SIMD_DOT) exactly as Register_DB_File binds them (Feature.hpp:473-483): text, a blob
    of D native fp32 values, a double holding the fp32 norm.  `features` (N, D) float32 array-like, `simd_norms` (N,)."""
    F = np.ascontiguousarray(np.asarray(features, dtype=np.float32))
    nr = np.asarray(simd_norms, dtype=np.float32).reshape(-1)
    if F.ndim != 2 or len(names) != F.shape[0] or nr.shape[0] != F.shape[0]:
        raise ValueError("names / features / simd_norms disagree: %d, %s, %d" % (len(names), F.shape, nr.shape[0]))
    con = sqlite3.connect(path)
    try:
        con.executemany(INSERT_REG_FACE, [(str(n), F[i].tobytes(), float(nr[i])) for i, n in enumerate(names)])
        con.commit()
    finally:
        con.close()


def select_registered(path: str, fv_size: int):
    """Select_Registed_Data (Feature.hpp:393-419): every row of `reg_face` in table order -> (ids int32 (N,), names list,
    G float32 (N, fv_size), gnorm float32 (N,)).  Like the reference's memcpy, only the first fv_size floats of a blob are
    used; a shorter blob is an error here (the reference would read past its end)."""
    con = sqlite3.connect(path)
    try:
        rows = con.execute(SELECT_REG_FACE).fetchall()
    finally:
        con.close()
    ids = np.empty(len(rows), dtype=np.int32)
    names = []
    G = np.empty((len(rows), fv_size), dtype=np.float32)
    gnorm = np.empty(len(rows), dtype=np.float32)
    for i, (pid, name, blob, sd) in enumerate(rows):
        if blob is None or len(blob) < 4 * fv_size:
            raise ValueError("reg_face row %r holds %d bytes, need %d" % (pid, 0 if blob is None else len(blob), 4 * fv_size))
        ids[i] = pid
        names.append(name)
        G[i] = np.frombuffer(blob, dtype=np.float32, count=fv_size)
        gnorm[i] = np.float32(sd)                           # (float)sqlite3_column_double (Feature.hpp:408)
    return ids, names, G, gnorm


def load_gallery(path: str, fv_size: int, device="cuda"):
    """The register DB as a device-resident `Gallery` with the STORED norms (Feature.hpp:412) + (ids, names)."""
    import torch

    from .gallery import Gallery
    ids, names, G, gnorm = select_registered(path, fv_size)
    gal = Gallery(torch.from_numpy(G).to(device), gnorm=torch.from_numpy(gnorm).to(device),
                  person_ids=torch.from_numpy(ids).to(device))
    return gal, ids, names


def register_db_file(path: str, image_paths, features, fv_size=None, dedupe_th=0.99999):
    """Register_DB_File (Feature.hpp:453-489) on device tensors: rows whose best cosine against the members ALREADY in the
    DB (loaded once, before the loop, exactly like `AllMembers`) reaches `dedupe_th` are skipped; the others are appended
    with NAME = directory of the image and SIMD_DOT = sqrt(simd_dot(fv, fv)) in the reference's summation order (computed
    by the library, bit-identical to the SSE code).  `features`: (N, D) float32 CUDA tensor.  Returns the number of rows
    written."""
    import torch

    from . import _lib as L
    from .gallery import Gallery
    L.require_cuda(features)
    F = features.to(torch.float32).contiguous()
    D = F.shape[1] if fv_size is None else int(fv_size)
    ids, _, G, gnorm = select_registered(path, D)
    keep = np.ones(F.shape[0], dtype=bool)
    if len(ids):
        gal = Gallery(torch.from_numpy(G).to(F.device), gnorm=torch.from_numpy(gnorm).to(F.device))
        idx, _ = gal.argmax(F[:, :D].contiguous(), dedupe_th)
        keep = idx.cpu().numpy() < 0                         # Compare_Face_From_DB(...) >= 0 -> "same feature vector", skipped
    norms = Gallery(F[:, :D].contiguous()).gnorm.cpu().numpy()
    rows = np.flatnonzero(keep)
    insert_registered(path, [registered_name(image_paths[i]) for i in rows], F[:, :D].cpu().numpy()[rows], norms[rows])
    return int(rows.size)


# ------------------------------------------------------------------------------------------------ slice_celeb1m.py
def read_fv(path: str):
    """slice_celeb1m.py:17-32: `[f, f, ...]<2 chars>id\\n` per line -> OrderedDict id -> list of feature lists (file order
    of first appearance, which is what the train/test split iterates over)."""
    fv_id = OrderedDict()
    with open(path, "r") as f:
        for line in f:
            parts = line.split("]")
            fv = [float(e) for e in parts[0][1:].split(",")]
            key = parts[1][2:-1]
            fv_id.setdefault(key, []).append(fv)
    return fv_id


def format_fv_line(fv, ident) -> str:
    """A line `read_fv` (and the reference's reader) parses back: Python list repr, two separator characters, the id."""
    return "{}, {}\n".format([float(v) for v in fv], ident)


def split_train_test(fv_id, train_portion: float):
    """slice_celeb1m.py:56-72: identities are visited in dict order; identity number i (0-based) goes to the training set
    while i < total_ids * train_portion.  Returns (train_imgs, train_ids, test_imgs, test_ids)."""
    limit = len(fv_id) * train_portion
    tr_i, tr_y, te_i, te_y = [], [], [], []
    for i, key in enumerate(fv_id):
        imgs, ys = (tr_i, tr_y) if i < limit else (te_i, te_y)
        for fv in fv_id[key]:
            ys.append(key)
            imgs.append(fv)
    return tr_i, tr_y, te_i, te_y


def write_img_csv(path: str, rows):
    """write2file(is_img=True), slice_celeb1m.py:34-42: '{},' per element, '{}\\n' for the last (no trailing comma)."""
    with open(path, "w") as f:
        for fv in rows:
            f.write(",".join("{}".format(v) for v in fv) + "\n")


def write_id_txt(path: str, ids):
    """write2file(is_img=False), slice_celeb1m.py:43-45: one id per line."""
    with open(path, "w") as f:
        for line in ids:
            f.write("{}\n".format(line))
