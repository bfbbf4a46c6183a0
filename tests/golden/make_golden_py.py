#!/usr/bin/env python
"""Generates tests/golden/ref_python.npz by RUNNING THE REFERENCE'S OWN PYTHON FUNCTIONS in this container.

The reference's scripts cannot run as a whole here (MXNet / TensorFlow / matplotlib are not installable), but three of
its modules hold plain numpy / Python functions whose only obstacle is a module-level import of an absent package:

  * draw_cos_dis_real.py            calc_counts, read_csv, calc_lines                      
        obstacles: `import matplotlib.pyplot`, `np.float` (removed from numpy >= 1.24)
  * feature_extraction/c_version/slice_celeb1m.py   read_fv, write2file, main()             
  * feature_extraction/facenet_version/facenet.py   distance, calculate_accuracy, calculate_roc, calculate_val,
        calculate_val_far                                                                    (LFW protocol)
        obstacles: `import tensorflow`, `tensorflow.python.*`, `scipy.misc`
  * train_efm.py                    define_pos, Batch, DataIter (the 2B-row batch layout)  
        obstacles: `import mxnet` (used here only as the array CARRIER), module-level training code after the definitions

This script puts EMPTY stand-in modules for exactly those absent packages into sys.modules (none of the functions
above touches them), restores `np.float = float`, imports the reference modules from /root/reference UNMODIFIED, calls the
functions on seeded inputs and stores inputs + outputs.  No reference source is copied.  The vectors pin the oracle /
host restatements of these rows against the reference itself (tests/test_oracle_cpu.py, tests/test_formats_cpu.py,
tests/test_verification_cpu.py).

Run (needs /root/reference):  python tests/golden/make_golden_py.py
"""
import importlib.util
import io
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("EFM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def stub(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not found at %s" % REF)
    # ---- stand-ins for the absent packages (never called by the functions we run)
    for n in ("matplotlib", "matplotlib.pyplot", "tensorflow", "tensorflow.python", "tensorflow.python.training",
              "tensorflow.python.training.training", "tensorflow.python.platform", "tensorflow.python.platform.gfile"):
        if n not in sys.modules:
            stub(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["tensorflow.python.training"].training = sys.modules["tensorflow.python.training.training"]
    sys.modules["tensorflow.python.platform"].gfile = sys.modules["tensorflow.python.platform.gfile"]
    import scipy
    if not hasattr(scipy, "misc"):
        scipy.misc = stub("scipy.misc")
    elif "scipy.misc" not in sys.modules:
        sys.modules["scipy.misc"] = scipy.misc
    if not hasattr(np, "float"):
        np.float = float                                     # the shim SURVEY 8a5 names

    out = {}
    rng = np.random.default_rng(20261018)

    # ------------------------------------------------------------------ a5: draw_cos_dis_real.py
    dcr = load(os.path.join(REF, "draw_cos_dis_real.py"), "ref_draw_cos_dis_real")
    n_rows = 1000
    pos = np.clip(0.7 + 0.15 * rng.standard_normal(n_rows), -1, 1).astype(np.float32)
    neg = np.clip(0.1 + 0.25 * rng.standard_normal(n_rows), -1, 1).astype(np.float32)
    import csv
    buf = io.StringIO()
    w = csv.writer(buf, delimiter=" ")                      # how train_efm.py:252-255 writes the file
    for a, b in zip(pos.tolist(), neg.tolist()):
        w.writerow([a, b])
    csv_text = buf.getvalue().replace("\r\n", "\n")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            with open("cosine_similarity.csv", "w", newline="") as f:
                f.write(csv_text)
            for ep in (1, 5, 7):
                p_list, n_list = dcr.read_csv(desire_epoch=ep)
                out["a5_read_pos_ep%d" % ep] = np.array(p_list, dtype=np.float64)
                out["a5_read_neg_ep%d" % ep] = np.array(n_list, dtype=np.float64)
            p_list, n_list = dcr.read_csv(desire_epoch=5)
        finally:
            os.chdir(cwd)
    out["a5_csv_text"] = np.frombuffer(csv_text.encode(), dtype=np.uint8)
    cnt, edges = dcr.calc_counts(p_list)
    out["a5_counts_pos"] = cnt; out["a5_edges_pos"] = edges
    names = ["pos_line", "neg_line", "pos_cdf", "neg_cdf", "pos_bins", "neg_bins"]
    for k, v in zip(names, dcr.calc_lines(pos_dis=p_list, neg_dis=n_list)):
        out["a5_" + k] = np.asarray(v)

    # ------------------------------------------------------------------ 8f-2: slice_celeb1m.py
    sc = load(os.path.join(REF, "feature_extraction", "c_version", "slice_celeb1m.py"), "ref_slice_celeb1m")
    ids = ["m.0%dx" % (i % 9) for i in range(60)]
    rows = rng.standard_normal((60, 6)).astype(np.float32)
    fv_text = "".join("{}, {}\n".format([float(v) for v in fv], ident) for fv, ident in zip(rows, ids))
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "feature_vector.txt"), "w") as f:
            f.write(fv_text)
        fv_id = sc.read_fv(os.path.join(td, "feature_vector.txt"))
        out["fv_keys"] = np.array(list(fv_id.keys()))
        out["fv_counts"] = np.array([len(fv_id[k]) for k in fv_id])
        out["fv_first"] = np.array([fv_id[k][0] for k in fv_id], dtype=np.float64)
        os.makedirs(os.path.join(td, "out"))
        argv = sys.argv
        sys.argv = ["slice_celeb1m.py", "--root2files", td, "--output_root", os.path.join(td, "out"), "--train_portion", "0.6"]
        stdout = sys.stdout
        sys.stdout = io.StringIO()
        try:
            sc.main()
        finally:
            sys.argv = argv
            sys.stdout = stdout
        for fn in ("train_img.csv", "train_id.txt", "test_img.csv", "test_id.txt"):
            with open(os.path.join(td, "out", fn), "rb") as f:
                out["fv_out_" + fn.replace(".", "_")] = np.frombuffer(f.read(), dtype=np.uint8)
    out["fv_text"] = np.frombuffer(fv_text.encode(), dtype=np.uint8)
    out["fv_train_portion"] = np.array([0.6])

    # ------------------------------------------------------------------ 8f-4: facenet.py (LFW protocol)
    fn = load(os.path.join(REF, "feature_extraction", "facenet_version", "facenet.py"), "ref_facenet")
    P, D = 600, 32                                           # 600 pairs, 10 folds of 60 (LFW: 6000 pairs, 10 folds of 600)
    cent = rng.standard_normal((P, D))
    issame = (np.arange(P) % 2 == 0)
    e1 = cent + 0.9 * rng.standard_normal((P, D))                      # hard enough for a non-trivial ROC
    e2 = np.where(issame[:, None], cent + 0.9 * rng.standard_normal((P, D)), rng.standard_normal((P, D)))
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True); e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
    out["lfw_e1"] = e1; out["lfw_e2"] = e2; out["lfw_issame"] = issame
    for metric in (0, 1):
        out["lfw_dist_m%d" % metric] = fn.distance(e1, e2, metric)
        thr = np.arange(0, 4, 0.01) if metric == 0 else np.arange(0, 1, 0.0025)
        out["lfw_thr_m%d" % metric] = thr
        for sm in (False, True):
            tpr, fpr, acc = fn.calculate_roc(thr, e1, e2, issame, nrof_folds=10, distance_metric=metric, subtract_mean=sm)
            tag = "m%d_sm%d" % (metric, int(sm))
            out["lfw_tpr_" + tag] = tpr; out["lfw_fpr_" + tag] = fpr; out["lfw_acc_" + tag] = acc
            try:
                val, val_std, far = fn.calculate_val(thr, e1, e2, issame, 1e-2, nrof_folds=10, distance_metric=metric, subtract_mean=sm)
                out["lfw_val_" + tag] = np.array([val, val_std, far])
            except ValueError as e:
                # scipy >= 1.10: interp1d(kind='slinear') refuses duplicate x, and the FAR curve of a threshold sweep always
                # has duplicates -> the reference's calculate_val cannot run on this scipy; recorded, not pinned
                out["lfw_val_error"] = np.frombuffer(("%s: %s" % (type(e).__name__, e)).encode(), dtype=np.uint8)
    d0 = out["lfw_dist_m0"]
    out["lfw_acc_at_1"] = np.array(fn.calculate_accuracy(1.0, d0, issame))
    out["lfw_valfar_at_1"] = np.array(fn.calculate_val_far(1.0, d0, issame))

    # ------------------------------------------------------------------ 8f-1: define_pos + DataIter (train_efm.py:37-114)
    # train_efm.py trains at module level, so only its definitions (everything before the "Data & Parameter setting" block)
    # are executed, read from the reference tree at run time, against a CARRIER-ONLY stand-in for mxnet: an NDArray that
    # wraps numpy and offers the four carrier methods these definitions use (indexing, asnumpy, asscalar, shape) -- no
    # arithmetic of the reference goes through it.
    class FakeND(object):
        def __init__(self, a):
            self.a = np.asarray(a, dtype=np.float32)
        def __getitem__(self, i):
            return FakeND(self.a[i])
        def asnumpy(self):
            return self.a
        def asscalar(self):
            return self.a.reshape(-1)[0]
        @property
        def shape(self):
            return self.a.shape
        def __len__(self):
            return len(self.a)
    mx = stub("mxnet")
    mx.nd = stub("mxnet.nd"); mx.nd.array = lambda x: FakeND(np.array([e.a if isinstance(e, FakeND) else e for e in x]) if isinstance(x, list) else x)
    mx.io = stub("mxnet.io")
    class _DataIterBase(object):
        def __init__(self, *a, **k):
            pass
    mx.io.DataIter = _DataIterBase
    for n in ("mxnet.autograd", "mxnet.gluon", "mxnet.init", "mxnet.gluon.nn", "lightcnn", "gluoncv"):
        stub(n)
    mx.autograd = sys.modules["mxnet.autograd"]; mx.gluon = sys.modules["mxnet.gluon"]; mx.init = sys.modules["mxnet.init"]
    mx.gluon.nn = sys.modules["mxnet.gluon.nn"]
    src = open(os.path.join(REF, "train_efm.py")).read()
    cut = src.index('""" Data & Parameter setting """')
    ns = {"__name__": "ref_train_efm_defs"}
    exec(compile(src[:cut], os.path.join(REF, "train_efm.py"), "exec"), ns)
    Bsz, nb = 4, 6
    lab = rng.integers(0, 7, Bsz * nb).astype(np.float32)
    imgs = rng.random((Bsz * nb, 1, 2, 3)).astype(np.float32)
    class _B(object):
        def __init__(self, d, l):
            self.data = [FakeND(d)]; self.label = [FakeND(l)]
    batches = [_B(imgs[i * Bsz:(i + 1) * Bsz], lab[i * Bsz:(i + 1) * Bsz]) for i in range(nb)]
    pos_img = ns["define_pos"](batches, nb, Bsz)
    out["f1_imgs"] = imgs; out["f1_labels"] = lab
    out["f1_pos_keys"] = np.array(list(pos_img.keys()), dtype=np.int64)
    out["f1_pos_imgs"] = np.stack([pos_img[k].asnumpy() for k in pos_img])
    it = ns["DataIter"](batches, nb, pos_img, Bsz, (Bsz, 1, 2, 3))
    got_d, got_l = [], []
    for b in it:
        got_d.append(b.data[0].asnumpy()); got_l.append(b.label[0].asnumpy())
    out["f1_batches_data"] = np.stack(got_d); out["f1_batches_label"] = np.stack(got_l)

    # ------------------------------------------------------------------ a2: negative selection (train_efm.py:234-238, :268-272)
    # The rejection-sampling loops live inline in the training / validation loops.  Their source lines are read from the
    # reference tree at run time, dedented and executed with carrier stand-ins: `label` a FakeND, `fc` a FakeND whose row j
    # holds the number j (so the appended rows reveal the drawn indices), Python's module-level `random` seeded.
    import random
    import textwrap
    lines = src.splitlines()
    def loop_after(marker_line_no):
        k = next(i for i in range(marker_line_no, len(lines)) if lines[i].strip().startswith("for i in range(batch_size):"))
        end = next(i for i in range(k + 1, len(lines)) if "neg.append" in lines[i])
        return textwrap.dedent("\n".join(lines[k:end + 1]))
    train_loop = loop_after(next(i for i, l in enumerate(lines) if '""" training loop """' in l))
    val_loop = loop_after(next(i for i, l in enumerate(lines) if '""" validation loop """' in l))
    assert "batch_size - 1" in train_loop and "batch_size*2 - 1" in val_loop     # pools: anchor half (train), both halves (val)
    Bn = 16
    lab2 = np.concatenate([rng.integers(0, 5, Bn)] * 2).astype(np.float32)    # duplicated halves like the loader emits
    out["a2_labels"] = lab2
    for tag, code in (("train", train_loop), ("val", val_loop)):
        for seed in (0, 1, 12345):
            random.seed(seed)
            env = {"random": random, "batch_size": Bn, "label": FakeND(lab2), "fc": FakeND(np.arange(2 * Bn).reshape(-1, 1)), "neg": []}
            exec(code, env)
            out["a2_%s_seed%d" % (tag, seed)] = np.array([int(r[0]) for r in env["neg"]], dtype=np.int64)

    # ------------------------------------------------------------------ a7: feature / label export text (extract_feacture_v2.py:67-79)
    # The two `with open(...)` blocks of the training loop, read from the reference tree and executed in a scratch directory.
    # What is pinned is the TEXT FORMAT ('{},' per element with a trailing comma, numpy-float32 label repr); the normalisation
    # itself runs through a numpy-float32 stand-in for `mx.nd.norm` / NDArray division and its result is stored next to
    # the text, so the test formats exactly these values.
    class FakeND2(FakeND):
        def __truediv__(self, o):
            return FakeND2(self.a / (o.a if isinstance(o, FakeND) else o))
        def __getitem__(self, i):
            return FakeND2(self.a[i])
    mx.nd.norm = lambda x: FakeND2(np.sqrt(np.sum(np.square(x.a), dtype=np.float32), dtype=np.float32).reshape(1))
    esrc = open(os.path.join(REF, "extract_feacture_v2.py")).read().splitlines()
    k0 = next(i for i, l in enumerate(esrc) if '""" save feature vectors """' in l)
    k1 = next(i for i in range(k0, len(esrc)) if esrc[i].strip().startswith("print("))
    block = textwrap.dedent("\n".join(esrc[k0:k1]))
    fc = rng.standard_normal((5, 7)).astype(np.float32)
    labv = np.array([12, 0, 3, 78578, 7], dtype=np.float32)
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            exec(block, {"mx": mx, "fc": FakeND2(fc), "label": FakeND2(labv), "batch_size": 5})
            out["a7_feature_csv"] = np.frombuffer(open("feature_vector_train.csv", "rb").read(), dtype=np.uint8)
            out["a7_label_csv"] = np.frombuffer(open("label_train.csv", "rb").read(), dtype=np.uint8)
        finally:
            os.chdir(cwd)
    out["a7_fc_normalised"] = np.stack([fc[v] / np.sqrt(np.sum(np.square(fc[v]), dtype=np.float32), dtype=np.float32) for v in range(5)])
    out["a7_labels"] = labv

    # ------------------------------------------------------------------ a4: cosine_similarity.csv rows (train_efm.py:250-255)
    # The block that appends "pos neg" rows, read from the reference tree; `cosine_dist` is replaced by a stand-in that
    # returns the given similarities as (1,) carriers whose asscalar() is a numpy.float32 -- as MXNet's is, which is what
    # decides the text (csv.writer stringifies a numpy.float32 with str(): shortest float32 repr).
    k0 = next(i for i, l in enumerate(lines) if "save positive distance and negative distance to csv" in l)
    k1 = next(i for i in range(k0, len(lines)) if "csvwriter.writerow" in lines[i])
    block = textwrap.dedent("\n".join(lines[k0:k1 + 1]))
    pd_ = np.array([0.70710678, 1.0, -0.1, 1e-8, 0.33333334], dtype=np.float32)
    nd_ = np.array([-0.25, 0.5, 0.123456789, -1.0, 3e-5], dtype=np.float32)
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            exec(block, {"csv": csv, "batch_size": 5, "anc": None, "pos": None, "neg": None,
                         "cosine_dist": lambda a, p, n, b: ([FakeND(pd_[i:i + 1]) for i in range(b)], [FakeND(nd_[i:i + 1]) for i in range(b)])})
            out["a4_csv_bytes"] = np.frombuffer(open("cosine_similarity.csv", "rb").read(), dtype=np.uint8)
        finally:
            os.chdir(cwd)
    out["a4_pos"] = pd_; out["a4_neg"] = nd_

    np.savez_compressed(os.path.join(OUT, "ref_python.npz"), **out)
    print("wrote", os.path.join(OUT, "ref_python.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
