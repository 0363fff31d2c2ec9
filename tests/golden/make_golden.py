#!/usr/bin/env python
"""Generates tests/golden/ref_fixture.db and tests/golden/ref_golden.npz by running the REAL
reference code (/root/reference, read-only, only present in the build container).

What is pinned: the non-faiss half of the hot path exactly as the reference executes it —
schema DDL (main/create_db.py:49-85), blob format (vector_scripts/create_vector_base.py:142-145),
join order / row order / skip-on-decode-error / concat (main/create_index.py:115-189), offsets
(:236-249, :301-313), query-type ordering, concat and mean (main/search_from_image.py:256-317)
and offset -> path mapping (:346-379).  faiss itself is absent (un-vendored faiss_cpu==1.10.0), so
a RECORDING STUB stands in for it: the stub's arithmetic is never used as a golden value.

Run:  python tests/golden/make_golden.py      (then commit the two artefacts)
"""
import os
import pickle
import sqlite3
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
DIMS = {"color": 48, "sift": 128, "dreamsim": 1792}


class _Recorder:
    adds = []
    written = []
    normalized = []


def _install_stubs():
    faiss = types.ModuleType("faiss")

    class _Hnsw:
        efConstruction = 0
        efSearch = 0

    class _Index:
        def __init__(self, *a, **k):
            self.args = a
            self.hnsw = _Hnsw()
            self.is_trained = True
            self.ntotal = 0

        def train(self, x):
            pass

        def add(self, x):
            assert x.dtype == np.float32 and x.flags.c_contiguous
            _Recorder.adds.append(np.array(x))
            self.ntotal += x.shape[0]

    faiss.IndexHNSWFlat = _Index
    faiss.IndexIVFPQ = _Index
    faiss.write_index = lambda index, path: _Recorder.written.append((index.ntotal, path))
    faiss.normalize_L2 = lambda x: _Recorder.normalized.append(np.array(x))   # recorder, no arithmetic
    faiss.read_index = lambda path: (_ for _ in ()).throw(RuntimeError("stub"))
    sys.modules["faiss"] = faiss
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, str(REF))
    # the extractor modules are imported lazily by extract_*_features even on a cache hit; they pull
    # cv2 / dreamsim (absent, out of scope).  Their classes are never constructed on the cache path.
    import vector_scripts  # noqa: F401  (the real package, for create_vector_base)
    for mod, cls in (("create_color_vector", "ColorVectorIndexer"), ("create_sift_vector", "SIFTVLADVectorIndexer"),
                     ("create_dreamsim_vector", "DreamSimVectorIndexer")):
        m = types.ModuleType(f"vector_scripts.{mod}")
        setattr(m, cls, type(cls, (), {}))
        sys.modules[f"vector_scripts.{mod}"] = m


def _vec(rng, d, nonneg=False):
    v = rng.standard_normal(d).astype(np.float32)
    if nonneg:
        v = np.abs(v)
    return (v / np.linalg.norm(v)).astype(np.float32)


def build_fixture_db(db_path: Path):
    """Schema by the reference's own ImageDBCreator; blobs in the reference's writer format."""
    from main.create_db import ImageDBCreator
    import torch
    if db_path.exists():
        db_path.unlink()
    creator = ImageDBCreator.__new__(ImageDBCreator)          # bypass the folder scan; keep create_tables
    creator.db_path = str(db_path)
    creator.journal_mode, creator.synchronous, creator.timeout = "WAL", "OFF", 30000
    creator.create_tables()
    rng = np.random.default_rng(20261018)
    conn = sqlite3.connect(db_path)
    n_images = 30
    for i in range(n_images):
        conn.execute("INSERT INTO images (path) VALUES (?)", (f"image_data/set/img_{i:04d}.jpg",))
    conn.execute("DELETE FROM images WHERE id IN (4, 11)")                   # id gaps
    ids = [r[0] for r in conn.execute("SELECT id FROM images ORDER BY id")]
    blob = lambda v: sqlite3.Binary(pickle.dumps(v, protocol=pickle.HIGHEST_PROTOCOL))  # noqa: E731
    for rid in ids:
        conn.execute("INSERT INTO color_vectors VALUES (?, ?)", (rid, blob(_vec(rng, 48, True))))
        if rid not in (7, 20, 21):                                            # incomplete images
            v = _vec(rng, 128)
            if rid == 9:
                v = v.reshape(1, -1)                                          # 2-D blob: ravel()ed
            conn.execute("INSERT INTO sift_vectors VALUES (?, ?)", (rid, blob(v)))
        v = _vec(rng, 1792)
        if rid == 13:
            payload = blob(torch.from_numpy(v))                               # tensor blob (.cpu())
        elif rid == 17:
            payload = sqlite3.Binary(b"\x80\x05not a pickle")                 # undecodable: skipped
        else:
            payload = blob(v)
        conn.execute("INSERT INTO dreamsim_vectors VALUES (?, ?)", (rid, payload))
    conn.commit()
    conn.execute("PRAGMA wal_checkpoint(TRUNCATE)")
    conn.close()
    for ext in ("-wal", "-shm"):
        p = Path(str(db_path) + ext)
        if p.exists():
            p.unlink()


def main():
    _install_stubs()
    db = HERE / "ref_fixture.db"
    build_fixture_db(db)
    work = HERE / "_work"
    work.mkdir(exist_ok=True)
    os.chdir(work)
    import shutil
    shutil.copy(db, work / "images.db")

    from main.create_index import FAISSIndexBuilderDB
    out = {}
    for types_ in (["color"], ["color", "sift", "dreamsim"], ["dreamsim", "color"]):
        _Recorder.adds.clear()
        b = FAISSIndexBuilderDB(db_path="images.db", vector_types=types_, batch_size=8, log_dir=str(work / "logs"))
        name = "_".join(types_)
        out[f"{name}/count"] = np.array(b._count_records())
        b.build_index(update_index=False)
        out[f"{name}/added"] = np.concatenate(_Recorder.adds, axis=0)
        out[f"{name}/batch_sizes"] = np.array([a.shape[0] for a in _Recorder.adds])
        conn = sqlite3.connect("images.db")
        rows = conn.execute(f"SELECT image_id, offset FROM faiss_index_offsets_{name} ORDER BY offset").fetchall()
        conn.close()
        out[f"{name}/offsets"] = np.array(rows, dtype=np.int64)
        out[f"{name}/index_file"] = np.array(str(b.index_file))

    from main.search_from_image import ImageRecommender
    rec = ImageRecommender(images_root="image_data", db_path="images.db")
    cases = ["color", "sift,color", "dreamsim, COLOR ,sift", "hog,color", "bogus"]
    out["ordered/cases"] = np.array(cases)
    out["ordered/results"] = np.array(["|".join(rec._get_ordered_index_types(c)) for c in cases])
    # query vector (pre-normalisation; normalize_L2 is faiss arithmetic and only recorded)
    import builtins
    real_print = builtins.print
    builtins.print = lambda *a, **k: None
    try:
        for tag, paths, ordered in (("q1", ["set/img_0002.jpg"], ["color", "dreamsim", "sift"]),
                                    ("q2", ["set/img_0002.jpg", "set/img_0005.jpg"], ["color", "dreamsim", "sift"]),
                                    ("q3", ["set/img_0008.jpg"], ["color"])):
            # the reference resolves against images_root but create_db stores parent-relative paths;
            # query with the stored form so that the cache path (the in-scope one) is exercised
            _Recorder.normalized.clear()
            stored = [f"image_data/{p}" for p in paths]
            q = rec._extract_query_vector(stored, ordered)
            out[f"{tag}/combined"] = np.array(_Recorder.normalized[-1])
            out[f"{tag}/shape"] = np.array(q.shape)
    finally:
        builtins.print = real_print
    # offset -> path mapping with the offsets written by the combo build
    b = FAISSIndexBuilderDB(db_path="images.db", vector_types=["color", "sift", "dreamsim"], batch_size=8,
                            log_dir=str(work / "logs"))
    b.build_index()
    idx = np.array([[3, 0, 10, 7]])
    dst = np.array([[0.5, 0.25, 0.75, 0.125]], dtype=np.float32)
    res = rec._fetch_results(idx, dst, "faiss_index_offsets_color_sift_dreamsim")
    out["fetch/paths"] = np.array([str(p.relative_to(rec.base_dir)) for p, _ in res])
    out["fetch/dists"] = np.array([d for _, d in res], dtype=np.float64)
    np.savez_compressed(HERE / "ref_golden.npz", **out)
    shutil.rmtree(work, ignore_errors=True)
    print("wrote", db, "and", HERE / "ref_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
