"""Host-side mirror (main/create_index.py, main/search_from_image.py) and the oracle against the
golden vectors produced by the REAL reference code (tests/golden/make_golden.py)."""
import pickle
import shutil
import sqlite3
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle

GOLD = Path(__file__).resolve().parent / "golden"
G = np.load(GOLD / "ref_golden.npz")
COMBOS = [["color"], ["color", "sift", "dreamsim"], ["dreamsim", "color"]]


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    shutil.copy(GOLD / "ref_fixture.db", tmp_path / "images.db")
    monkeypatch.chdir(tmp_path)
    return tmp_path


def _builder(types_, **kw):
    from main.create_index import FAISSIndexBuilderDB
    return FAISSIndexBuilderDB(db_path="images.db", vector_types=types_, batch_size=8, log_dir="logs", **kw)


@pytest.mark.parametrize("types_", COMBOS)
def test_decode_concat_order_match_reference(workdir, types_):
    """SQL join order, skip-on-decode-error, tensor/2-D blobs, concat: bit-equal to what the
    reference handed to index.add (create_index.py:115-189, 304-311)."""
    name = "_".join(types_)
    b = _builder(types_)
    assert b._count_records() == int(G[f"{name}/count"])
    assert str(b.index_file) == str(G[f"{name}/index_file"])
    assert b.offset_table == f"faiss_index_offsets_{name}"
    ids, rows, sizes, off = [], [], [], 0
    for batch in b._batch_records():
        i, e = b._process_batch(batch)
        if not e:
            continue
        b._store_offsets(i, off)
        off += len(i)
        ids += i
        rows.append(np.stack(e).astype("float32"))
        sizes.append(len(i))
    got = np.concatenate(rows)
    want = G[f"{name}/added"]
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert sizes == G[f"{name}/batch_sizes"].tolist()
    conn = sqlite3.connect("images.db")
    tab = conn.execute(f"SELECT image_id, offset FROM {b.offset_table} ORDER BY offset").fetchall()
    conn.close()
    assert np.array_equal(np.array(tab, dtype=np.int64), G[f"{name}/offsets"])
    assert ids == G[f"{name}/offsets"][:, 0].tolist()


def test_oracle_concat_matches_reference_rows(workdir):
    """oracle.pack(normalize=False) is the reference's _process_batch concat, on its own rows; and
    since the fixture parts are unit-norm, Spec P's normalisation moves them by a few ulp only."""
    b = _builder(["color", "sift", "dreamsim"])
    parts = []
    for batch in b._batch_records():
        _, p = b._decode_batch(batch)
        parts += p
    tables = [np.stack([p[t] for p in parts]) for t in range(3)]
    want = G["color_sift_dreamsim/added"]
    assert np.array_equal(oracle.pack(tables, normalize=False)["f32"].view(np.uint32), want.view(np.uint32))
    normed = oracle.pack(tables, normalize=True)["f32"]
    ulp = np.abs(normed.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
    assert ulp.max() <= 4
    assert np.allclose(np.linalg.norm(want.astype(np.float64), axis=1), np.sqrt(3.0), atol=1e-6)


def test_ordered_index_types_match_reference(workdir):
    from main.search_from_image import ImageRecommender
    rec = ImageRecommender(images_root="image_data", db_path="images.db")
    for case, want in zip(G["ordered/cases"].tolist(), G["ordered/results"].tolist()):
        assert "|".join(rec._get_ordered_index_types(case)) == want
    # README form (not parsed by the reference)
    assert rec._get_ordered_index_types("combo_color_sift_dreamsim") == ["color", "dreamsim", "sift"]


def test_query_vector_concat_mean_match_reference(workdir, monkeypatch):
    """concat (axis=1) -> mean over images: bit-equal to what the reference passed to
    faiss.normalize_L2 (search_from_image.py:305-322)."""
    import main.search_from_image as sfi
    seen = []
    monkeypatch.setattr(sfi, "normalize_L2", lambda x, device=0: seen.append(np.array(x)))
    rec = sfi.ImageRecommender(images_root="image_data", db_path="images.db")
    cases = (("q1", ["image_data/set/img_0002.jpg"], ["color", "dreamsim", "sift"]),
             ("q2", ["image_data/set/img_0002.jpg", "image_data/set/img_0005.jpg"], ["color", "dreamsim", "sift"]),
             ("q3", ["image_data/set/img_0008.jpg"], ["color"]))
    for tag, paths, ordered in cases:
        q = rec._extract_query_vector(paths, ordered)
        assert list(q.shape) == G[f"{tag}/shape"].tolist()
        assert np.array_equal(seen[-1].view(np.uint32), G[f"{tag}/combined"].view(np.uint32))
    # the images_root-relative form the reference computes (search_from_image.py:230-232) resolves too
    q = rec._extract_query_vector(["set/img_0008.jpg"], ["color"])
    assert np.array_equal(seen[-1].view(np.uint32), G["q3/combined"].view(np.uint32))
    # raw little-endian float32 blob fallback (search_from_image.py:82-91)
    conn = sqlite3.connect("images.db")
    v = np.arange(48, dtype=np.float32)
    conn.execute("UPDATE color_vectors SET color_vector_blob = ? WHERE image_id = 1", (sqlite3.Binary(v.tobytes()),))
    conn.commit(); conn.close()
    got = rec._get_db_vector("image_data/set/img_0000.jpg", "color_vectors", "color_vector_blob")
    assert np.array_equal(got, v.reshape(1, -1))
    assert rec._extract_query_vector(["image_data/set/missing.jpg"], ["color"]) is None


def test_fetch_results_match_reference(workdir):
    from main.search_from_image import ImageRecommender
    b = _builder(["color", "sift", "dreamsim"])
    off = 0
    for batch in b._batch_records():
        i, _ = b._process_batch(batch)
        b._store_offsets(i, off)
        off += len(i)
    rec = ImageRecommender(images_root="image_data", db_path="images.db")
    res = rec._fetch_results(np.array([[3, 0, 10, 7]]), np.array([[0.5, 0.25, 0.75, 0.125]], dtype=np.float32),
                             "faiss_index_offsets_color_sift_dreamsim")
    assert [str(p.relative_to(rec.base_dir)) for p, _ in res] == G["fetch/paths"].tolist()
    assert [d for _, d in res] == G["fetch/dists"].tolist()
    # -1 padding (k > ntotal) is dropped, not looked up
    res = rec._fetch_results(np.array([[3, -1]]), np.array([[0.5, 3.4e38]], dtype=np.float32),
                             "faiss_index_offsets_color_sift_dreamsim")
    assert len(res) == 1


def test_batched_query_acquisition_equals_per_image_path(workdir, monkeypatch):
    """search_batch's IN-select acquisition builds, per group, the same bits as the
    reference-shaped per-image path (_extract_query_vector), and maps hits like _fetch_results."""
    import main.search_from_image as sfi
    monkeypatch.setattr(sfi, "normalize_L2", lambda x, device=0: None)      # compare what is handed to it
    rec = sfi.ImageRecommender(images_root="image_data", db_path="images.db")
    conn = sqlite3.connect("images.db")
    paths = [p for (p,) in conn.execute("SELECT path FROM images ORDER BY id")]
    conn.close()
    ordered = ["color", "dreamsim", "sift"]
    i2, i5 = "image_data/set/img_0002.jpg", "image_data/set/img_0005.jpg"      # the golden q1 / q2 images
    groups = [[i2], [i2, i5], ["set/" + Path(paths[8]).name], [paths[0], "image_data/set/missing.jpg"],
              ["image_data/set/missing.jpg"], [paths[3], paths[3], paths[9]]]
    q, live = rec._extract_query_matrix(groups, ordered)
    want = [rec._extract_query_vector(g, ordered) for g in groups]
    assert live == [i for i, w in enumerate(want) if w is not None] and 4 not in live
    for row, gi in enumerate(live):
        assert np.array_equal(q[row:row + 1].view(np.uint32), want[gi].view(np.uint32))
    assert np.array_equal(q[0:1].view(np.uint32), G["q1/combined"].view(np.uint32))
    assert np.array_equal(q[1:2].view(np.uint32), G["q2/combined"].view(np.uint32))
    # hits -> paths: through the offset table, and through ids kept with the index
    b = _builder(["color", "sift", "dreamsim"])
    off, all_ids = 0, []
    for batch in b._batch_records():
        i, _ = b._process_batch(batch)
        b._store_offsets(i, off)
        off += len(i)
        all_ids += i
    idx = np.array([[3, 0, 10, 7], [5, -1, 2, 2], [23, 22, 21, 20]])
    dist = np.array([[0.5, 0.25, 0.75, 0.125], [0.1, 3.4e38, 0.3, 0.2], [4, 3, 2, 1]], dtype=np.float32)
    table = "faiss_index_offsets_color_sift_dreamsim"
    single = [rec._fetch_results(idx[i:i + 1], dist[i:i + 1], table) for i in range(3)]
    assert rec._fetch_results_batch(idx, dist, table) == single
    token = object()
    rec._resident_ids[id(token)] = np.asarray(all_ids, dtype=np.int64)
    assert rec._fetch_results_batch(idx, dist, table, token) == single
    assert [str(p.relative_to(rec.base_dir)) for p, _ in single[0]] == G["fetch/paths"].tolist()


def test_cli_parsers():
    import main.create_index as ci
    import main.search_from_image as sfi
    with pytest.raises(SystemExit):
        sfi.main(["--db-path", "x.db"])          # --query is required
    with pytest.raises(SystemExit):
        ci.main(["--no-such-flag"])


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_build_and_search_end_to_end(workdir, gpu):
    """README flow: build `color sift dreamsim`, query `combo_color_sift_dreamsim`; offsets equal
    the reference's; stored rows = Spec P of the reference's rows; every image finds itself."""
    from main import create_index as ci
    from main.search_from_image import ImageRecommender
    import image_recommender_b200 as irb
    ci.main(["--db-path", "images.db", "--vector-types", "color", "sift", "dreamsim", "--batch-size", "8"])
    f = Path("index_hnsw_color_sift_dreamsim.faiss")
    assert f.exists()
    conn = sqlite3.connect("images.db")
    tab = conn.execute("SELECT image_id, offset FROM faiss_index_offsets_color_sift_dreamsim ORDER BY offset").fetchall()
    paths = dict(conn.execute("SELECT id, path FROM images").fetchall())
    conn.close()
    assert np.array_equal(np.array(tab, dtype=np.int64), G["color_sift_dreamsim/offsets"])
    info = irb.file_info(f)
    assert info["n_rows"] == 24 and info["table_dims"] == [48, 128, 1792] and info["has_ids"]
    assert irb.load_ids(f, 0, 24).tolist() == [t[0] for t in tab]
    want_rows = G["color_sift_dreamsim/added"]
    pk = oracle.pack([want_rows[:, :48], want_rows[:, 48:176], want_rows[:, 176:]])
    ix = irb.FlatShard.load(f, device=gpu)
    got, _, _ = ix.get_rows(0, 24)
    assert np.array_equal(got.view(np.uint32), pk["f32"].view(np.uint32))
    ix.close()

    rec = ImageRecommender(images_root="image_data", db_path="images.db", top_k=5)
    for image_id, offset in tab[:6]:
        q_path = str(Path("image_data").resolve().parent / paths[image_id])
        res = rec.search_similar_images([q_path], index_type="combo_color_sift_dreamsim")
        assert res is not None and len(res) == 5
        assert res[0][0] == rec.base_dir / paths[image_id]
        assert abs(res[0][1] - (np.sqrt(3.0) - 1) ** 2) < 1e-5          # SURVEY §8c self-query property
        assert [d for _, d in res] == sorted(d for _, d in res)
        # same answer as the oracle on the same query vector
        parts = [want_rows[offset:offset + 1, :48], want_rows[offset:offset + 1, 48:176],
                 want_rows[offset:offset + 1, 176:]]
        order = [0, 1, 2]                                                # file order = color, sift, dreamsim
        qv = oracle.normalize_l2(np.concatenate([parts[i] for i in order], axis=1))
        d, lab, _ = oracle.search_exact(pk["f32"], qv, 5, pk["norm2"])
        assert [float(x) for x in d[0]] == [r[1] for r in res]
    # averaged multi-image query + batch API
    two = [str(Path("image_data").resolve().parent / paths[tab[0][0]]),
           str(Path("image_data").resolve().parent / paths[tab[1][0]])]
    assert len(rec.search_similar_images(two, index_type="color,sift,dreamsim")) == 5
    groups = rec.search_batch([[two[0]], [two[1]], two], index_type="color,sift,dreamsim")
    assert [len(g) for g in groups] == [5, 5, 5] and groups[0][0][0] == rec.base_dir / paths[tab[0][0]]
    # resident CLI mode: one process, index loaded once, one query per stdin line
    import io
    import contextlib
    from main import search_from_image as sfi
    out = io.StringIO()
    stdin = sys.stdin
    sys.stdin = io.StringIO(f"{two[0]}\n\n{two[0]} {two[1]}\nimage_data/set/missing.jpg\n{two[1]}\n")
    try:
        with contextlib.redirect_stdout(out):
            rc = sfi.main(["--db-path", "images.db", "--images-root", "image_data", "--index", "combo_color_sift_dreamsim",
                           "--top-k", "3", "--serve"])
    finally:
        sys.stdin = stdin
    blocks = [b.strip().splitlines() for b in out.getvalue().split("\n\n")]
    answers = [b for b in blocks if b and "\t" in b[0]]
    assert rc == 0 and len(answers) == 3 and all(len(b) == 3 for b in answers)
    assert answers[0][0].endswith(paths[tab[0][0]]) and answers[2][0].endswith(paths[tab[1][0]])
    # update_index appends only images without an offset
    conn = sqlite3.connect("images.db")
    v = lambda d: sqlite3.Binary(pickle.dumps((np.ones(d, np.float32) / np.sqrt(d)), protocol=pickle.HIGHEST_PROTOCOL))  # noqa: E731
    conn.execute("INSERT INTO sift_vectors VALUES (7, ?)", (v(128),))
    conn.commit(); conn.close()
    rec.close()
    ci.main(["--db-path", "images.db", "--vector-types", "color", "sift", "dreamsim", "--update"])
    assert irb.file_info(f)["n_rows"] == 25
    conn = sqlite3.connect("images.db")
    assert conn.execute("SELECT offset FROM faiss_index_offsets_color_sift_dreamsim WHERE image_id = 7").fetchone()[0] == 24
    conn.close()


def test_blob_fast_path_equals_unpickler():
    """The in-place view of numpy float32 pickles must agree with pickle.loads on every layout the
    tables can hold, and must fall back (not misread) on everything else."""
    import torch
    from main.create_index import FAISSIndexBuilderDB as B
    rng = np.random.default_rng(3)
    dump = lambda v: pickle.dumps(v, protocol=pickle.HIGHEST_PROTOCOL)  # noqa: E731
    for d in (1, 48, 128, 255, 256, 1792, 32768, 70000):
        v = rng.standard_normal(d).astype(np.float32)
        blob = dump(v)
        got = B._decode_blob(blob)
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), v.view(np.uint32))
        assert np.shares_memory(got, np.frombuffer(blob, np.uint8)) or d == 0     # fast path taken
    slow = [rng.standard_normal((1, 64)).astype(np.float32),                       # 2-D
            rng.standard_normal(64),                                                # float64
            rng.standard_normal(64).astype(np.float16),
            rng.standard_normal(64).astype(">f4"),                                  # big endian
            np.asfortranarray(rng.standard_normal((4, 16)).astype(np.float32)),
            rng.standard_normal(128).astype(np.float32)[::2],                       # strided
            torch.from_numpy(rng.standard_normal(64).astype(np.float32)),
            [0.5, 1.5, 2.5]]
    for v in slow:
        want = np.asarray(v.cpu().numpy() if hasattr(v, "cpu") else v, dtype="float32").ravel()
        assert np.array_equal(B._decode_blob(dump(v)), want)
        assert np.array_equal(B._decode_blob(pickle.dumps(v, protocol=4)), want)
    with pytest.raises(Exception):
        B._decode_blob(b"\x80\x05not a pickle")
