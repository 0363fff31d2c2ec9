"""Native ingest (csrc/ingest.cu): the strict blob recogniser against the Python decoder on the CPU,
and — on the GPU — the whole SQL -> decode -> pinned staging -> K-pack loop against the Python build
loop that mirrors the reference (main/create_index.py:144-189, 301-313)."""
import pickle
import sqlite3
from pathlib import Path

import numpy as np
import pytest

DDL = """
CREATE TABLE IF NOT EXISTS images (id INTEGER PRIMARY KEY AUTOINCREMENT, path TEXT UNIQUE);
CREATE TABLE IF NOT EXISTS color_vectors (image_id INTEGER PRIMARY KEY, color_vector_blob BLOB,
    FOREIGN KEY(image_id) REFERENCES images(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS sift_vectors (image_id INTEGER PRIMARY KEY, sift_vector_blob BLOB,
    FOREIGN KEY(image_id) REFERENCES images(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS dreamsim_vectors (image_id INTEGER PRIMARY KEY, dreamsim_vector_blob BLOB,
    FOREIGN KEY(image_id) REFERENCES images(id) ON DELETE CASCADE);
"""   # the reference's schema (main/create_db.py:59-85)
DIMS = {"color": 48, "sift": 128, "dreamsim": 1792}


def _dump(v):
    # the reference's blob writer (vector_scripts/create_vector_base.py:142-145)
    return pickle.dumps(v, protocol=pickle.HIGHEST_PROTOCOL)


def _make_db(path, n, seed=0, foreign_at=None, missing=()):
    rng = np.random.default_rng(seed)
    conn = sqlite3.connect(path)
    conn.executescript(DDL)
    conn.executemany("INSERT INTO images (id, path) VALUES (?, ?)", [(i + 1, f"image_data/{i:06d}.jpg") for i in range(n)])
    for t, d in DIMS.items():
        rows = []
        for i in range(n):
            if (t, i + 1) in missing:
                continue
            v = rng.standard_normal(d).astype(np.float32)
            if t == "color":
                v = np.abs(v)
            if foreign_at == (t, i + 1):
                import torch
                v = torch.from_numpy(v)          # a tensor pickle: the reference decodes it, the recogniser must decline
            rows.append((i + 1, sqlite3.Binary(_dump(v))))
        conn.executemany(f"INSERT INTO {t}_vectors VALUES (?, ?)", rows)
    conn.commit()
    conn.close()


def test_native_blob_recogniser_agrees_with_python_fast_path():
    """b2k_parse_f32_blob accepts exactly the blobs FAISSIndexBuilderDB._decode_blob views in place,
    returns the same floats, and declines everything else (never misreads)."""
    import torch
    import image_recommender_b200 as irb
    from main.create_index import FAISSIndexBuilderDB as B
    rng = np.random.default_rng(11)
    for d in (1, 2, 48, 128, 255, 256, 257, 1792, 32768, 65535, 65536, 70000):
        v = rng.standard_normal(d).astype(np.float32)
        blob = _dump(v)
        got = irb.parse_f32_blob(blob)
        py = B._decode_blob(blob)
        assert got is not None and np.array_equal(got.view(np.uint32), v.view(np.uint32))
        assert np.shares_memory(py, np.frombuffer(blob, np.uint8))          # Python took its fast path too
    others = [rng.standard_normal((1, 64)).astype(np.float32), rng.standard_normal(64),
              rng.standard_normal(64).astype(np.float16), rng.standard_normal(64).astype(">f4"),
              np.asfortranarray(rng.standard_normal((4, 16)).astype(np.float32)),
              rng.standard_normal(128).astype(np.float32)[::2],
              torch.from_numpy(rng.standard_normal(64).astype(np.float32)), [0.5, 1.5, 2.5],
              np.zeros(0, np.float32)]
    for v in others:
        for proto in (pickle.HIGHEST_PROTOCOL, 4):
            assert irb.parse_f32_blob(pickle.dumps(v, protocol=proto)) is None
    good = _dump(rng.standard_normal(48).astype(np.float32))
    for bad in (b"", b"\x80\x05", b"\x80\x05not a pickle", good[:-1], good[:100], good[:-9] + b"X" + good[-8:],
                good.replace(b"f4", b"f8"), good.replace(b"\x8c\x01<", b"\x8c\x01>")):
        assert irb.parse_f32_blob(bad) is None
    # a lying length field cannot make the recogniser read outside the blob
    i = good.index(b"(\x96") + 2
    for n_bytes in (0, 3, 188, 196, 2 ** 40):
        assert irb.parse_f32_blob(good[:i] + int(n_bytes).to_bytes(8, "little") + good[i + 8:]) is None


def test_stage_api_rejects_use_without_device():
    """The staging entry points validate their arguments on a GPU-less host (no compute)."""
    import ctypes as C
    from image_recommender_b200 import _capi
    lib = _capi.load_library()
    p = C.c_void_p()
    assert lib.b2k_stage_open(None, 16) == _capi.E_INVALID
    assert lib.b2k_stage_ptr(None, 0, 0, C.byref(p)) == _capi.E_INVALID
    assert lib.b2k_stage_commit(None, 0, 1) == _capi.E_INVALID
    assert lib.b2k_stage_rows(None) == 0
    n = C.c_int64(0)
    assert lib.b2k_ingest_sqlite(None, b"x.db", b"select 1", None, 0, C.byref(n)) == _capi.E_INVALID


# ------------------------------------------------------------------------------------------ GPU
def _build(tmp, types_, native, **kw):
    from main.create_index import FAISSIndexBuilderDB
    out = tmp / f"index_{'native' if native else 'python'}.faiss"
    b = FAISSIndexBuilderDB(db_path=str(tmp / "images.db"), vector_types=types_, batch_size=500, index_file=str(out),
                            log_dir=str(tmp / "logs"), native_ingest=native, **kw)
    logged = []
    orig = b._log
    b._log = lambda m, level="info": (logged.append(m), orig(m, level))[1]
    b.build_index()
    conn = sqlite3.connect(tmp / "images.db")
    tab = conn.execute(f"SELECT image_id, offset FROM {b.offset_table} ORDER BY offset").fetchall()
    conn.close()
    return out, tab, logged


@pytest.mark.gpu
@pytest.mark.parametrize("types_", [["color", "sift", "dreamsim"], ["dreamsim"], ["sift", "color"]])
def test_native_ingest_equals_python_build(tmp_path, gpu, types_):
    """Index file and offset table of the native loop are byte-identical to the Python loop's
    (inner join drops images with a missing part; ids ascending; more rows than one staging slot)."""
    _make_db(tmp_path / "images.db", 2500, seed=1, missing={("sift", 7), ("sift", 1200), ("color", 2500), ("dreamsim", 1)})
    f_nat, tab_nat, log_nat = _build(tmp_path, types_, True)
    f_py, tab_py, log_py = _build(tmp_path, types_, False)
    assert any("Native ingest: added" in m for m in log_nat) and not any("Native ingest" in m for m in log_py)
    assert tab_nat == tab_py and len(tab_nat) >= 2496
    assert f_nat.read_bytes() == f_py.read_bytes()


@pytest.mark.gpu
def test_native_ingest_small_slots_and_staging_reuse(tmp_path, gpu):
    """Many tiny staging slots (rows_per_slot = 64) exercise the two-slot ping-pong."""
    import image_recommender_b200 as irb
    _make_db(tmp_path / "images.db", 1000, seed=2)
    sql = ("SELECT i.id, c.color_vector_blob, s.sift_vector_blob, d.dreamsim_vector_blob FROM images i "
           "JOIN color_vectors c ON i.id = c.image_id JOIN sift_vectors s ON i.id = s.image_id "
           "JOIN dreamsim_vectors d ON i.id = d.image_id")
    a = irb.FlatShard([48, 128, 1792], 1000, device=gpu)
    ids = a.ingest_sqlite(tmp_path / "images.db", sql, 1000, rows_per_slot=64)
    b = irb.FlatShard([48, 128, 1792], 1000, device=gpu)
    ids_b = b.ingest_sqlite(tmp_path / "images.db", sql, 1000)
    assert ids.tolist() == list(range(1, 1001)) == ids_b.tolist() and a.ntotal == b.ntotal == 1000
    for x, y in zip(a.get_rows(0, 1000), b.get_rows(0, 1000)):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    with pytest.raises(irb.B2KError) as e:          # capacity of ids_out is enforced
        a.reset(); a.ingest_sqlite(tmp_path / "images.db", sql, 10)
    assert e.value.status == irb._capi.E_CAPACITY
    a.close(); b.close()


@pytest.mark.gpu
def test_native_ingest_reader_threads_equal_single_thread(tmp_path, gpu):
    """b2k_ingest_sqlite_mt: 1, 3 and 8 reader threads over id chunks of 64 images (many more chunks than
    threads, images with a missing part inside and at the edges of chunks) append exactly the rows, in exactly
    the order, of the single-threaded loop; a foreign blob stops every thread with E_UNSUPPORTED; the capacity of
    ids_out is enforced."""
    import image_recommender_b200 as irb
    from main.create_index import FAISSIndexBuilderDB
    _make_db(tmp_path / "images.db", 1500, seed=4, missing={("sift", 1), ("color", 64), ("dreamsim", 65), ("sift", 1500)})
    b = FAISSIndexBuilderDB(db_path=str(tmp_path / "images.db"), vector_types=["color", "sift", "dreamsim"],
                            log_dir=str(tmp_path / "logs"))
    sel, joins = b._make_select_and_joins()
    sql = f"SELECT {sel} FROM images i {joins} ORDER BY i.id"
    sql_range = f"SELECT {sel} FROM images i {joins} WHERE i.id >= ?1 AND i.id < ?2 ORDER BY i.id"
    bounds = b._id_chunk_bounds(64)
    assert bounds[0] == 1 and bounds[-1] == 1501 and len(bounds) == 25
    ref = irb.FlatShard([48, 128, 1792], 1500, device=gpu)
    ids_ref = ref.ingest_sqlite(tmp_path / "images.db", sql, 1500)
    assert len(ids_ref) == 1496
    for threads in (1, 3, 8):
        a = irb.FlatShard([48, 128, 1792], 1500, device=gpu)
        ids = a.ingest_sqlite_mt(tmp_path / "images.db", sql_range, bounds, 1500, threads, rows_per_slot=64)
        assert ids.tolist() == ids_ref.tolist() and a.ntotal == ref.ntotal
        for x, y in zip(a.get_rows(0, a.ntotal), ref.get_rows(0, ref.ntotal)):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
        a.reset()
        with pytest.raises(irb.B2KError) as e:
            a.ingest_sqlite_mt(tmp_path / "images.db", sql_range, bounds, 100, threads, rows_per_slot=64)
        assert e.value.status == irb._capi.E_CAPACITY
        a.close()
    ref.close()
    # the builder: an explicit thread count is taken as given, the default needs 50 000 rows per thread
    assert b._ingest_threads(1500) == 1 and 1 <= b._ingest_threads(10_000_000) <= 8
    b.ingest_threads = 4
    assert b._ingest_threads(1500) == 4
    b.read_conn.close(); b.write_conn.close()
    f_mt, tab_mt, log_mt = _build(tmp_path, ["color", "sift", "dreamsim"], True, ingest_threads=3)
    f_py, tab_py, _ = _build(tmp_path, ["color", "sift", "dreamsim"], False)
    assert any("Native ingest: added" in m for m in log_mt) and tab_mt == tab_py and f_mt.read_bytes() == f_py.read_bytes()
    _make_db(tmp_path / "foreign.db", 900, seed=5, foreign_at=("dreamsim", 700))
    a = irb.FlatShard([48, 128, 1792], 900, device=gpu)
    fb = FAISSIndexBuilderDB(db_path=str(tmp_path / "foreign.db"), vector_types=["color", "sift", "dreamsim"],
                             log_dir=str(tmp_path / "logs"))
    with pytest.raises(irb.B2KError) as e:
        a.ingest_sqlite_mt(tmp_path / "foreign.db", sql_range, fb._id_chunk_bounds(64), 900, 4, rows_per_slot=64)
    assert e.value.status == irb._capi.E_UNSUPPORTED and "image id 700" in str(e.value)
    fb.read_conn.close(); fb.write_conn.close()
    a.close()


@pytest.mark.gpu
def test_native_ingest_declines_foreign_blob_and_python_takes_over(tmp_path, gpu):
    """One tensor pickle in the middle of the table: the native loop stops with E_UNSUPPORTED, the
    builder rebuilds in Python (which decodes it as the reference does): same file either way."""
    _make_db(tmp_path / "images.db", 700, seed=3, foreign_at=("sift", 400))
    f_nat, tab_nat, log_nat = _build(tmp_path, ["color", "sift", "dreamsim"], True)
    f_py, tab_py, _ = _build(tmp_path, ["color", "sift", "dreamsim"], False)
    assert any("Native ingest declined" in m for m in log_nat)
    assert len(tab_nat) == 700 and tab_nat == tab_py
    assert f_nat.read_bytes() == f_py.read_bytes()


# ------------------------------------------------------------------------------- property tests (CPU)
def test_blob_recogniser_never_misreads_mutated_blobs():
    """Fuzz: random byte edits / truncations of valid blobs.  Whatever the recogniser accepts must be
    exactly what the unpickler (the reference's decoder) yields for the same bytes — it may decline
    anything, it may never invent floats — and it must not read outside the buffer (no crash)."""
    from hypothesis import given, settings, strategies as st
    import image_recommender_b200 as irb

    base = [_dump(np.random.default_rng(i).standard_normal(d).astype(np.float32)) for i, d in enumerate((1, 7, 48, 128, 300))]

    @settings(max_examples=400, deadline=None)
    @given(which=st.integers(0, len(base) - 1), edits=st.lists(st.tuples(st.integers(0, 1 << 30), st.integers(0, 255)), max_size=4),
           cut=st.one_of(st.none(), st.integers(0, 1 << 30)))
    def run(which, edits, cut):
        b = bytearray(base[which])
        for pos, val in edits:
            b[pos % len(b)] = val
        if cut is not None:
            b = b[:cut % (len(b) + 1)]
        blob = bytes(b)
        got = irb.parse_f32_blob(blob)
        if got is None:
            return
        try:
            want = pickle.loads(blob)
        except Exception:
            want = None
        if isinstance(want, np.ndarray) and want.dtype == np.float32 and want.ndim == 1:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        else:
            # accepted although the unpickler disagrees: only tolerable when the edit hit pickle
            # framing/memo bytes the recogniser does not interpret, never the payload length or dtype
            head = blob.index(b"(\x96")
            n_bytes = int.from_bytes(blob[head + 2:head + 10], "little")
            assert got.size * 4 == n_bytes and blob[head + 10:head + 10 + n_bytes] == got.tobytes()

    run()


def test_shard_ranges_partition_the_rows():
    """shard_range(): contiguous, ordered, disjoint ranges covering [0, n) exactly — the property that
    makes global offset = base + local row (create_index.py:236-249) hold on any number of GPUs."""
    from hypothesis import given, settings, strategies as st
    from image_recommender_b200.sharded import shard_range

    @settings(max_examples=300, deadline=None)
    @given(n=st.integers(0, 10 ** 9), world=st.integers(1, 64))
    def run(n, world):
        edges = [shard_range(n, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
            assert a0 <= a1 == b0 <= b1
        sizes = [b - a for a, b in edges]
        assert sizes == sorted(sizes, reverse=True)               # ceil(n/world) rows each, only tail shards shorter
        assert max(sizes) == -(-n // world)

    run()


def test_id_chunk_bounds_cover_sparse_ids(tmp_path):
    """The reader threads' chunking (FAISSIndexBuilderDB._id_chunk_bounds): bounds are strictly increasing, every
    chunk [b[c], b[c+1]) holds at most rows_per_chunk images, and together they hold every image exactly once — with
    gaps in the ids (deleted images, AUTOINCREMENT jumps)."""
    from main.create_index import FAISSIndexBuilderDB
    rng = np.random.default_rng(3)
    ids = np.unique(rng.integers(1, 10**7, size=5000)).tolist()
    conn = sqlite3.connect(tmp_path / "images.db")
    conn.executescript(DDL)
    conn.executemany("INSERT INTO images (id, path) VALUES (?, ?)", [(i, f"p/{i}.jpg") for i in ids])
    conn.commit()
    conn.close()
    b = FAISSIndexBuilderDB(db_path=str(tmp_path / "images.db"), vector_types=["color"], log_dir=str(tmp_path / "logs"))
    for per in (1, 7, 64, 4096, 10**6):
        bounds = b._id_chunk_bounds(per)
        assert bounds[0] == ids[0] and bounds[-1] == ids[-1] + 1
        assert all(x < y for x, y in zip(bounds, bounds[1:]))
        arr = np.asarray(ids)
        counts = [int(((arr >= lo) & (arr < hi)).sum()) for lo, hi in zip(bounds, bounds[1:])]
        assert sum(counts) == len(ids) and max(counts) <= per and min(counts) >= 1
    assert b._ingest_threads(10) == 1 and b._ingest_threads(10**7, share=8) >= 1
    b.read_conn.close(); b.write_conn.close()
    empty = sqlite3.connect(tmp_path / "empty.db")
    empty.executescript(DDL)
    empty.commit(); empty.close()
    e = FAISSIndexBuilderDB(db_path=str(tmp_path / "empty.db"), vector_types=["color"], log_dir=str(tmp_path / "logs"))
    assert e._id_chunk_bounds(64) == [0, 0]
    e.read_conn.close(); e.write_conn.close()
