"""GPU box: build-half throughput, native ingest (csrc/ingest.cu) vs the Python decode loop, on a
synthetic SQLite DB in the reference's schema and blob format.  Usage: python scripts/bench_ingest.py [rows]"""
import json, sys, time, tempfile, shutil
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np
from test_ingest import _make_db
from main.create_index import FAISSIndexBuilderDB

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
tmp = Path(tempfile.mkdtemp(prefix="b2k_ingest_"))
t0 = time.perf_counter(); _make_db(tmp / "images.db", n, seed=5); t_db = time.perf_counter() - t0
res = {"rows": n, "db_mb": round((tmp / "images.db").stat().st_size / 1e6, 1), "make_db_s": round(t_db, 2)}
files = {}
for name, native in (("warmup", True), ("native", True), ("python", False)):
    out = tmp / f"{name}.faiss"
    b = FAISSIndexBuilderDB(db_path=str(tmp / "images.db"), vector_types=["color", "sift", "dreamsim"], index_file=str(out),
                            log_dir=str(tmp / "logs"), native_ingest=native)
    b._log = lambda m, level="info": None
    t0 = time.perf_counter(); b.build_index(); dt = time.perf_counter() - t0
    res[name] = {"s": round(dt, 3), "rows_per_s": round(n / dt)}
    files[name] = out
# the ingest loops alone (no COUNT(*), offset table, index file)
import image_recommender_b200 as irb
b = FAISSIndexBuilderDB(db_path=str(tmp / "images.db"), vector_types=["color", "sift", "dreamsim"], log_dir=str(tmp / "logs"))
b._log = lambda m, level="info": None
sel, joins = b._make_select_and_joins()
sql = f"SELECT {sel} FROM images i {joins}"
for rep in range(2):
    ix = irb.FlatShard([48, 128, 1792], n, device=0)
    from image_recommender_b200.index import _lib, check
    import ctypes as C
    t0 = time.perf_counter(); check(_lib.b2k_stage_open(ix._h, 4096)); t_open = time.perf_counter() - t0
    ids = np.empty(n, np.int64); cnt = C.c_int64(0)
    t0 = time.perf_counter()
    check(_lib.b2k_ingest_sqlite(ix._h, str(tmp / "images.db").encode(), sql.encode(), ids.ctypes.data, n, C.byref(cnt)))
    dt = time.perf_counter() - t0
    t0 = time.perf_counter(); _lib.b2k_stage_close(ix._h); t_close = time.perf_counter() - t0
    assert cnt.value == n
    res["native_loop"] = {"s": round(dt, 3), "rows_per_s": round(n / dt), "gb_per_s": round(n * 7872 / dt / 1e9, 2),
                          "stage_open_s": round(t_open, 3), "stage_close_s": round(t_close, 3)}
    ix.close()
    # reader threads (b2k_ingest_sqlite_mt): own connection + two staging slots each, chunks committed in id order
    sql_range = f"SELECT {sel} FROM images i {joins} WHERE i.id >= ?1 AND i.id < ?2 ORDER BY i.id"
    bounds = b._id_chunk_bounds(2048)
    res["native_threads"] = {}
    for threads in (1, 2, 4, 8, 16):
        ix = irb.FlatShard([48, 128, 1792], n, device=0)
        bnd = np.ascontiguousarray(bounds, dtype=np.int64)
        got = np.empty(n, np.int64); cnt = C.c_int64(0)
        t0 = time.perf_counter(); check(_lib.b2k_stage_open_n(ix._h, 2048, 2 * threads)); t_open = time.perf_counter() - t0
        t0 = time.perf_counter()
        check(_lib.b2k_ingest_sqlite_mt(ix._h, str(tmp / "images.db").encode(), sql_range.encode(), bnd.ctypes.data, bnd.size - 1,
                                        threads, got.ctypes.data, n, C.byref(cnt)))
        dt = time.perf_counter() - t0
        t0 = time.perf_counter(); _lib.b2k_stage_close(ix._h); t_close = time.perf_counter() - t0
        assert cnt.value == n and (got == ids[:n]).all()
        res["native_threads"][str(threads)] = {"s": round(dt, 3), "rows_per_s": round(n / dt), "gb_per_s": round(n * 7872 / dt / 1e9, 2),
                                               "stage_open_s": round(t_open, 3), "stage_close_s": round(t_close, 3)}
        ix.close()
    ix = irb.FlatShard([48, 128, 1792], n, device=0)
    t0 = time.perf_counter()
    for batch in b._batch_records():
        ids_b, parts = b._decode_batch(batch)
        ix.add_tables([np.stack([p[t] for p in parts]).astype("float32") for t in range(3)])
    dt = time.perf_counter() - t0
    res["python_loop"] = {"s": round(dt, 3), "rows_per_s": round(n / dt)}
    ix.close()
res["identical_files"] = files["native"].read_bytes() == files["python"].read_bytes()
print(json.dumps(res))
shutil.rmtree(tmp, ignore_errors=True)
