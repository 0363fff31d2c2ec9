"""Index build for the retrieval hot path — drop-in for the reference's main/create_index.py.

Same class, constructor kwargs, method names, SQLite schema and offset table as the reference
(`FAISSIndexBuilderDB`, /root/reference/main/create_index.py:13-325); the FAISS HNSW / IVFPQ
index is replaced by an exact flat store on the GPU (`image_recommender_b200.FlatShard`):
rows are per-table L2-normalised, concatenated and bf16-packed by a CUDA kernel as they are
added.  There is no CPU path: without a B200 the build raises.

CLI (README.md:101-105 of the reference; the reference itself hard-codes its parameters):

    python -m main.create_index --db-path images.db --vector-types color sift dreamsim \
        --output index_hnsw.faiss [--batch-size 8192] [--hnsw_M 32] [--efConstruction 200] [--efSearch 64]
"""
from __future__ import annotations

import argparse
import logging
import os
import pickle
import sqlite3
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from image_recommender_b200 import FlatShard  # noqa: E402  (raises if the CUDA extension is missing)


class FAISSIndexBuilderDB:
    def __init__(
        self,
        db_path: str = "images.db",
        vector_types: list = None,
        batch_size: int = 8192,
        index_file: str = None,
        hnsw_M: int = 32,
        efConstruction: int = 200,
        efSearch: int = 64,
        log_file: str = "faiss_builder.log",
        log_dir: str = "logs",
        device: int = 0,
        native_ingest: bool = True,
        ingest_threads: int | None = None,
        sharded: bool = None,
    ):
        # reference: create_index.py:14-53 (same attribute names; `device` is new)
        self.log_dir = log_dir
        self.log_file = log_file
        self._setup_logging()

        self.db_path = db_path
        if not vector_types:
            raise TypeError("vector_types must be a non-empty list, e.g. ['color', 'sift', 'dreamsim']")
        self.vector_types = list(vector_types)
        self.vector_cols = [f"{t}_vector_blob" for t in self.vector_types]
        self.batch_size = batch_size

        name = "_".join(self.vector_types)
        self.index_file = Path(index_file) if index_file else Path(f"index_hnsw_{name}.faiss")

        # accepted for API compatibility; exact search has no graph to tune
        self.hnsw_M = hnsw_M
        self.efConstruction = efConstruction
        self.efSearch = efSearch
        self.device = device
        self.native_ingest = native_ingest     # new: SQL -> decode -> pack loop in libb2k.so (csrc/ingest.cu)
        self.ingest_threads = ingest_threads   # new: reader threads of the native ingest (None: half the cores, <= 8)
        # new: under torchrun (torch.distributed initialised, world > 1) every rank ingests and packs ITS row
        # range on its own GPU and all ranks write one index file together (None = auto, False = never)
        self.sharded = sharded

        self.offset_table = f"faiss_index_offsets_{name}"

        self.read_conn = sqlite3.connect(self.db_path, timeout=60)
        self._configure_db(self.read_conn)
        self.read_cur = self.read_conn.cursor()

        self.write_conn = sqlite3.connect(self.db_path, timeout=60)
        self._configure_db(self.write_conn)
        self.write_cur = self.write_conn.cursor()

        self._prepare_offset_table()

    # ---- logging (create_index.py:55-87) ---------------------------------------------------
    def _setup_logging(self):
        Path(self.log_dir).mkdir(parents=True, exist_ok=True)
        full_path = Path(self.log_dir) / self.log_file
        logging.basicConfig(
            level=logging.INFO,
            filename=str(full_path),
            filemode="a",
            format="%(asctime)s - %(levelname)s - %(message)s",
            encoding="utf-8",
        )
        self._log(f"Logging initialized (file={full_path})", level="info")

    def _log(self, message: str, level: str = "info"):
        print(message)
        getattr(logging, {"info": "info", "warning": "warning", "error": "error"}.get(level.lower(), "debug"))(message)

    # ---- SQLite (create_index.py:89-158) ------------------------------------------------------
    def _configure_db(self, conn):
        # several ranks of a sharded build open the database at once: switching the journal mode needs a
        # moment of exclusive access and does not wait on the busy handler, so retry it
        import time
        for attempt in range(200):
            try:
                conn.execute("PRAGMA journal_mode=WAL;")
                break
            except sqlite3.OperationalError as e:
                if "locked" not in str(e) or attempt == 199:
                    raise
                time.sleep(0.05)
        conn.execute("PRAGMA synchronous=OFF;")

    def _prepare_offset_table(self):
        self.write_cur.execute(
            f"""
            CREATE TABLE IF NOT EXISTS {self.offset_table} (
                image_id INTEGER PRIMARY KEY,
                offset   INTEGER
            );
            """
        )
        # new: the reference looks hits up with an unindexed `WHERE offset = ?` (F7)
        self.write_cur.execute(
            f"CREATE INDEX IF NOT EXISTS idx_{self.offset_table}_offset ON {self.offset_table} (offset);"
        )
        self.write_conn.commit()
        self._log(f"Offset table '{self.offset_table}' is ready.", level="info")

    def _make_select_and_joins(self):
        select_cols = ["i.id"]
        join_strs = []
        for vtype in self.vector_types:
            alias = vtype[0]
            vtable = f"{vtype}_vectors"
            vcol = f"{vtype}_vector_blob"
            select_cols.append(f"{alias}.{vcol}")
            join_strs.append(f"JOIN {vtable} {alias} ON i.id = {alias}.image_id")
        return ", ".join(select_cols), " ".join(join_strs)

    def _count_records(self):
        _, join_strs = self._make_select_and_joins()
        query = f"SELECT COUNT(*) FROM images i {join_strs}"
        return self.read_cur.execute(query).fetchone()[0]

    def _batch_records(self):
        select_cols, join_strs = self._make_select_and_joins()
        # the reference's join without an ORDER BY returns ascending images.id only because SQLite happens to
        # drive it from `images`; offsets are defined by that order, so it is made explicit here
        # (and is the order of the row-sharded build, whose files must equal a single-GPU build byte for byte)
        query = f"SELECT {select_cols} FROM images i {join_strs} ORDER BY i.id"
        self.read_cur.execute(query)
        while True:
            rows = self.read_cur.fetchmany(self.batch_size)
            if not rows:
                break
            yield rows

    # ---- decode (create_index.py:160-189) -------------------------------------------------------
    # Fast path for the blobs the reference's extractors write (create_vector_base.py:142-145):
    # pickle protocol 5 of a 1-D C-contiguous little-endian float32 ndarray under numpy >= 2 is
    #   ... '_frombuffer' MEMO STACK_GLOBAL MEMO MARK BYTEARRAY8 <len:8> <raw> ... 'f4' ... '<' ... (d,) 'C' ...
    # so the payload can be viewed in place instead of running the unpickler (4x faster per blob;
    # the build is bound by this Python loop).  Anything else falls back to pickle.loads.
    _FAST_MARK = b"_frombuffer\x94\x93\x94(\x96"
    _FAST_TAIL = b"\x8c\x01C\x94t\x94R\x94."

    @classmethod
    def _decode_blob(cls, blob) -> np.ndarray:
        if isinstance(blob, (bytes, bytearray, memoryview)) and blob[:2] == b"\x80\x05":
            head = bytes(blob[:80])
            m = head.find(cls._FAST_MARK)
            if m > 0:
                off = m + len(cls._FAST_MARK) + 8
                n_bytes = int.from_bytes(head[off - 8:off], "little")
                tail = bytes(blob[off + n_bytes:])
                d = n_bytes // 4
                shape = (b"K" + bytes([d]) if d < 256 else
                         b"M" + d.to_bytes(2, "little") if d < 65536 else b"J" + d.to_bytes(4, "little")) + b"\x85"
                if (n_bytes % 4 == 0 and 0 < len(tail) < 160 and b"\x8c\x02f4\x94" in tail
                        and b"\x8c\x01<\x94" in tail and tail.endswith(shape + b"\x94" + cls._FAST_TAIL)):
                    return np.frombuffer(blob, dtype="<f4", count=d, offset=off)
        vec = pickle.loads(blob)
        if hasattr(vec, "cpu"):
            vec = vec.cpu().numpy()
        return np.asarray(vec, dtype="float32").ravel()

    def _process_batch(self, rows):
        """Same contract as the reference: (ids, [concatenated float32 row per id]); a row with an
        undecodable blob is logged and skipped."""
        ids, parts = self._decode_batch(rows)
        return ids, [np.concatenate(p) for p in parts]

    def _decode_batch(self, rows):
        ids, parts_per_row = [], []
        for rec_id, *blobs in rows:
            parts = []
            skip = False
            for vt, blob in zip(self.vector_types, blobs):
                try:
                    parts.append(self._decode_blob(blob))
                except Exception as e:
                    self._log(f"ID {rec_id}: error loading {vt}: {e}", level="warning")
                    skip = True
                    break
            if skip:
                continue
            ids.append(rec_id)
            parts_per_row.append(parts)
        return ids, parts_per_row

    def find_valid_m(self, dim, candidates=(64, 56, 48, 32, 28, 24, 16, 12, 8)):
        # kept for API compatibility (create_index.py:191-205); product quantisation is not used
        for m in candidates:
            if dim % m == 0:
                return m
        return 1

    def _initialize_index(self, table_dims, capacity, use_pq=True):
        """reference: IndexIVFPQ over an HNSW coarse quantiser, or IndexHNSWFlat
        (create_index.py:207-234).  Here: one exact flat shard; M/efConstruction/efSearch/use_pq
        are accepted and ignored."""
        index = FlatShard(list(table_dims), int(capacity), device=self.device)
        self._log(
            f"Created exact flat GPU index (dims={list(table_dims)}, D={index.d}, capacity={capacity}; "
            f"hnsw_M={self.hnsw_M}, efConstruction={self.efConstruction}, efSearch={self.efSearch} ignored)",
            level="info",
        )
        return index

    def _build_native(self, total):
        """Fresh builds: the whole SQL -> decode -> stage -> pack loop runs in libb2k.so
        (b2k_ingest_sqlite).  The table dims come from the first row, decoded here.  Returns
        (index, ids), or (None, []) when some blob is not in the extractors' format — the Python
        decoder below then rebuilds from scratch (pickle.loads semantics, row skipping included)."""
        from image_recommender_b200 import B2KError, _capi
        select_cols, join_strs = self._make_select_and_joins()
        sql = f"SELECT {select_cols} FROM images i {join_strs} ORDER BY i.id"
        first = self.read_cur.execute(sql + " LIMIT 1").fetchone()
        if first is None:
            return None, []
        try:
            dims = [int(self._decode_blob(b).shape[0]) for b in first[1:]]
        except Exception:
            return None, []
        index = self._initialize_index(dims, total)
        try:
            threads = self._ingest_threads(total)
            if threads > 1:
                # several reader threads, each on its own connection over a range of image ids, committed in id
                # order: the same rows, offsets and ids as the single-threaded loop
                rows_per_slot = 2048
                bounds = self._id_chunk_bounds(rows_per_slot)
                sql_range = f"SELECT {select_cols} FROM images i {join_strs} WHERE i.id >= ?1 AND i.id < ?2 ORDER BY i.id"
                ids = index.ingest_sqlite_mt(self.db_path, sql_range, bounds, total, threads, rows_per_slot)
            else:
                ids = index.ingest_sqlite(self.db_path, sql, total)
        except B2KError as e:
            index.close()
            if e.status != _capi.E_UNSUPPORTED:
                raise
            self._log(f"Native ingest declined ({e}); decoding in Python.", level="warning")
            return None, []
        return index, [int(i) for i in ids]

    def _ingest_threads(self, rows=None, share=1):
        """Reader threads of the native ingest.  `ingest_threads` (constructor) or B2K_INGEST_THREADS are taken as
        given; the default is half the host cores (divided by `share` ranks), at most 8 — SQLite page reads and row
        copies scale until the memory bus does not — and one thread per 50 000 rows at least: pinning the extra
        staging slots costs more than a small database's whole ingest."""
        n = getattr(self, "ingest_threads", None)
        if n is None:
            n = int(os.environ.get("B2K_INGEST_THREADS", "0")) or None
        if n is None:
            n = max(1, min(8, len(os.sched_getaffinity(0)) // (2 * max(1, share))))
            if rows is not None:
                n = max(1, min(n, int(rows) // 50_000))
        return max(1, min(int(n), 16))

    def _id_chunk_bounds(self, rows_per_chunk):
        """Every rows_per_chunk-th image id (in id order) plus last id + 1: chunk c = ids in [b[c], b[c+1])."""
        cur = self.read_cur
        try:
            firsts = [r[0] for r in cur.execute(
                "SELECT id FROM (SELECT id, ROW_NUMBER() OVER (ORDER BY id) AS rn FROM images) "
                f"WHERE (rn - 1) % {int(rows_per_chunk)} = 0 ORDER BY id")]
        except sqlite3.OperationalError:          # SQLite < 3.25: no window functions
            firsts = [r[0] for r in cur.execute("SELECT id FROM images ORDER BY id")][::int(rows_per_chunk)]
        last = cur.execute("SELECT MAX(id) FROM images").fetchone()[0]
        if not firsts or last is None:
            return [0, 0]
        return [int(x) for x in firsts] + [int(last) + 1]

    @staticmethod
    def _dist_world():
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_world_size(), dist.get_rank()
        except Exception:
            pass
        return 1, 0

    def _build_sharded(self, world, rank):
        """Row-sharded build (ingest shards like the search): rank r takes rows
        [r0, r1) of the join in id order, packs them on its GPU, and all ranks write ONE index file
        (`b2k_save_shard`) whose offsets equal those of a single-GPU build.  A database too large for one
        GPU (BASELINE config 5: 100 M rows = 134 GB per GPU on eight) can only be built this way."""
        import torch
        import torch.distributed as dist
        from image_recommender_b200 import B2KError, _capi
        from image_recommender_b200.sharded import shard_range
        dev = torch.device("cuda", self.device)
        if rank == 0:
            if self.index_file.exists():
                self.index_file.unlink()
            self.write_cur.execute(f"DELETE FROM {self.offset_table}")
            self.write_conn.commit()
        dist.barrier()
        total = self._count_records()
        self._log(f"[rank {rank}/{world}] {total} complete records found.", level="info")
        if total == 0:
            return
        r0, r1 = shard_range(total, world, rank)
        select_cols, join_strs = self._make_select_and_joins()
        base_sql = f"SELECT {select_cols} FROM images i {join_strs}"
        # keyset page instead of LIMIT/OFFSET (which re-scans the r0 joined rows before the page on every rank):
        # the ids of the complete records, in order, give this rank's first and last id
        all_ids = [r[0] for r in self.read_cur.execute(f"SELECT i.id FROM images i {join_strs} ORDER BY i.id")]
        mt_bounds = None
        if r1 > r0:
            sql = f"{base_sql} WHERE i.id BETWEEN {int(all_ids[r0])} AND {int(all_ids[r1 - 1])} ORDER BY i.id"
            # reader threads of this rank (the host cores are shared by the ranks): chunks of 2048 complete records
            threads = self._ingest_threads(r1 - r0, share=world)
            if threads > 1:
                mt_bounds = [int(x) for x in all_ids[r0:r1:2048]] + [int(all_ids[r1 - 1]) + 1]
        else:
            sql = f"{base_sql} WHERE 0"
        del all_ids
        first = self.read_cur.execute(base_sql + " ORDER BY i.id LIMIT 1").fetchone()
        dims = [int(self._decode_blob(b).shape[0]) for b in first[1:]]
        index = self._initialize_index(dims, max(r1 - r0, 1))
        ids = None
        if self.native_ingest and r1 > r0:
            try:
                if mt_bounds is not None:
                    sql_range = f"{base_sql} WHERE i.id >= ?1 AND i.id < ?2 ORDER BY i.id"
                    ids = [int(i) for i in index.ingest_sqlite_mt(self.db_path, sql_range, mt_bounds, r1 - r0, threads, 2048)]
                else:
                    ids = [int(i) for i in index.ingest_sqlite(self.db_path, sql, r1 - r0)]
            except B2KError as e:
                if e.status != _capi.E_UNSUPPORTED:
                    raise
                self._log(f"[rank {rank}] native ingest declined ({e}); decoding in Python.", level="warning")
                index.reset()
        if ids is None:
            ids = []
            cur = self.read_conn.cursor()
            cur.execute(sql)
            while True:
                rows = cur.fetchmany(self.batch_size)
                if not rows:
                    break
                b_ids, parts = self._decode_batch(rows)
                if b_ids:
                    index.add_tables([np.stack([p[t] for p in parts]).astype("float32") for t in range(len(dims))])
                    ids.extend(b_ids)
        # rows skipped on decode errors shift the offsets of every later rank: exchange the counts
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        mine = torch.tensor([len(ids)], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, mine)
        counts = counts.cpu().tolist()
        begin, n_all = sum(counts[:rank]), sum(counts)
        for r in range(world):            # one writer at a time: the creating rank first, SQLite one process at a time
            if r == rank:
                index.save_shard(self.index_file, np.asarray(ids, dtype=np.int64), begin, n_all, create=(rank == 0))
                for lo in range(0, len(ids), self.batch_size):
                    self._store_offsets(ids[lo:lo + self.batch_size], begin + lo)
            dist.barrier()
        self._log(f"[rank {rank}] wrote rows [{begin}, {begin + len(ids)}) of {n_all} to {self.index_file}", level="info")
        index.close()
        self.read_conn.close()
        self.write_conn.close()

    def _store_offsets(self, ids, start_offset):
        pairs = [(rid, start_offset + i) for i, rid in enumerate(ids)]
        self.write_cur.executemany(
            f"INSERT OR REPLACE INTO {self.offset_table} (image_id, offset) VALUES (?, ?)",
            pairs,
        )
        self.write_conn.commit()

    # ---- build (create_index.py:251-325) ----------------------------------------------------------
    def build_index(self, update_index: bool = False):
        """One pass over the joined tables: decode -> per-table arrays -> GPU pack (normalise,
        concatenate, bf16) -> offsets.  The reference's separate "training" pass
        (create_index.py:283-299) has no counterpart: nothing is trained.

        update_index=True appends the images that have no offset yet to the existing index file
        (the reference's flag restarts at offset 0 and is unusable)."""
        combo = "_".join(self.vector_types)
        self._log(f"Starting index build for [{combo}]…", level="info")
        world, rank = self._dist_world()
        if world > 1 and self.sharded is not False:
            if update_index:
                raise ValueError("update_index is not supported for a row-sharded build")
            return self._build_sharded(world, rank)

        index = None
        all_ids: list[int] = []
        offset_counter = 0
        known = set()
        if update_index and self.index_file.exists():
            from image_recommender_b200 import file_info, load_ids
            info = file_info(self.index_file)
            if not info["has_ids"] and info["n_rows"] > 0:
                # without the id column (files written through faiss_shim.write_index) nothing tells which images
                # the file already holds: every row would be added again.  Refuse before any work is done.
                raise ValueError(f"{self.index_file} carries no image-id column: it cannot be updated in place; "
                                 f"rebuild it (update_index=False)")
            # straight into the final capacity: appending never re-allocates (a re-allocation holds the old and
            # the new arrays at once, which a shard filling most of the GPU cannot afford)
            index = FlatShard.load(self.index_file, device=self.device,
                                   capacity=max(info["n_rows"], self._count_records()))
            offset_counter = index.ntotal
            all_ids = load_ids(self.index_file, 0, info["n_rows"]).tolist() if info["has_ids"] else []
            known = set(all_ids)
            self._log(f"Appending to {self.index_file} ({offset_counter} vectors).", level="info")
        else:
            if self.index_file.exists():
                self._log(f"Removing existing index {self.index_file}", level="info")
                self.index_file.unlink()
            self._log(f"Clearing offset table {self.offset_table}", level="info")
            self.write_cur.execute(f"DELETE FROM {self.offset_table}")
            self.write_conn.commit()

        total = self._count_records()
        self._log(f"{total} complete records found.", level="info")
        if total == 0:
            self._log("No complete embeddings found; aborting.", level="error")
            return

        batch_num = 0
        native_done = False
        if index is None and self.native_ingest:
            index, ids = self._build_native(total)
            if index is not None:
                for lo in range(0, len(ids), self.batch_size):
                    self._store_offsets(ids[lo:lo + self.batch_size], lo)
                all_ids, offset_counter, native_done = ids, len(ids), True
                self._log(f"Native ingest: added {offset_counter} vectors.", level="info")
        for batch in (() if native_done else self._batch_records()):
            batch_num += 1
            ids, parts = self._decode_batch(batch)
            if known:
                keep = [i for i, rid in enumerate(ids) if rid not in known]
                ids = [ids[i] for i in keep]
                parts = [parts[i] for i in keep]
            if not ids:
                continue
            tables = [np.stack([p[t] for p in parts]).astype("float32") for t in range(len(self.vector_types))]
            if index is None:
                index = self._initialize_index([t.shape[1] for t in tables], total)
            index.add_tables(tables)
            self._store_offsets(ids, offset_counter)
            all_ids.extend(ids)
            offset_counter += len(ids)
            self._log(f"Batch {batch_num}: added {len(ids)} vectors (total {offset_counter}).", level="info")

        if index is None:
            self._log("No decodable embeddings found; aborting.", level="error")
            return
        self._log(f"Writing index to {self.index_file.resolve()}", level="info")
        index.save(self.index_file, np.asarray(all_ids, dtype=np.int64))
        self._log(f"Index saved ({index.ntotal} vectors).", level="info")
        index.close()

        self.read_conn.close()
        self.write_conn.close()
        self._log("Done.", level="info")


def main(argv=None):
    ap = argparse.ArgumentParser(description="Build the exact GPU index from the SQLite vector tables")
    ap.add_argument("--db-path", default="images.db")
    ap.add_argument("--vector-types", nargs="+", default=["color"], help="any of: color sift dreamsim")
    ap.add_argument("--output", default=None, help="index file (default index_hnsw_<types>.faiss)")
    ap.add_argument("--batch-size", type=int, default=8192)
    ap.add_argument("--hnsw_M", type=int, default=32)
    ap.add_argument("--efConstruction", type=int, default=200)
    ap.add_argument("--efSearch", type=int, default=64)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--update", action="store_true", help="append images that have no offset yet")
    a = ap.parse_args(argv)
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:       # torchrun --nproc-per-node G -m main.create_index ...: row-sharded build on G GPUs
        import torch
        import torch.distributed as dist
        a.device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(a.device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", a.device))
    builder = FAISSIndexBuilderDB(
        db_path=a.db_path, vector_types=a.vector_types, batch_size=a.batch_size, index_file=a.output,
        hnsw_M=a.hnsw_M, efConstruction=a.efConstruction, efSearch=a.efSearch, device=a.device,
    )
    builder.build_index(update_index=a.update)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
