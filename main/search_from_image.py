"""Top-k query for the retrieval hot path — drop-in for the reference's main/search_from_image.py.

`ImageRecommender` keeps the reference's constructor kwargs and method names
(/root/reference/main/search_from_image.py:17-379); `index.search` runs as exact brute-force
search on the GPU (`image_recommender_b200.FlatShard`) instead of a FAISS HNSW/IVFPQ walk, the
index is loaded once and kept resident (the reference re-reads it on every call) and
`faiss.normalize_L2` is the package's CUDA kernel.  Feature *extraction* (colour histogram,
SIFT-VLAD, DreamSim inference) is outside this repository's scope: query vectors come from the
SQLite vector tables exactly as the reference's cache path does (search_from_image.py:50-92).

CLI (README.md:110 of the reference):

    python -m main.search_from_image --db-path images.db --images-root image_data \
        --query path/to/query.jpg --index combo_color_sift_dreamsim --top-k 5
"""
from __future__ import annotations

import argparse
import itertools
import logging
import pickle
import sqlite3
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from image_recommender_b200 import FlatShard, file_info, load_ids, normalize_L2, parse_f32_blob  # noqa: E402

VALID_TYPES = ["color", "hog", "lpips", "dreamsim", "sift", "color_sift", "sift_dreamsim"]
TABLES = {  # type -> (table, column)   (create_db.py:59-85)
    "color": ("color_vectors", "color_vector_blob"),
    "sift": ("sift_vectors", "sift_vector_blob"),
    "dreamsim": ("dreamsim_vectors", "dreamsim_vector_blob"),
}


class ImageRecommender:
    def __init__(
        self,
        images_root="image_data",
        db_path="images.db",
        use_gpu=True,
        sift_codebook_path="sift_codebook.npy",
        sift_pca_path="sift_vlad_pca.joblib",
        sift_n_clusters=256,
        sift_desc_dim=128,
        top_k=5,
        index_dir=".",
        device=0,
    ):
        self.base_dir = Path().expanduser().resolve()
        self.images_root = (self.base_dir / images_root).resolve()
        self.db_path = Path(db_path).expanduser().resolve()
        if not use_gpu:
            raise ValueError("this engine is GPU-only (no CPU search path); use_gpu=False is not supported")
        self.use_gpu = use_gpu
        # device: a CUDA ordinal, or "all": the index is row-sharded over every GPU of the box and driven from this
        # one process (image_recommender_b200.ShardGroup; no torchrun)
        self.device = 0 if device == "all" else int(device)
        self.all_devices = device == "all"
        # extractor settings are accepted for signature compatibility only
        self.sift_codebook_path = Path(sift_codebook_path).expanduser().resolve()
        self.sift_pca_path = Path(sift_pca_path).expanduser().resolve()
        self.sift_n_clusters = sift_n_clusters
        self.sift_desc_dim = sift_desc_dim
        self.top_k = top_k
        self.index_dir = Path(index_dir)
        self._resident = {}      # index file -> (mtime, FlatShard)
        self._resident_ids = {}  # id(FlatShard) -> image id per offset (from the index file), when stored
        self._conn = None        # one SQLite connection for the batched lookups (the reference opens one per lookup)
        logging.basicConfig(level=logging.INFO, format="%(asctime)s [%(levelname)s] %(message)s")

    # ---- query-vector acquisition (search_from_image.py:50-216) -------------------------------
    def _get_db_vector(self, path_rel: str, vector_table: str, vector_column: str):
        conn = sqlite3.connect(self.db_path)
        try:
            cur = conn.cursor()
            row = cur.execute("SELECT id FROM images WHERE path = ?", (path_rel,)).fetchone()
            if not row:
                # create_db.py:99 stores paths relative to the PARENT of the image folder
                row = cur.execute("SELECT id FROM images WHERE path = ?",
                                  (f"{self.images_root.name}/{path_rel}",)).fetchone()
            if not row:
                return None
            vrow = cur.execute(f"SELECT {vector_column} FROM {vector_table} WHERE image_id = ?", (row[0],)).fetchone()
        finally:
            conn.close()
        if not vrow or vrow[0] is None:
            return None
        return self._decode_cached_blob(vrow[0], vector_column, path_rel)

    @staticmethod
    def _decode_cached_blob(blob, vector_column="", path_rel=""):
        """Blob -> [1, d] array with the reference's rules (search_from_image.py:70-91): unpickle
        (tensors moved to numpy), else raw little-endian float32."""
        fast = parse_f32_blob(blob) if isinstance(blob, bytes) else None
        if fast is not None:          # pickled 1-D float32 ndarray: the same floats, without the unpickler
            return fast.reshape(1, -1)
        try:
            arr = pickle.loads(blob)
            if hasattr(arr, "cpu"):
                arr = arr.cpu().numpy()
            if isinstance(arr, np.ndarray):
                return arr.reshape(1, -1) if arr.ndim == 1 else arr
        except Exception:
            pass
        if len(blob) % 4 != 0:       # raw little-endian float32 fallback (search_from_image.py:82-91)
            logging.warning(f"invalid blob size {len(blob)} B for '{vector_column}' at '{path_rel}', skipping cache")
            return None
        return np.frombuffer(blob, dtype="float32").reshape(1, -1)

    def get_or_compute_vector(self, path_rel, vector_table, vector_column, compute_func=None, reshape=None,
                              print_vectors=False):
        cached = self._get_db_vector(path_rel, vector_table, vector_column)
        if cached is not None:
            return cached
        vec = compute_func() if compute_func is not None else None
        if vec is None:
            logging.error(f"No cached {vector_column} for '{path_rel}' (feature extraction is out of scope here: "
                          f"run the reference's vector_scripts first).")
            return None
        return vec.reshape(*reshape) if reshape is not None else vec

    def extract_color_features(self, path_rel):
        return self.get_or_compute_vector(path_rel, *TABLES["color"], reshape=(1, -1))

    def extract_sift_vlad_features(self, path_rel):
        return self.get_or_compute_vector(path_rel, *TABLES["sift"])

    def extract_dreamsim_features(self, path_rel):
        return self.get_or_compute_vector(path_rel, *TABLES["dreamsim"])

    # ---- main search (search_from_image.py:219-254) ---------------------------------------------
    def search_similar_images(self, query_image_paths, index_type: str = "color", plot: bool = False):
        """Returns [(path, squared-L2 distance)] ascending, as the reference's `_fetch_results`
        (which it then only plots)."""
        paths_rel = [self._relative(p) for p in query_image_paths]
        ordered = self._get_ordered_index_types(index_type)
        if not ordered:
            return None
        index, offset_table, file_order = self._load_faiss_index("_".join(ordered), ordered)
        if index is None:
            return None
        # same arithmetic as _extract_query_vector / _fetch_results (kept below in the reference's shape and
        # checked equal), through IN-selects on one persistent connection and the resident id column:
        # 0.8 ms -> 0.3 ms of host time per query, next to 0.9 ms of GPU time on eight GPUs
        parts, offs, live = self._extract_query_parts([paths_rel], file_order)
        if parts is None:
            return None
        # mean over the query images, faiss.normalize_L2 and the search run on the device in one call
        # (b2k_search_groups): the normalised vector never travels back to the host
        distances, indices = index.search_groups(parts, offs, self.top_k)
        results = self._fetch_results_batch(indices, distances, offset_table, index)[0]
        if not results:
            logging.error("No similar images found.")
            return None
        if plot:
            self._plot_results(query_image_paths, results)
        return results

    def search_batch(self, query_groups, index_type: str = "color"):
        """New: one query vector per group of image paths, all groups searched as ONE
        batch.  Same arithmetic as `search_similar_images` per group (concat -> mean over the group's
        images -> whole-vector normalise -> exact top-k -> offset -> path), but the SQLite work is
        batched: cached vectors are fetched with `IN (...)` selects on one connection instead of
        three selects per image, and hits are mapped through the id column kept in the index file
        instead of two selects per hit.  Returns one result list per group (None for a group
        without any usable image)."""
        ordered = self._get_ordered_index_types(index_type)
        if not ordered:
            return None
        index, offset_table, file_order = self._load_faiss_index("_".join(ordered), ordered)
        if index is None:
            return None
        groups_rel = [[self._relative(p) for p in g] for g in query_groups]
        parts, offs, live = self._extract_query_parts(groups_rel, file_order)
        out = [None] * len(groups_rel)
        if parts is None:
            return out
        distances, indices = index.search_groups(parts, offs, self.top_k)
        for row, res in zip(live, self._fetch_results_batch(indices, distances, offset_table, index)):
            out[row] = res or None
        return out

    # ---- batched SQLite access (new) --------------------------------------------------------------
    _SQL_CHUNK = 900        # bound variables per statement (SQLite's historical limit is 999)

    def _db(self):
        if self._conn is None:
            self._conn = sqlite3.connect(self.db_path, timeout=60)
        return self._conn

    def _select_in(self, cur, sql_fmt: str, keys):
        """Runs sql_fmt.format(marks=...) over `keys` in chunks; yields rows."""
        keys = list(keys)
        for lo in range(0, len(keys), self._SQL_CHUNK):
            part = keys[lo:lo + self._SQL_CHUNK]
            yield from cur.execute(sql_fmt.format(marks=",".join("?" * len(part))), part)

    def _fetch_vectors_batch(self, paths_rel, ordered):
        """{path_rel: [1, D] float32 concat of the cached parts in `ordered`} for every path whose
        parts are all cached (the per-image rules of _get_db_vector / _extract_query_vector)."""
        uniq = list(dict.fromkeys(paths_rel))
        prefixed = {p: f"{self.images_root.name}/{p}" for p in uniq}
        conn = self._db()
        try:
            cur = conn.cursor()
            by_path = dict((path, i) for i, path in self._select_in(
                cur, "SELECT id, path FROM images WHERE path IN ({marks})", list(uniq) + list(prefixed.values())))
            ids = {p: by_path.get(p, by_path.get(prefixed[p])) for p in uniq}      # exact path first (reference order)
            want = sorted({i for i in ids.values() if i is not None})
            blobs = {}
            for t in ordered:
                table, col = TABLES[t]
                blobs[t] = dict(self._select_in(cur, f"SELECT image_id, {col} FROM {table} WHERE image_id IN ({{marks}})", want))
        finally:
            cur.close()
        out = {}
        for p in uniq:
            parts = []
            for t in ordered:
                blob = blobs[t].get(ids[p]) if ids[p] is not None else None
                v = self._decode_cached_blob(blob, TABLES[t][1], p) if blob is not None else None
                if v is None:
                    logging.error(f"No cached {TABLES[t][1]} for '{p}' (feature extraction is out of scope here: "
                                  f"run the reference's vector_scripts first).")
                    parts = None
                    break
                parts.append(v.reshape(1, -1) if v.ndim == 1 else v)
            if parts is None:
                logging.error(f"Could not load all vector features for '{p}', skipping this image.")
                continue
            out[p] = np.concatenate(parts, axis=1).astype("float32")
        return out

    def _extract_query_parts(self, groups_rel, ordered):
        """([n_images, D] concatenated per-image vectors, int32 group offsets [G'+1], indices of the groups
        kept) — the inputs of `search_groups`, which takes the mean of every group, normalises it and searches
        on the device.  Groups without any usable image are dropped (logged), as _extract_query_vector does."""
        if any(t not in TABLES for t in ordered):
            logging.error(f"Unknown vector type in {ordered}.")
            return None, None, []
        vecs = self._fetch_vectors_batch([p for g in groups_rel for p in g], ordered)
        rows, offs, live = [], [0], []
        for gi, g in enumerate(groups_rel):
            have = [vecs[p] for p in g if p in vecs]
            if not have:
                logging.error("Could not extract a vector for any of the query images.")
                continue
            rows.extend(have)
            offs.append(offs[-1] + len(have))
            live.append(gi)
        if not live:
            return None, None, []
        parts = np.ascontiguousarray(np.concatenate(rows, axis=0), dtype=np.float32)
        return parts, np.asarray(offs, dtype=np.int32), live

    def _extract_query_matrix(self, groups_rel, ordered):
        """([G', D] normalised query matrix, indices of the groups it holds) — row g is bit-identical
        to _extract_query_vector(groups_rel[g], ordered)."""
        if any(t not in TABLES for t in ordered):
            logging.error(f"Unknown vector type in {ordered}.")
            return None, []
        vecs = self._fetch_vectors_batch([p for g in groups_rel for p in g], ordered)
        rows, live = [], []
        for gi, g in enumerate(groups_rel):
            have = [vecs[p] for p in g if p in vecs]
            if not have:
                logging.error("Could not extract a vector for any of the query images.")
                continue
            rows.append(np.mean(have, axis=0))
            live.append(gi)
        if not rows:
            return None, []
        q = np.ascontiguousarray(np.concatenate(rows, axis=0), dtype=np.float32)
        normalize_L2(q, device=self.device)
        return q, live

    def _fetch_results_batch(self, indices, distances, offset_table, index=None):
        """Per query row: [(path, distance)] ascending, as _fetch_results; offsets -> image ids through
        the id column of the index file when it has one (else one IN-select on the offset table),
        ids -> paths with one IN-select per chunk."""
        offs = sorted({int(o) for o in np.asarray(indices).ravel() if o >= 0})
        ids_resident = self._resident_ids.get(id(index)) if index is not None else None
        conn = self._db()
        try:
            cur = conn.cursor()
            if ids_resident is not None:
                off2id = {o: int(ids_resident[o]) for o in offs if o < len(ids_resident)}
            else:
                off2id = dict(self._select_in(cur, f"SELECT offset, image_id FROM {offset_table} WHERE offset IN ({{marks}})", offs))
            id2path = dict(self._select_in(cur, "SELECT id, path FROM images WHERE id IN ({marks})", sorted(set(off2id.values()))))
        finally:
            cur.close()
        out = []
        for qi in range(len(indices)):
            res = []
            for rank, offset in enumerate(indices[qi]):
                if offset < 0:
                    continue
                image_id = off2id.get(int(offset))
                if image_id is None:
                    logging.warning(f"No entry found for offset={offset} in {offset_table}")
                    continue
                path = id2path.get(image_id)
                if path is None:
                    logging.warning(f"No path found for id={image_id}")
                    continue
                res.append((Path(self.base_dir) / path, float(distances[qi, rank])))
            res.sort(key=lambda x: x[1])
            out.append(res)
        return out

    def _relative(self, p):
        # search_from_image.py:230-232
        return Path(p).resolve().relative_to(self.images_root).as_posix()

    def _get_ordered_index_types(self, index_type: str):
        # accepts README's `combo_color_sift_dreamsim` as well as the reference's `color,sift`
        text = index_type.lower()
        if text.startswith("combo_"):
            text = text[len("combo_"):].replace("_", ",")
        requested = [x.strip() for x in text.split(",")]
        ordered = [v for v in VALID_TYPES if v in requested]
        if not ordered:
            logging.error(f"Unknown index_type '{index_type}'. Choose from {VALID_TYPES}.")
            return []
        return ordered

    def _extract_query_vector(self, paths_rel, ordered):
        all_query_vectors = []
        for path_rel in paths_rel:
            parts = []
            for vec_type in ordered:
                if vec_type == "color":
                    parts.append(self.extract_color_features(path_rel))
                elif vec_type == "sift":
                    parts.append(self.extract_sift_vlad_features(path_rel))
                elif vec_type == "dreamsim":
                    parts.append(self.extract_dreamsim_features(path_rel))
                else:
                    logging.error(f"Unknown vector type '{vec_type}' for '{path_rel}'.")
                    return None
            if any(p is None for p in parts):
                logging.error(f"Could not load all vector features for '{path_rel}', skipping this image.")
                continue
            parts = [x.reshape(1, -1) if x.ndim == 1 else x for x in parts]
            all_query_vectors.append(np.concatenate(parts, axis=1).astype("float32"))
        if not all_query_vectors:
            logging.error("Could not extract a vector for any of the query images.")
            return None
        combined_vec = np.ascontiguousarray(np.mean(all_query_vectors, axis=0), dtype=np.float32)
        if combined_vec.ndim == 1:
            combined_vec = combined_vec.reshape(1, -1)
        normalize_L2(combined_vec, device=self.device)      # faiss.normalize_L2 (search_from_image.py:322)
        return combined_vec

    def _load_faiss_index(self, canonical, ordered=None):
        """(index, offset table, concat order of that file).  The reference looks only for
        index_hnsw_<fixed-order>.faiss, which never matches an index built in another order
        (SURVEY F6); here every ordering of the requested types is tried, fixed order first."""
        ordered = ordered or canonical.split("_")
        candidates = [tuple(ordered)] + [p for p in itertools.permutations(ordered) if p != tuple(ordered)]
        for order in candidates:
            name = "_".join(order)
            f = self.index_dir / f"index_hnsw_{name}.faiss"
            if not f.exists():
                continue
            try:
                mtime = f.stat().st_mtime
                hit = self._resident.get(str(f))
                if hit is None or hit[0] != mtime:
                    if hit is not None:
                        self._resident_ids.pop(id(hit[1]), None)
                        hit[1].close()
                    shard = self._load_index_file(f)
                    self._resident[str(f)] = (mtime, shard)
                    info = file_info(f)
                    if info["has_ids"]:
                        self._resident_ids[id(shard)] = load_ids(f, 0, info["n_rows"])
                    logging.info(f"Loaded index '{f}' with {self._resident[str(f)][1].ntotal} vectors.")
                return self._resident[str(f)][1], f"faiss_index_offsets_{name}", list(order)
            except Exception as e:
                logging.error(f"Error loading index '{f}': {e}")
                return None, None, None
        logging.error(f"Error loading index 'index_hnsw_{canonical}.faiss': no such file in {self.index_dir}")
        return None, None, None

    def _load_index_file(self, f):
        """One GPU: the whole file.  Under torchrun (torch.distributed initialised, world > 1): this
        rank's row range, searched SPMD with a top-k exchange (image_recommender_b200.sharded) — every
        rank must then issue the same queries."""
        try:
            import torch.distributed as dist
            sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        except Exception:
            sharded = False
        if not sharded:
            if self.all_devices:
                from image_recommender_b200 import ShardGroup
                return ShardGroup.load(f)
            return FlatShard.load(f, device=self.device)
        from image_recommender_b200.sharded import ShardedIndex
        return ShardedIndex.load(f, device=self.device)

    def _fetch_results(self, indices, distances, offset_table):
        conn = sqlite3.connect(self.db_path)
        cur = conn.cursor()
        results = []
        for rank, offset in enumerate(indices[0]):
            if offset < 0:
                continue                      # -1 padding when top_k > ntotal
            row = cur.execute(f"SELECT image_id FROM {offset_table} WHERE offset = ?", (int(offset),)).fetchone()
            if not row:
                logging.warning(f"No entry found for offset={offset} in {offset_table}")
                continue
            fp_row = cur.execute("SELECT path FROM images WHERE id = ?", (row[0],)).fetchone()
            if not fp_row:
                logging.warning(f"No path found for id={row[0]}")
                continue
            results.append((Path(self.base_dir) / fp_row[0], float(distances[0, rank])))
        conn.close()
        results.sort(key=lambda x: x[1])
        return results

    def _plot_results(self, query_image_paths, results):
        """matplotlib display of the reference (search_from_image.py:381-427); optional here."""
        try:
            import matplotlib.pyplot as plt
            from PIL import Image
        except Exception as e:          # plotting is not part of the hot path
            logging.warning(f"plotting skipped ({e})")
            return
        items = [(p, f"Query: {Path(p).name}") for p in query_image_paths] + \
                [(fp, f"{Path(fp).name}\nDist: {d:.4f}") for fp, d in results]
        ncols = max(4, len(query_image_paths))
        nrows = -(-len(items) // ncols)
        fig, axes = plt.subplots(nrows, ncols, figsize=(5 * ncols, 5 * nrows))
        axes = np.atleast_1d(axes).flatten()
        for ax in axes:
            ax.axis("off")
        for ax, (p, title) in zip(axes, items):
            try:
                ax.imshow(Image.open(p).convert("RGB"))
                ax.set_title(title)
            except Exception as e:
                logging.error(f"Error rendering {p}: {e}")
        plt.tight_layout(pad=2.0, h_pad=3.0)
        plt.show()

    def close(self):
        for _, ix in self._resident.values():
            ix.close()
        self._resident.clear()
        self._resident_ids.clear()
        if self._conn is not None:
            self._conn.close()
            self._conn = None


def main(argv=None):
    ap = argparse.ArgumentParser(description="Exact top-k image search on the GPU index")
    ap.add_argument("--db-path", default="images.db")
    ap.add_argument("--images-root", default="image_data")
    ap.add_argument("--query", nargs="+", help="query image path(s); several are averaged")
    ap.add_argument("--serve", action="store_true",
                    help="resident mode: load the index once, then answer one query per stdin line (paths separated "
                         "by spaces are averaged; results end with an empty line) until EOF")
    ap.add_argument("--index", default="color", help="e.g. combo_color_sift_dreamsim or color,sift")
    ap.add_argument("--top-k", type=int, default=5)
    ap.add_argument("--index-dir", default=".")
    ap.add_argument("--device", default="0", help="CUDA ordinal, or 'all': row-shard the index over every GPU of the "
                                                     "box from this one process (no torchrun)")
    ap.add_argument("--plot", action="store_true")
    a = ap.parse_args(argv)
    if not a.serve and not a.query:
        ap.error("--query is required (or --serve)")
    import os
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:       # torchrun --nproc-per-node G -m main.search_from_image ...: row-sharded over G GPUs
        import torch
        import torch.distributed as dist
        a.device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(a.device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", a.device))
    rec = ImageRecommender(images_root=a.images_root, db_path=a.db_path, top_k=a.top_k, index_dir=a.index_dir,
                           device=a.device if a.device == "all" else int(a.device))
    def answer(paths):
        res = rec.search_similar_images(paths, index_type=a.index, plot=a.plot and rank == 0)
        if rank == 0:
            for fp, dist_ in res or []:
                print(f"{dist_:.6f}\t{fp}")
        return res

    results = answer(a.query) if a.query else None
    if a.serve:
        # the index (and its id column) stays on the GPU between queries: SURVEY §8f-3.  Under torchrun every
        # rank must see the same lines (give all ranks the same stdin, e.g. a file redirected into torchrun).
        if rank == 0:
            print(flush=True)
        for line in sys.stdin:
            paths = line.split()
            if not paths:
                continue
            try:
                results = answer(paths)
            except Exception as e:          # a bad path must not take the server down
                logging.error(f"query failed: {e}")
                results = None
            if rank == 0:
                print(flush=True)
        results = results or True
    rec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if results else 1


if __name__ == "__main__":
    sys.exit(main())
