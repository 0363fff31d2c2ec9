"""torchrun check of the row-sharded build (main/create_index.py under torch.distributed): every rank ingests
its row range on its own GPU and all write one index file; file and offset table must equal a single-GPU build.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded_build.py"""
import json, os, shutil, sqlite3, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.distributed as dist
from test_ingest import _make_db
from main.create_index import FAISSIndexBuilderDB

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tmp = Path(tempfile.gettempdir()) / "b2k_check_sharded_build"
res = {"world": world}
ok = True
for case, foreign in (("clean", None), ("foreign_blob", ("sift", 4000)), ("reader_threads", None)):
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True); tmp.mkdir(parents=True)
        _make_db(tmp / "images.db", 6001, seed=7, foreign_at=foreign, missing={("sift", 9), ("color", 3000), ("dreamsim", 6001)})
    dist.barrier()
    types = ["color", "sift", "dreamsim"]
    b = FAISSIndexBuilderDB(db_path=str(tmp / "images.db"), vector_types=types, index_file=str(tmp / "sharded.faiss"),
                            log_dir=str(tmp / f"logs{rank}"), batch_size=700, device=local,
                            ingest_threads=3 if case == "reader_threads" else None)
    b._log = lambda m, level="info": None
    b.build_index()
    dist.barrier()
    if rank == 0:
        conn = sqlite3.connect(tmp / "images.db")
        tab_sh = conn.execute(f"SELECT image_id, offset FROM {b.offset_table} ORDER BY offset").fetchall()
        conn.close()
        s = FAISSIndexBuilderDB(db_path=str(tmp / "images.db"), vector_types=types, index_file=str(tmp / "single.faiss"),
                                log_dir=str(tmp / "logs_single"), batch_size=700, device=local, sharded=False)
        s._log = lambda m, level="info": None
        s.build_index()
        conn = sqlite3.connect(tmp / "images.db")
        tab_1 = conn.execute(f"SELECT image_id, offset FROM {s.offset_table} ORDER BY offset").fetchall()
        conn.close()
        same_file = (tmp / "sharded.faiss").read_bytes() == (tmp / "single.faiss").read_bytes()
        res[case] = {"rows": len(tab_1), "same_file": same_file, "same_offsets": tab_sh == tab_1}
        ok = ok and same_file and tab_sh == tab_1 and len(tab_1) == 5998
    dist.barrier()
    if case == "clean":
        # the reference-shaped query API on the sharded index: every rank loads its row range of the file just
        # built, issues the same query, and gets the single-GPU answer
        from main.search_from_image import ImageRecommender
        import image_recommender_b200 as irb
        os.chdir(tmp)
        shutil.copy(tmp / "sharded.faiss", tmp / "index_hnsw_color_sift_dreamsim.faiss") if rank == 0 else None
        dist.barrier()
        rec = ImageRecommender(images_root="image_data", db_path=str(tmp / "images.db"), top_k=7, index_dir=str(tmp), device=local)
        qpaths = [str(tmp / "image_data" / "000123.jpg"), str(tmp / "image_data" / "004321.jpg")]
        got = [rec.search_similar_images([p], index_type="combo_color_sift_dreamsim") for p in qpaths]
        got.append(rec.search_similar_images(qpaths, index_type="color,sift,dreamsim"))
        mine = json.dumps([[(str(p), d) for p, d in g] for g in got])
        alls = [None] * world
        dist.all_gather_object(alls, mine)
        same_on_all_ranks = all(a == alls[0] for a in alls)
        self_first = got[0][0][0].name == "000123.jpg" and got[1][0][0].name == "004321.jpg"
        single_ok = True
        if rank == 0:
            whole = irb.FlatShard.load(tmp / "single.faiss", device=local)
            qv = rec._extract_query_vector([rec._relative(qpaths[0])], ["color", "sift", "dreamsim"])
            d1, l1 = whole.search(qv, 7)
            single_ok = [round(float(x), 6) for x in d1[0]] == [round(d, 6) for _, d in got[0]]
            whole.close()
            res["recommender"] = {"same_on_all_ranks": same_on_all_ranks, "self_first": self_first, "equals_single_gpu": single_ok,
                                  "kind": type(next(iter(rec._resident.values()))[1]).__name__}
        ok = ok and same_on_all_ranks and self_first and single_ok
        rec.close()
        dist.barrier()
flag = torch.tensor([int(ok)], device="cuda")
dist.broadcast(flag, 0)
if rank == 0:
    print(json.dumps(res), flush=True)
    shutil.rmtree(tmp, ignore_errors=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
