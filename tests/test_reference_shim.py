"""The one-line integration of INTEGRATION.md §2 on the UNMODIFIED reference file: with
`import image_recommender_b200.faiss_shim as faiss` standing in for faiss, the reference's own
FAISSIndexBuilderDB runs its SQL / decode / stack code unchanged and reaches the engine exactly at
index.add.  /root/reference only exists in the build container, so this test skips elsewhere; without a GPU
the engine must refuse at that seam (no CPU path); with one, the reference builds a real index."""
import importlib.util
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference/main/create_index.py")
GOLD = Path(__file__).resolve().parent / "golden"


def _load_reference_builder(monkeypatch):
    import image_recommender_b200.faiss_shim as shim
    monkeypatch.setitem(sys.modules, "faiss", shim)          # the one-line patch, applied from outside
    spec = importlib.util.spec_from_file_location("ref_create_index", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not REF.exists(), reason="the reference checkout is only present in the build container")
def test_unmodified_reference_builder_reaches_the_engine_at_index_add(tmp_path, monkeypatch):
    import image_recommender_b200 as irb
    shutil.copy(GOLD / "ref_fixture.db", tmp_path / "images.db")
    monkeypatch.chdir(tmp_path)
    mod = _load_reference_builder(monkeypatch)
    b = mod.FAISSIndexBuilderDB(db_path="images.db", vector_types=["color"], batch_size=8, log_dir="logs")
    assert b.offset_table == "faiss_index_offsets_color" and str(b.index_file) == "index_hnsw_color.faiss"
    if irb.device_count() == 0:
        with pytest.raises(irb.B2KError) as e:
            b.build_index()
        assert e.value.status == irb._capi.E_NODEVICE          # refused at index.add: there is no CPU engine
        return
    b.build_index()                                            # on a B200: the reference builds through the shim
    G = np.load(GOLD / "ref_golden.npz")
    info = irb.file_info("index_hnsw_color.faiss")
    assert info["n_rows"] == int(G["color/count"]) and info["table_dims"] == [48]
