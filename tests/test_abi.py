"""C-ABI surface: the library loads on a GPU-less host, exports every symbol include/b2k.h
declares, and its compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "b2k.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2k_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from image_recommender_b200 import _capi
    lib = _capi.load_library()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b2k.h but not exported by libb2k.so"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_capi.SIGNATURES) == set(names)
    assert lib.b2k_abi_version() == 2


def test_stats_struct_matches_header():
    from image_recommender_b200 import _capi
    text = (ROOT / "include" / "b2k.h").read_text()
    body = re.search(r"typedef struct b2k_stats \{(.*?)\} b2k_stats;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(?:int32_t|float)\s+(\w+);", body)
    assert fields == [f for f, _ in _capi.Stats._fields_]


def test_no_cpu_fallback_without_device():
    import image_recommender_b200 as irb
    if irb.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(irb.B2KError) as e:
        irb.FlatShard([48, 128, 1792], 10)
    assert e.value.status == -4
    x = np.ones((2, 8), np.float32)
    with pytest.raises(irb.B2KError):
        irb.normalize_L2(x)
    assert (x == 1).all()


def test_argument_validation():
    from image_recommender_b200 import _capi
    lib = _capi.load_library()
    h = C.c_void_p()
    dims = (C.c_int32 * 1)(0)
    assert lib.b2k_create(dims, 1, 10, 0, 0, C.byref(h)) == _capi.E_INVALID
    assert lib.b2k_create(dims, 9, 10, 0, 0, C.byref(h)) == _capi.E_INVALID
    assert b"bad argument" in lib.b2k_last_error()
    assert lib.b2k_file_info(b"/nonexistent/x.faiss", None, None, None, None) == _capi.E_IO


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package or main/ may reference it."""
    for base in ("image_recommender_b200", "main"):
        for p in (ROOT / base).rglob("*.py"):
            src = p.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
    # the CUDA sources may NAME the oracle in comments (the shared bit-level specs live in oracle/b2k_oracle.c),
    # but never include, link or dlopen anything of it
    for p in (ROOT / "image_recommender_b200" / "csrc").glob("*"):
        src = p.read_text()
        code = re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", src, flags=re.S))
        assert "oracle" not in code, p
        assert not re.search(r"#\s*include\s*[\"<][^\">]*oracle", src), p
    build = (ROOT / "image_recommender_b200" / "build_ext.py").read_text()
    assert "oracle" not in build


def _build_c_demo(tmp_path):
    import subprocess
    exe = tmp_path / "c_abi_demo"
    lib = ROOT / "image_recommender_b200" / "libb2k.so"
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-std=c99", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "c_abi_demo.c"),
                        "-o", str(exe), str(lib), "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe, {"LD_LIBRARY_PATH": str(lib.parent)}


def test_header_is_plain_c_and_links(tmp_path):
    """include/b2k.h compiles as C99 and a C client links against libb2k.so (no torch / Python types at
    the boundary); without a GPU the client stops at b2k_device_count."""
    import os
    import subprocess
    import image_recommender_b200 as irb
    exe, env = _build_c_demo(tmp_path)
    if irb.device_count() > 0:
        pytest.skip("a GPU is visible: the gpu-marked twin runs the client")
    r = subprocess.run([str(exe), "100", "2", "5"], env={**os.environ, **env}, capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_client_end_to_end(tmp_path):
    """examples/c_abi_demo.c: add in two batches, search, brute-force check in C, save/load round trip, the same file
    as a row-sharded group (b2k_group_*) from the one C thread."""
    import os
    import subprocess
    exe, env = _build_c_demo(tmp_path)
    for args in (["20000", "9", "10"], ["3000", "130", "5"], ["5000", "3", "40"]):
        r = subprocess.run([str(exe), *args], env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout, r.stderr)


def test_faiss_shim_refuses_rows_of_unequal_norm_without_table_dims():
    """without table_dims the shim would normalise whole rows and rank by cosine — a different ranking
    from the reference's L2 index when the row norms differ.  It refuses before touching the GPU."""
    import image_recommender_b200.faiss_shim as faiss
    index = faiss.IndexHNSWFlat(8, 32)
    x = np.ones((4, 8), np.float32)
    x[2] *= 3.0
    with pytest.raises(ValueError, match="norms"):
        index.add(x)
    assert index.ntotal == 0
