import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    # Build what is missing (nvcc cross-compiles here; on the GPU box the .so files travel).
    # load the build script by path: importing the package itself requires the built library
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2k_build_ext", ROOT / "image_recommender_b200" / "build_ext.py")
    build_ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build_ext)
    if build_ext.needs_build():
        build_ext.build()
    import oracle
    oracle.build_oracle()


def _has_gpu() -> bool:
    try:
        import image_recommender_b200 as irb
        return irb.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _has_gpu():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    return 0


DIMS = [48, 128, 1792]   # color, sift, dreamsim
