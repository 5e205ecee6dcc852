import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The product library and the oracle are built in-tree (prebuilt files are reused as they are)."""
    from rheotool_b200 import abi, build
    if not abi.lib_path().exists():
        build.build_library()
    if not (ROOT / "oracle" / "_build" / "liboracle.so").exists():
        from oracle import oracle as orc
        orc.build()
    yield


def has_gpu() -> bool:
    from rheotool_b200 import abi
    return abi.lib().rheo_gpu_device_count() > 0
