import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the generator, the oracle and the CUDA library once per session."""
    from audio_formats_b200 import build as b
    from audio_formats_b200 import synth
    import oracle

    synth.build()
    oracle.build()
    b.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    import audio_formats_b200 as af

    if af.device_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests must run on the GPU box (there is no CPU fallback)")
    c = af.Context(0)
    yield c
    c.close()
