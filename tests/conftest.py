import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_ref():
    """The strongest oracle available: the reference's own decoders (oracle/_ref/libmspack_ref.so, built
    from /root/reference where it exists and shipped prebuilt to the GPU box), else the plain-C port."""
    from oracle import oracle as orc
    return orc.load("reference")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import oracle as orc
    return orc.Oracle("port")


@pytest.fixture(scope="session")
def decoder():
    from libmspack_b200.codec import BatchDecoder
    d = BatchDecoder(0)
    yield d
    d.close()
