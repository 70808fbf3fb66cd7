import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def eng_oracle():
    """Oracle model on eng.aspell.lexicon + simple.alphabet.tsv (BASELINE config 1 model)."""
    import workloads
    from oracle import orc
    m = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build()
    return m
