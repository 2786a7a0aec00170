"""Shared fixtures.  GPU tests are marked `@pytest.mark.gpu`; everything else runs on CPU."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "birdnet-stm32_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

FIXTURES = os.path.join(ROOT, "tests", "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")
TFLITE = os.path.join(FIXTURES, "birdnet_stm32n6_100.tflite")
CONFIG = os.path.join(FIXTURES, "birdnet_stm32n6_100_model_config.json")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cfg():
    with open(CONFIG) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def graph():
    from birdnet_stm32.conversion.tflite_reader import read_tflite

    return read_tflite(TFLITE)


@pytest.fixture(scope="session")
def blob(graph, cfg):
    from birdnet_stm32.conversion.export_blob import export_blob

    return export_blob(graph, cfg)


@pytest.fixture(scope="session")
def oracle_model(blob):
    """The C oracle, running on requantisation constants the oracle derived itself from the `.tflite` scales
    (`oracle/tflite_quant.py`), not on the integers the product's exporter wrote into the blob."""
    from oracle import bn_oracle, tflite_quant

    bn_oracle.build()
    return bn_oracle.OracleModel(tflite_quant.patch_blob(blob, tflite_quant.derive(TFLITE)))


@pytest.fixture(scope="session")
def synth():
    from birdnet_stm32.audio import synth as s

    return s


@pytest.fixture(scope="session")
def pcm_batch(synth):
    """8 synthetic 22.05 kHz chunks incl. the edge cases of SURVEY 8(d) config 1."""
    pcm = synth.synth_pcm16(8, 66150, 22050, seed=1234, edge_cases=True)
    peak = synth.file_peaks(pcm)
    return pcm, peak
