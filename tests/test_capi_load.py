"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/gpsiq.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from pluto_gps_sim_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported_and_bound():
    hdr = open(os.path.join(REPO, "include", "gpsiq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gpsiq_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    raw = C.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "libgpsiq.so lacks " + name
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)


def test_desc_layout_is_64_bytes():
    assert capi.DESC_DTYPE.itemsize == 64
    assert capi.DESC_DTYPE.fields["code_phase0"][1] == 16 and capi.DESC_DTYPE.fields["flags"][1] == 56


def test_version_and_strerror():
    assert "sm_100a" in capi.version()
    assert capi.lib.gpsiq_strerror(capi.ERR_CUDA).decode() == "CUDA error"


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pluto_gps_sim_b200 import Synthesizer

    with pytest.raises(capi.GpsiqError) as ei:
        Synthesizer(max_chan=12, samples_per_epoch=1000, max_epochs=1)
    assert ei.value.status == capi.ERR_CUDA


def test_product_does_not_reference_oracle():
    pkg = os.path.join(REPO, "pluto_gps_sim_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".inc")):
                text = open(os.path.join(root, f)).read()
                for needle in ("liboracle", "oracle_lib", "gpsiq_oracle", "import oracle", "from oracle", "oracle/_ref",
                               "/oracle/", "oracle_synth"):
                    assert needle not in text, (needle, os.path.join(root, f))
