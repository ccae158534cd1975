"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/vulcan_b200.h declares; compute entry points fail loudly (no CPU fallback)."""
import os
import re

import pytest

from helpers import REPO


def _header_symbols():
    src = open(os.path.join(REPO, "include", "vulcan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vk_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_list():
    from vulcan_b200 import _abi
    assert _header_symbols() == sorted(_abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from vulcan_b200 import _abi, build
    build.build()
    lib = _abi.load()
    for name in _header_symbols():
        assert hasattr(lib, name), name
    assert lib.vk_abi_version() == 3


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vulcan_b200 import _abi
    from helpers import load_network
    with pytest.raises(_abi.VulcanB200Error):
        _abi.DeviceNetwork(load_network("HD189"))


def test_product_does_not_import_oracle():
    """the product package must never route through oracle/ (it is test infrastructure)."""
    pkg = os.path.join(REPO, "vulcan_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "vk_oracle" not in text and "libvk_oracle" not in text, f
