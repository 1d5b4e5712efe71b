"""The C-ABI library loads on a CPU-only box and exports every symbol include/trx2dyn.h
declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import trx2dyn  # noqa: F401
from trx2dyn import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "trx2dyn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = capi.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert lib.trx_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.TrxError, match="no CUDA device|CPU fallback"):
        capi.Context(0)


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both headers must compile as C99 (no C++-isms, no torch types), and every entry
    point must say which reference code it replaces."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "trx2dyn.h"\n#include "trx_centroid_model.h"\nint main(void) { return trx_abi_version() * 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(ROOT, "include", "trx2dyn.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                  # comments may mention torch streams
    assert "torch" not in code.lower() and "std::" not in code and "at::" not in code
    # the compute entry points cite the reference lines they replace
    for fn in ("trx_tables_create", "trx_energy_grad", "trx_fold_create", "trx_fold_run", "trx_glocon_matrix", "trx_tmscore_matrix"):
        head = text[:text.index("int " + fn + "(")]
        comment = head[head.rindex("/*"):]
        assert "Replaces" in comment and (".py:" in comment), fn
