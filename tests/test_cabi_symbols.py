"""The C-ABI library loads on a CPU-only box and exports every symbol include/bmb200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

from bitmagic_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "bmb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bmb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libbmb200.so does not export {s}"
    assert sorted(capi.SYMBOLS) == syms, "capi.SYMBOLS out of sync with include/bmb200.h"


def test_error_messages_and_no_device_is_loud():
    lib = capi.lib()
    assert lib.bmb200_error_msg(0) == b"ok"
    assert b"no CPU fallback" in lib.bmb200_error_msg(capi.ERR_NODEVICE)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(capi.BMB200Error) as e:
            capi.Context(0)
        assert e.value.code == capi.ERR_NODEVICE


def test_bad_arguments_rejected_without_gpu():
    lib = capi.lib()
    assert lib.bmb200_init(0, None) == capi.ERR_BADARG
    assert lib.bmb200_set_free(None) == capi.ERR_BADARG
    assert lib.bmb200_rank_batch(None, None, C.c_uint64(0), None) == capi.ERR_RS_IDX_MISSING
