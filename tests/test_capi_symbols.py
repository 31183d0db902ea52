"""The C-ABI library loads on a CPU box and exports every entry point that include/agcn_b200.h declares
(no compute calls here -- those need a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "agcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agcn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from fusion_gcn_b200 import build, capi
    build.build()
    names = declared_symbols()
    assert len(names) >= 15
    handle = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in agcn_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == names, "fusion_gcn_b200/capi.py must bind exactly the declared entry points"
    lib = capi.lib()
    assert lib.agcn_version() >= 100
    assert lib.agcn_bn_workspace_bytes(64) > 0
    assert lib.agcn_conv_wgrad_workspace_bytes(2, 8, 8, 25, 64, 64, 9) > 0
    assert lib.agcn_conv_fwd_workspace_bytes(64, 64, 9, capi.PREC_FP32) == 2 * 64 * 64 * 9 * 4
    assert lib.agcn_conv_fwd_workspace_bytes(64, 64, 9, capi.PREC_TF32) == 0


def test_argument_validation_without_gpu():
    """Shape / null checks run before any CUDA call, so the error paths are testable on a CPU box."""
    from fusion_gcn_b200 import capi
    lib = capi.lib()
    rc = lib.agcn_conv_fwd(None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, None, 0, None)
    assert rc == 6 and b"null" in lib.agcn_last_error_string()
    rc = lib.agcn_joint_mix(1, 1, 1, 1, 4, 25, 8, 24, 8, 7, 0, 0, None, 0, None)       # mix mode 7 does not exist
    assert rc == 2 and b"unknown mode" in lib.agcn_last_error_string()
    rc = lib.agcn_joint_gram(1, 1, 1, 1, 4, 40, 8, 8, 3, 0, 0, 0, 0, 8, 2, 0, None)    # V = 40 > 32 runs as ONE chunk
    assert rc == 2 and b"one chunk" in lib.agcn_last_error_string()
    rc = lib.agcn_conv_fwd(1, 1, None, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, None, 0, None)
    assert rc == 1
    with pytest.raises(RuntimeError, match="status 1"):
        capi.check(rc, "agcn_conv_fwd")
