"""ctypes binding of ``include/magat_gat.h`` (``lib/libmagat_gat.so``).

The library is the product: there is no Python or CPU fallback.  If it is missing, loading
raises with the build command; on a machine with a GPU every op of this package goes through
these entry points with raw device pointers and the caller's CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmagat_gat.so")
ABI_VERSION = 13

MODE_KEYQUERY, MODE_GAT_MODIFIED, MODE_GSO_VALUES = 0, 1, 2
DT_F32, DT_F64 = 0, 1
PATH_AUTO, PATH_SIMT, PATH_TCGEN05 = 0, 1, 2

EXPORTS = (
    "magat_abi_version", "magat_last_error", "magat_device_check", "magat_gso_scan",
    "magat_gso_build_ell", "magat_gat_wprep_floats", "magat_gat_forward", "magat_gat_forward_taps_valid",
    "magat_gat_forward_relu_bits_valid", "magat_gat_relu_bits_words",
    "magat_gat_actions_supported", "magat_gat_forward_actions", "magat_head_mean_forward", "magat_head_mean_backward",
    "magat_gso_self_loops",
    "magat_gat_bwd_partial_floats", "magat_gat_backward", "magat_gat_attention_dense",
    "magat_launch_count", "magat_profile_enable", "magat_profile_collect",
    "magat_gat_small_supported", "magat_gat_forward_small", "magat_gso_from_positions",
    "magat_gat_fused_supported", "magat_gat_fused_workspace_bytes", "magat_gat_forward_fused",
    "magat_gso_scan_nonzero", "magat_gso_edge_values", "magat_gso_pack_host", "magat_gso_from_rowbits",
    "magat_gat_backward_workspace_bytes", "magat_gat_backward_ws",
)

_i32, _i64, _ptr = C.c_int32, C.c_int64, C.c_void_p


class FwdArgs(C.Structure):
    _fields_ = [(n, _i32) for n in ("B", "N", "G", "F", "K", "P", "D", "mode", "concat", "relu", "path",
                                    "reserved")] + [
        ("x", _ptr), ("x_sb", _i64), ("x_sn", _i64),
        ("nbr_out", _ptr), ("nbr_in", _ptr), ("slot_in", _ptr), ("slot_out", _ptr),
        ("weight", _ptr), ("mixer", _ptr), ("weight_bias", _ptr), ("filterWeight", _ptr), ("bias", _ptr),
        ("y", _ptr), ("y_sb", _i64), ("y_sn", _i64), ("y_sc", _i64),
        ("att", _ptr), ("ain", _ptr), ("taps", _ptr), ("wprep", _ptr), ("sproj", _ptr), ("relu_bits", _ptr),
    ]


class BwdArgs(C.Structure):
    _fields_ = [(n, _i32) for n in ("B", "N", "G", "F", "K", "P", "D", "mode", "concat", "relu", "path",
                                    "need_dx", "need_dweight", "need_dfilter", "need_dbias",
                                    "need_dmixer", "taps_valid", "reserved")] + [
        ("x", _ptr), ("x_sb", _i64), ("x_sn", _i64),
        ("nbr_out", _ptr), ("nbr_in", _ptr), ("slot_in", _ptr),
        ("weight", _ptr), ("mixer", _ptr), ("weight_bias", _ptr), ("filterWeight", _ptr),
        ("y", _ptr), ("y_sb", _i64), ("y_sn", _i64), ("y_sc", _i64),
        ("att", _ptr), ("taps", _ptr), ("wprep", _ptr), ("sproj", _ptr),
        ("dy", _ptr), ("dy_sb", _i64), ("dy_sn", _i64), ("dy_sc", _i64),
        ("dx", _ptr), ("dweight", _ptr), ("dmixer", _ptr), ("dweight_bias", _ptr),
        ("dfilterWeight", _ptr), ("dbias", _ptr),
        ("gz", _ptr), ("datt", _ptr), ("rc", _ptr), ("partial", _ptr), ("relu_bits", _ptr),
    ]


class FusedArgs(C.Structure):
    _fields_ = [(n, _i32) for n in ("B", "N", "G", "F", "K", "P", "D", "mode", "concat", "relu", "s_dtype",
                                    "save", "team", "reserved")] + [
        ("S", _ptr), ("x", _ptr), ("x_sb", _i64), ("x_sn", _i64),
        ("weight", _ptr), ("mixer", _ptr), ("weight_bias", _ptr), ("filterWeight", _ptr), ("bias", _ptr),
        ("y", _ptr), ("y_sb", _i64), ("y_sn", _i64), ("y_sc", _i64),
        ("nbr_out", _ptr), ("nbr_in", _ptr), ("slot_in", _ptr), ("slot_out", _ptr),
        ("att", _ptr), ("taps", _ptr), ("sproj", _ptr), ("wprep", _ptr),
        ("workspace", _ptr), ("ws_bytes", C.c_size_t),
    ]


class MagatError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmagat_gat error {code}: {msg}")
        self.code = code


_lock = threading.Lock()
_lib = None


def lib():
    """Load (once per process) and return the shared library.  Never stored on modules, so
    ``GraphFilterBatchAttentional`` stays picklable for ``torch.multiprocessing.spawn``."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path.")
        L = C.CDLL(LIB_PATH)
        L.magat_abi_version.restype = C.c_int
        L.magat_last_error.restype = C.c_char_p
        L.magat_device_check.restype = C.c_int
        L.magat_gso_scan.argtypes = [_ptr, C.c_int, C.c_int, C.c_int, _ptr, _ptr, _ptr, _ptr]
        L.magat_gso_scan_nonzero.argtypes = [_ptr, C.c_int, C.c_int, C.c_int, _ptr, _ptr, _ptr, _ptr]
        L.magat_gso_scan_nonzero.restype = C.c_int
        L.magat_gso_edge_values.argtypes = [_ptr, C.c_int, _ptr, C.c_int, C.c_int, C.c_int, _ptr, _ptr]
        L.magat_gso_edge_values.restype = C.c_int
        L.magat_gso_pack_host.argtypes = [_ptr, C.c_int, C.c_long, C.c_int, _ptr, C.c_int]
        L.magat_gso_pack_host.restype = C.c_int
        L.magat_gso_from_rowbits.argtypes = [_ptr, C.c_int, C.c_int, _ptr, _ptr, _ptr]
        L.magat_gso_self_loops.argtypes = [_ptr, C.c_int, C.c_int, C.c_int, _ptr, _ptr, _ptr, _ptr]
        L.magat_gso_self_loops.restype = C.c_int
        L.magat_gso_from_rowbits.restype = C.c_int
        L.magat_gso_build_ell.argtypes = [_ptr, _ptr, C.c_int, C.c_int, C.c_int, _ptr, _ptr, _ptr, _ptr, _ptr]
        L.magat_gso_from_positions.argtypes = [_ptr, C.c_int, C.c_int, C.c_int, C.c_double, _ptr, _ptr, _ptr, _ptr]
        L.magat_gso_from_positions.restype = C.c_int
        L.magat_gat_wprep_floats.argtypes = [C.c_int] * 5
        L.magat_gat_wprep_floats.restype = C.c_size_t
        L.magat_gat_forward.argtypes = [C.POINTER(FwdArgs), _ptr]
        L.magat_gat_forward_taps_valid.argtypes = [C.POINTER(FwdArgs)]
        L.magat_gat_forward_taps_valid.restype = C.c_int
        L.magat_gat_forward_relu_bits_valid.argtypes = [C.POINTER(FwdArgs)]
        L.magat_gat_forward_relu_bits_valid.restype = C.c_int
        L.magat_gat_relu_bits_words.argtypes = [C.c_int] * 4
        L.magat_gat_relu_bits_words.restype = C.c_size_t
        L.magat_head_mean_forward.argtypes = [_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ptr, _ptr]
        L.magat_head_mean_forward.restype = C.c_int
        L.magat_head_mean_backward.argtypes = [_ptr, _ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ptr, _ptr]
        L.magat_head_mean_backward.restype = C.c_int
        L.magat_gat_actions_supported.argtypes = [C.POINTER(FwdArgs), C.c_int]
        L.magat_gat_actions_supported.restype = C.c_int
        L.magat_gat_forward_actions.argtypes = [C.POINTER(FwdArgs), _ptr, _ptr, C.c_int, _ptr, _ptr, _ptr, _ptr]
        L.magat_gat_forward_actions.restype = C.c_int
        L.magat_gat_bwd_partial_floats.argtypes = [C.c_int] * 7
        L.magat_gat_bwd_partial_floats.restype = C.c_size_t
        L.magat_gat_backward.argtypes = [C.POINTER(BwdArgs), _ptr]
        L.magat_gat_backward_workspace_bytes.argtypes = [C.c_int] * 8
        L.magat_gat_backward_workspace_bytes.restype = C.c_size_t
        L.magat_gat_backward_ws.argtypes = [C.POINTER(BwdArgs), _ptr, C.c_size_t, _ptr]
        L.magat_gat_backward_ws.restype = C.c_int
        L.magat_gat_attention_dense.argtypes = [_ptr, _ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ptr,
                                                _ptr]
        for name in ("magat_gso_scan", "magat_gso_build_ell", "magat_gat_forward", "magat_gat_backward",
                     "magat_gat_attention_dense"):
            getattr(L, name).restype = C.c_int
        L.magat_gat_small_supported.argtypes = [C.c_int] * 6
        L.magat_gat_small_supported.restype = C.c_int
        L.magat_gat_forward_small.argtypes = ([_ptr, C.c_int, _ptr, _i64, _i64] + [_ptr] * 5 + [_ptr, _i64, _i64, _i64, _ptr]
                                              + [C.c_int] * 9 + [_ptr])
        L.magat_gat_forward_small.restype = C.c_int
        L.magat_gat_fused_supported.argtypes = [C.c_int] * 8
        L.magat_gat_fused_supported.restype = C.c_int
        L.magat_gat_fused_workspace_bytes.argtypes = [C.c_int] * 8
        L.magat_gat_fused_workspace_bytes.restype = C.c_size_t
        L.magat_gat_forward_fused.argtypes = [C.POINTER(FusedArgs), _ptr]
        L.magat_gat_forward_fused.restype = C.c_int
        L.magat_launch_count.restype = C.c_long
        L.magat_profile_enable.argtypes = [C.c_int]
        L.magat_profile_enable.restype = None
        L.magat_profile_collect.argtypes = [C.c_char_p, C.c_long]
        L.magat_profile_collect.restype = C.c_long
        got = L.magat_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH}: ABI version {got}, expected {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise MagatError(code, lib().magat_last_error().decode(errors="replace"))


def profile_collect():
    """[(kernel name, launches, total ms)] recorded since magat_profile_enable(1)."""
    L = lib()
    buf = C.create_string_buffer(1 << 16)
    L.magat_profile_collect(buf, len(buf))
    out = []
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(",", 2)
        out.append((name, int(n), float(ms)))
    return out
