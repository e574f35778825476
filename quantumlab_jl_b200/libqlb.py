"""ctypes binding of libqlb200.so (include/qlb200.h) -- the same C ABI a Julia `ccall` binds.

There is no CPU fallback: if the shared library is missing, or no sm_100 device is visible, every compute entry
point raises.  This module performs no arithmetic of its own.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QLB200_LIB", os.path.join(_HERE, "libqlb200.so"))
NQCLASS = 21
NCLASS = 6
CLASS_NAMES = ("ss", "ps", "pp", "ds", "dp", "dd")
QCLASS_NAMES = tuple(f"({CLASS_NAMES[X]}|{CLASS_NAMES[Y]})" for X in range(NCLASS) for Y in range(X + 1))


class QlbError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [
        ("quartets_unique", C.c_int64),
        ("quartets_kept", C.c_int64),
        ("quartets_kept_class", C.c_int64 * NQCLASS),
        ("prim_quartets_class", C.c_int64 * NQCLASS),
        ("pairs_total", C.c_int64),
        ("pairs_significant", C.c_int64),
        ("ms_total", C.c_double),
        ("ms_eri", C.c_double),
        ("ms_class", C.c_double * NQCLASS),
        ("flops_algorithmic", C.c_double),
        ("kernel_launches", C.c_int32),
        ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {
            "quartets_unique": self.quartets_unique,
            "quartets_kept": self.quartets_kept,
            "quartets_kept_class": {QCLASS_NAMES[i]: self.quartets_kept_class[i] for i in range(NQCLASS) if self.quartets_kept_class[i]},
            "prim_quartets_class": {QCLASS_NAMES[i]: self.prim_quartets_class[i] for i in range(NQCLASS) if self.prim_quartets_class[i]},
            "pairs_total": self.pairs_total,
            "pairs_significant": self.pairs_significant,
            "ms_total": self.ms_total,
            "ms_eri": self.ms_eri,
            "ms_class": {QCLASS_NAMES[i]: self.ms_class[i] for i in range(NQCLASS) if self.ms_class[i]},
            "flops_algorithmic": self.flops_algorithmic,
            "kernel_launches": self.kernel_launches,
        }


_lib = None
_initialised = False
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

EXPORTS = (
    "qlb_init", "qlb_finalize", "qlb_last_error", "qlb_device_info", "qlb_basis_create", "qlb_basis_destroy",
    "qlb_basis_nbf", "qlb_basis_npairs", "qlb_eri_block", "qlb_eri_tensor", "qlb_schwarz", "qlb_build_jk", "qlb_fock",
    "qlb_build_jk_device", "qlb_symmetrize_device", "qlb_fock_device", "qlb_comm_unique_id", "qlb_comm_init",
    "qlb_comm_destroy", "qlb_allreduce_jk_device", "qlb_overlap", "qlb_kinetic", "qlb_nuclear_attraction",
    "qlb_set_profiling", "qlb_fp64_peak_probe", "qlb_ri_create", "qlb_ri_destroy", "qlb_ri_naux", "qlb_ri_b_tensor",
    "qlb_ri_eri_tensor", "qlb_ri_build_jk", "qlb_schedule_pieces", "qlb_qlog",
    "qlb_scf_create", "qlb_scf_destroy", "qlb_scf_set_density", "qlb_scf_build", "qlb_scf_step", "qlb_scf_get",
    "qlb_dgemm_device", "qlb_dsymm_packed_device", "qlb_debug_set_schedule_feedback", "qlb_tensor_contract", "qlb_ri_slice", "qlb_ri_build_jk_occ", "qlb_fp64_tensor_peak_probe",
    "qlb_multi_create", "qlb_multi_destroy", "qlb_multi_ngpu", "qlb_multi_basis0", "qlb_multi_build_jk", "qlb_multi_fock",
)


def load():
    """dlopen the library and declare the prototypes (no device needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise QlbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.qlb_last_error.restype = C.c_char_p
    lib.qlb_init.argtypes = [C.c_int]
    lib.qlb_device_info.argtypes = [C.c_char_p, C.c_int, _ip, _ip, _ip]
    lib.qlb_basis_create.argtypes = [C.c_int, _dp, _ip, _ip, _dp, _dp, C.POINTER(C.c_void_p)]
    lib.qlb_basis_destroy.argtypes = [C.c_void_p]
    lib.qlb_basis_nbf.argtypes = [C.c_void_p, _ip]
    lib.qlb_basis_npairs.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.qlb_eri_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    lib.qlb_eri_tensor.argtypes = [C.c_void_p, _dp]
    lib.qlb_schwarz.argtypes = [C.c_void_p, _dp]
    lib.qlb_build_jk.argtypes = [C.c_void_p, _dp, C.c_double, _dp, _dp, C.POINTER(Stats)]
    lib.qlb_fock.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp, C.POINTER(Stats)]
    lib.qlb_build_jk_device.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.POINTER(Stats)]
    lib.qlb_symmetrize_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qlb_fock_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qlb_comm_unique_id.argtypes = [C.c_void_p]
    lib.qlb_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.qlb_allreduce_jk_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qlb_overlap.argtypes = [C.c_void_p, _dp]
    lib.qlb_kinetic.argtypes = [C.c_void_p, _dp]
    lib.qlb_nuclear_attraction.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    lib.qlb_set_profiling.argtypes = [C.c_int]
    lib.qlb_qlog.argtypes = [C.c_double]
    lib.qlb_qlog.restype = C.c_int
    lib.qlb_schedule_pieces.argtypes = [_dp, C.c_int, _ip, _ip]
    lib.qlb_ri_create.argtypes = [C.c_void_p, C.c_int, _dp, _ip, _ip, _dp, _dp, C.POINTER(C.c_void_p)]
    lib.qlb_ri_destroy.argtypes = [C.c_void_p]
    lib.qlb_ri_naux.argtypes = [C.c_void_p, _ip]
    lib.qlb_ri_b_tensor.argtypes = [C.c_void_p, _dp]
    lib.qlb_ri_eri_tensor.argtypes = [C.c_void_p, _dp]
    lib.qlb_ri_build_jk.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.qlb_fp64_peak_probe.argtypes = [_dp]
    ab_library = os.environ.get("QLB200_LIB") is not None  # an A/B library may predate the newest entry points

    def proto(name, argtypes):
        if not hasattr(lib, name):
            if ab_library:
                return
            raise QlbError(f"{LIB_PATH} lacks {name}: rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
        getattr(lib, name).argtypes = argtypes

    proto("qlb_scf_create", [C.c_void_p, _dp, _dp, _dp, C.c_int, C.POINTER(C.c_void_p)])
    proto("qlb_scf_destroy", [C.c_void_p])
    proto("qlb_scf_set_density", [C.c_void_p, _dp])
    proto("qlb_scf_build", [C.c_void_p, C.c_double, C.c_int, _dp, C.POINTER(Stats)])
    proto("qlb_scf_step", [C.c_void_p])
    proto("qlb_scf_get", [C.c_void_p, C.c_int, _dp])
    proto("qlb_multi_create", [C.c_int, C.c_int, _dp, _ip, _ip, _dp, _dp, C.POINTER(C.c_void_p)])
    proto("qlb_multi_destroy", [C.c_void_p])
    proto("qlb_multi_ngpu", [C.c_void_p, _ip])
    proto("qlb_multi_basis0", [C.c_void_p, C.POINTER(C.c_void_p)])
    proto("qlb_multi_build_jk", [C.c_void_p, _dp, C.c_double, _dp, _dp, C.POINTER(Stats)])
    proto("qlb_multi_fock", [C.c_void_p, _dp, _dp, C.c_double, _dp, C.POINTER(Stats)])
    proto("qlb_fp64_tensor_peak_probe", [_dp])
    proto("qlb_ri_slice", [C.c_void_p, _ip, _ip])
    proto("qlb_ri_build_jk_occ", [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp])
    proto("qlb_tensor_contract", [C.c_int, _dp, _dp, C.c_int, _dp])
    proto("qlb_debug_set_schedule_feedback", [C.c_void_p, C.c_int, _dp, _dp])
    proto("qlb_dsymm_packed_device", [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])
    proto("qlb_dgemm_device", [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p])
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise QlbError(f"libqlb200 error {rc}: {load().qlb_last_error().decode()}")


def init(device: int = -1):
    """qlb_init: must succeed before any compute call; raises when there is no usable B200."""
    global _initialised
    lib = load()
    if not _initialised:
        check(lib.qlb_init(device))
        _initialised = True
    return lib


def finalize():
    global _initialised
    if _lib is not None and _initialised:
        _lib.qlb_finalize()
    _initialised = False


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(_ip)


def device_info():
    lib = init()
    name = C.create_string_buffer(128)
    sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib.qlb_device_info(name, 128, C.byref(sm), C.byref(ma), C.byref(mi)))
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (ma.value, mi.value)}


def fp64_tensor_peak_probe() -> float:
    lib = init()
    v = C.c_double()
    check(lib.qlb_fp64_tensor_peak_probe(C.byref(v)))
    return v.value


def fp64_peak_probe() -> float:
    lib = init()
    v = C.c_double()
    check(lib.qlb_fp64_peak_probe(C.byref(v)))
    return v.value
