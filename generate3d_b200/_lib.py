"""ctypes binding of libg3d_b200.so (include/g3d.h). No fallback of any kind:
if the library is missing or has no CUDA device, the caller gets an exception."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libg3d_b200.so")

G3D_OK = 0
MODE_FAITHFUL, MODE_EXACT = 0, 1
MESH_BAD_INDEX, MESH_EMPTY, MESH_TOO_MANY_TRIS = 1, 2, 4


class TdfParams(C.Structure):
    _fields_ = [("ni", C.c_int32), ("nj", C.c_int32), ("nk", C.c_int32), ("origin", C.c_float * 3),
                ("dx", C.c_float), ("out_scale", C.c_float), ("trunc", C.c_float),
                ("occ_thresh", C.c_float), ("mode", C.c_int32), ("exact_band", C.c_int32)]


class TdfOutputs(C.Structure):
    _fields_ = [("sdf", C.c_void_p), ("tdf", C.c_void_p), ("tdflog", C.c_void_p),
                ("occ_bits", C.c_void_p), ("occ_f32", C.c_void_p), ("closest_tri", C.c_void_p),
                ("status", C.c_void_p)]


class G3DError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("g3d error %d: %s" % (code, text))
        self.code = code


# every symbol include/g3d.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
SYMBOLS = {
    "g3d_version": (C.c_int, []),
    "g3d_last_error": (C.c_char_p, []),
    "g3d_device_count": (C.c_int, []),
    "g3d_tdf_workspace_bytes": (_sz, [_i32, _i64, _i64, C.POINTER(TdfParams)]),
    "g3d_tdf_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, C.POINTER(TdfParams),
                                C.POINTER(TdfOutputs), _vp, _sz, _vp]),
    "g3d_tdf_from_sdf": (C.c_int, [_vp, _i64, _f32, _vp, _vp, _vp]),
    "g3d_unpack_occ": (C.c_int, [_vp, _i64, _vp, _vp]),
    "g3d_raycast_batch": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _i32, _vp, _vp,
                                    _vp, _vp]),
    "g3d_context_create": (C.c_int, [_i32, C.POINTER(_vp)]),
    "g3d_context_destroy": (None, [_vp]),
    "g3d_context_arena_bytes": (_sz, [_vp]),
    "g3d_tdf_batch_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, C.POINTER(TdfParams),
                                     C.POINTER(TdfOutputs)]),
    "g3d_tdf_batch_host_submit": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, C.POINTER(TdfParams),
                                            C.POINTER(TdfOutputs), C.POINTER(C.c_int32)]),
    "g3d_tdf_batch_host_wait": (C.c_int, [_vp, _i32]),
    "g3d_tdf_batch_host_submit_u16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, C.POINTER(TdfParams),
                                                C.POINTER(TdfOutputs), C.POINTER(C.c_int32)]),
    "g3d_raycast_batch_host": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _i32,
                                         _vp, _vp, _vp]),
    "g3d_dfgen_geometry": (C.c_int, [_i32, _i32, _i32, _vp, _vp, C.POINTER(TdfParams), _vp]),
    "g3d_write_df": (C.c_int, [C.c_char_p, _vp, _f32, _vp, _vp]),
    "g3d_read_df": (C.c_int, [C.c_char_p, _vp, _vp, _vp, _vp, _i64]),
    "g3d_occ_to_df_batch": (C.c_int, [_vp, _i32, _i32, _i32, _f32, _f32, _vp, _vp]),
    "g3d_raycast_color_batch": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _i32, _vp, _vp]),
    "g3d_raymarch": (C.c_int, [_vp, _vp, _i32, _i32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _i32,
                               _vp, _vp, _vp, _vp]),
    "g3d_zbuffer_index_map": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, C.c_float, C.c_float, C.c_float, _i32, _i32, _vp, _vp]),
    "g3d_profile_enable": (C.c_int, [_i32]),
    "g3d_profile_read": (C.c_int, [_vp, _vp]),
    "g3d_probe_fp32_tflops": (C.c_int, [_vp, _vp]),
    "g3d_host_alloc": (_vp, [_sz]),
    "g3d_host_free": (None, [_vp]),
    "g3d_vox2mesh": (C.c_int, [C.c_char_p, C.c_char_p, _i32, _i32, C.c_char_p, _f32, _i32]),
    "g3d_load_mesh": (C.c_int, [C.c_char_p, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp),
                                C.POINTER(_i64)]),
    "g3d_free": (None, [_vp]),
}

_lib = None


def lib():
    """The loaded library. Raises when it has not been built (python -m generate3d_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m generate3d_b200.build` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != G3D_OK:
        raise G3DError(rc, lib().g3d_last_error().decode(errors="replace"))
    return rc
