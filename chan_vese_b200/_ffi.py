"""ctypes binding of include/chan_vese_b200.h -- the C ABI is the product boundary; this file only declares it."""
import ctypes as C
import os

from . import build as _build

u8p = C.POINTER(C.c_uint8)
u8pp = C.POINTER(u8p)
f64p = C.POINTER(C.c_double)
intp = C.POINTER(C.c_int)
vp = C.c_void_p

OK, ERR_INVALID_ARGUMENT, ERR_NO_DEVICE, ERR_CUDA, ERR_OUT_OF_MEMORY, ERR_STATE, ERR_COMM, ERR_CALLBACK = range(8)
STATUS_NAMES = ["CVB_OK", "CVB_ERR_INVALID_ARGUMENT", "CVB_ERR_NO_DEVICE", "CVB_ERR_CUDA", "CVB_ERR_OUT_OF_MEMORY",
                "CVB_ERR_STATE", "CVB_ERR_COMM", "CVB_ERR_CALLBACK"]
PRECISION_F64, PRECISION_F32 = 0, 1
MATH_FAST, MATH_STRICT = 0, 1
COMM_ID_BYTES = 128


class CsvParams(C.Structure):
    _fields_ = [("mu", C.c_double), ("nu", C.c_double), ("dt", C.c_double), ("eps", C.c_double),
                ("lambda1", C.c_double * 3), ("lambda2", C.c_double * 3)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("csv_step_launches", C.c_uint64), ("pm_step_launches", C.c_uint64),
                ("csv_ms", C.c_double), ("pm_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("peer_wait_ms", C.c_double), ("peer_waits", C.c_uint64)]


FRAME_FN = C.CFUNCTYPE(C.c_int, f64p, C.c_int, C.c_int, C.c_int, vp)
MASK_FN = C.CFUNCTYPE(C.c_int, u8p, C.c_int, C.c_int, C.c_int, vp)
MASK_SEPARATE, MASK_CONTOUR = 0, 1
CsvParamsP = C.POINTER(CsvParams)

# name -> (restype, argtypes); every symbol include/chan_vese_b200.h declares
SIGNATURES = {
    "cvb_version": (C.c_char_p, []),
    "cvb_device_count": (C.c_int, []),
    "cvb_context_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "cvb_context_destroy": (None, [vp]),
    "cvb_last_error": (C.c_char_p, [vp]),
    "cvb_context_set_math_mode": (C.c_int, [vp, C.c_int]),
    "cvb_context_set_tile_rows": (C.c_int, [vp, C.c_int]),
    "cvb_context_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
    "cvb_context_reset_stats": (C.c_int, [vp]),
    "cvb_context_synchronize": (C.c_int, [vp]),
    "cvb_context_trim": (C.c_int, [vp]),
    "cvb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "cvb_host_free": (None, [vp]),
    "cvb_pm_num_steps": (C.c_int, [C.c_double, C.c_double]),
    "cvb_levelset_checkerboard": (C.c_int, [C.c_int, C.c_int, f64p]),
    "cvb_levelset_rect": (C.c_int, [C.c_int] * 6 + [f64p]),
    "cvb_levelset_circ": (C.c_int, [C.c_int] * 5 + [f64p]),
    "cvb_auto_tile_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "cvb_slab_partition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, intp, intp]),
    "cvb_perona_malik": (C.c_int, [vp, u8pp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, u8pp, intp]),
    "cvb_csv_run": (C.c_int, [vp, u8pp, C.c_int, C.c_int, C.c_int, f64p, CsvParamsP, C.c_double, C.c_int, intp, f64p,
                              FRAME_FN, vp]),
    "cvb_csv_run_masks": (C.c_int, [vp, u8pp, C.c_int, C.c_int, C.c_int, f64p, CsvParamsP, C.c_double, C.c_int, intp, f64p,
                                    C.c_int, MASK_FN, vp]),
    "cvb_segment": (C.c_int, [vp, u8pp, C.c_int, C.c_int, C.c_int, f64p, C.c_int, C.c_double, C.c_double, C.c_double,
                              u8pp, CsvParamsP, C.c_double, C.c_int, intp, f64p, C.c_int, u8p]),
    "cvb_region_means": (C.c_int, [vp, u8pp, C.c_int, C.c_int, C.c_int, f64p, C.c_double, f64p, f64p]),
    "cvb_curvature": (C.c_int, [vp, f64p, C.c_int, C.c_int, f64p]),
    "cvb_delta_map": (C.c_int, [vp, f64p, C.c_size_t, C.c_double]),
    "cvb_stop_condition": (C.c_int, [vp, u8pp, C.c_int, C.c_int, C.c_int, C.c_double, f64p]),
    "cvb_mask": (C.c_int, [vp, f64p, C.c_int, C.c_int, C.c_int, u8p]),
    "cvb_session_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "cvb_session_create_slab": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "cvb_session_destroy": (None, [vp]),
    "cvb_session_upload_image": (C.c_int, [vp, u8pp]),
    "cvb_session_upload_levelset": (C.c_int, [vp, f64p]),
    "cvb_session_init_checkerboard": (C.c_int, [vp]),
    "cvb_session_perona_malik": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, intp]),
    "cvb_session_csv_run": (C.c_int, [vp, CsvParamsP, C.c_double, C.c_int, intp, f64p, FRAME_FN, vp]),
    "cvb_session_csv_run_masks": (C.c_int, [vp, CsvParamsP, C.c_double, C.c_int, intp, f64p, C.c_int, MASK_FN, vp]),
    "cvb_session_csv_step": (C.c_int, [vp, CsvParamsP, f64p, f64p, f64p]),
    "cvb_session_region_means": (C.c_int, [vp, C.c_double, f64p, f64p]),
    "cvb_session_download_levelset": (C.c_int, [vp, f64p]),
    "cvb_session_download_image": (C.c_int, [vp, u8pp]),
    "cvb_session_download_pm_state": (C.c_int, [vp, C.POINTER(f64p)]),
    "cvb_session_mask": (C.c_int, [vp, C.c_int, u8p]),
    "cvb_session_mask_packed": (C.c_int, [vp, C.c_int, u8p]),
    "cvb_session_upload_image_smooth": (C.c_int, [vp, u8pp, C.c_double, C.c_double, C.c_double, intp]),
    "cvb_session_save_image": (C.c_int, [vp]),
    "cvb_session_restore_image": (C.c_int, [vp]),
    "cvb_session_prefetch_image": (C.c_int, [vp, u8pp]),
    "cvb_session_release_scratch": (C.c_int, [vp]),
    "cvb_comm_create_id": (C.c_int, [vp, vp]),
    "cvb_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "cvb_comm_destroy": (C.c_int, [vp]),
    "cvb_batch_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "cvb_batch_destroy": (None, [vp]),
    "cvb_batch_upload_images": (C.c_int, [vp, u8pp]),
    "cvb_batch_upload_levelset": (C.c_int, [vp, f64p]),
    "cvb_batch_init_checkerboard": (C.c_int, [vp]),
    "cvb_batch_perona_malik": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, intp]),
    "cvb_batch_csv_run": (C.c_int, [vp, CsvParamsP, C.c_double, C.c_int, intp, f64p]),
    "cvb_batch_download_levelset": (C.c_int, [vp, C.c_int, f64p]),
    "cvb_batch_download_image": (C.c_int, [vp, C.c_int, u8pp]),
    "cvb_batch_mask": (C.c_int, [vp, C.c_int, C.c_int, u8p]),
    "cvb_batch_mask_packed": (C.c_int, [vp, C.c_int, C.c_int, u8p]),
    "cvb_batch_masks_packed": (C.c_int, [vp, C.c_int, u8p]),
    "cvb_batch_upload_images_smooth": (C.c_int, [vp, u8pp, C.c_double, C.c_double, C.c_double, intp]),
    "cvb_batch_save_images": (C.c_int, [vp]),
    "cvb_batch_restore_images": (C.c_int, [vp]),
    "cvb_batch_prefetch_images": (C.c_int, [vp, u8pp]),
    "cvb_batch_release_scratch": (C.c_int, [vp]),
}

_lib = None


def lib():
    """Load (building if necessary) the CUDA library.  There is no CPU fallback: failure to build or load raises."""
    global _lib
    if _lib is None:
        path = os.environ.get("CVB_LIB") or _build.build()
        handle = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib
