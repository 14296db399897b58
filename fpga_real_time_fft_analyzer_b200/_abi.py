"""ctypes declarations of include/fra.h - one place, so the binding cannot drift
from the header.  `declare(cdll)` only sets argtypes/restype; it loads nothing."""
from __future__ import annotations

import ctypes as C

FRA_ABI_VERSION = 1

FRA_OK = 0
FRA_ERR_INVALID = -1
FRA_ERR_NO_DEVICE = -2
FRA_ERR_CUDA = -3
FRA_ERR_NOMEM = -4
FRA_ERR_UNSUPPORTED = -5
FRA_ERR_BUSY = -6

# command bytes (scripts/fft_analyzer_gui.py:28-37 of the reference)
UART_REQUEST_CMD = 0xA5
FPGA_RESET_CMD = 0xFF
ETHERNET_MODE_CMD = 0xEF
UART_MODE_CMD = 0xFE
START_COMMAND = 0x55
FILTER_UPDATE_CMD = 0xF1
FILTER_DEFAULT_CMD = 0x00
FILTER_CUSTOM_CMD = 0xA1
FILTER_NONE_CMD = 0xB1

FRA_WINDOW_LEN = 16384
FRA_SCALE_DEFAULT = 0x7FFFFFFF
FRA_ROUND_NEAREST = 0x1
FRA_K1_FORCE_LANE = 0x2
FRA_K1_FORCE_SPLIT = 0x4
FRA_K1_SPECULATE = 0x8
FRA_K1_FORCE_DUO = 0x20
FRA_PIPELINE = 0x40
FRA_K1_NO_BIASED = 0x80
FRA_K2_STAGED = 0x100
FRA_FFT_FIXED16 = 0x200
FRA_K2_64K_SPLIT = 0x400
FRA_HOST_HALF_SPECTRUM = 0x800
FRA_WINDOW_RTL_SKEW = 0x1000
FRA_K2_WIDE_CTA = 0x2000


class FraOutputs(C.Structure):
    """struct fra_outputs; every field is an address (0 = not requested)."""
    _fields_ = [("d_filtered", C.c_void_p), ("d_frames", C.c_void_p), ("d_iq", C.c_void_p),
                ("d_mag", C.c_void_p), ("d_phase", C.c_void_p)]


class FraStreamStats(C.Structure):
    """struct fra_stream_stats."""
    _fields_ = [("exact", C.c_int), ("n_chunks", C.c_int), ("chunk", C.c_int), ("warmup", C.c_int),
                ("n_mismatch", C.c_int), ("max_state_dev", C.c_int)]


# name -> (restype, argtypes); must list every function declared in include/fra.h
SIGNATURES = {
    "fra_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_uint]),
    "fra_destroy": (C.c_int, [C.c_void_p]),
    "fra_command": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "fra_load_bank1": (C.c_int, [C.c_void_p, C.POINTER(C.c_int8)]),
    "fra_load_sections": (C.c_int, [C.c_void_p, C.POINTER(C.c_int8)]),
    "fra_get_sections": (C.c_int, [C.c_void_p, C.POINTER(C.c_int8)]),
    "fra_set_mode": (C.c_int, [C.c_void_p, C.c_uint8]),
    "fra_set_mag_average": (C.c_int, [C.c_void_p, C.c_float]),
    "fra_reset": (C.c_int, [C.c_void_p]),
    "fra_get_mode": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint8)]),
    "fra_get_bank": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int8)]),
    "fra_get_transport": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint8)]),
    "fra_get_counters": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4),
    "fra_window_rom": (C.c_int, [C.POINTER(C.c_int16)]),
    "fra_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(FraOutputs), C.c_void_p]),
    "fra_process_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(FraOutputs)]),
    "fra_process_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(FraOutputs),
                                         C.POINTER(C.c_uint64)]),
    "fra_host_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "fra_set_host_half_share": (C.c_int, [C.c_void_p, C.c_double]),
    "fra_get_host_transfer": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fra_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fra_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fra_iir_stream": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                 C.POINTER(FraStreamStats)]),
    "fra_fft_only": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "fra_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "fra_profile_last": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fra_sync": (C.c_int, [C.c_void_p]),
    "fra_join": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fra_last_kernel_count": (C.c_int, [C.c_void_p]),
    "fra_last_cuda_error": (C.c_char_p, [C.c_void_p]),
    "fra_strerror": (C.c_char_p, [C.c_int]),
    "fra_abi_version": (C.c_int, []),
}


def declare(cdll):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)          # AttributeError if the library lacks an ABI symbol
        fn.restype = res
        fn.argtypes = args
    return cdll
