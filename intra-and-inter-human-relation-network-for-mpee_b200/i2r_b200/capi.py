"""ctypes binding of include/i2r.h.  There is no fallback: a missing library is an error."""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# I2R_LIB: debugging aid (tools/hang_hunt.py loads instrumented builds of the same ABI); the product path never sets it
LIB_PATH = os.environ.get("I2R_LIB") or os.path.join(HERE, "libi2r_sm100.so")

I2R_MAX_TAPS = 9
I2R_MAX_GROUP = 6
I2R_MAX_CHAIN_PROBLEMS = 40
I2R_MAX_CHAIN_LAYERS = 16
E_UNSUPPORTED = -2
F_RELU = 1
F_OUT_NCHW_F32 = 2
F_OUT_F32 = 4
F_SPLIT = 8
F_OUT_T16 = 16
F_GELU = 32
F_ACT_FIRST = 64

EXPORTS = [
    "i2r_version", "i2r_last_error", "i2r_device_check", "i2r_sm_count", "i2r_conv_igemm", "i2r_conv_halo", "i2r_conv_halo_supported", "i2r_conv_halo_chain", "i2r_conv_halo_chain_workspace", "i2r_debug_trace", "i2r_debug_flags", "i2r_debug_chain_flags", "i2r_debug_hang_buffer", "i2r_hflip_f32", "i2r_flip_merge", "i2r_decode_heatmaps", "i2r_crop_persons", "i2r_box_masks", "i2r_mask_res_stem",
    "i2r_sizeof_conv_problem", "i2r_stem_conv3x3s2", "i2r_stem_conv3x3s2_tc", "i2r_stem_tc_weight_bytes", "i2r_maxpool3x3s2", "i2r_attention_varlen", "i2r_attention_workspace_bytes", "i2r_attention_tc", "i2r_encoder_tail", "i2r_encoder_tail_weight_bytes", "i2r_attention_tc_workspace_bytes", "i2r_layernorm", "i2r_add_f16", "i2r_upsum", "i2r_dwconv3x3", "i2r_upsum_bilinear", "i2r_layernorm_padded", "i2r_window_rows", "i2r_ln_window_gather", "i2r_window_scatter_add", "i2r_window_attention", "i2r_window_attention_tc",
]


class ConvProblem(ctypes.Structure):
    _fields_ = [
        ("x", ctypes.c_void_p), ("w", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("bias", ctypes.c_void_p),
        ("add0", ctypes.c_void_p), ("add1", ctypes.c_void_p), ("y", ctypes.c_void_p),
        ("NB", ctypes.c_int32), ("IH", ctypes.c_int32), ("IW", ctypes.c_int32),
        ("Cin", ctypes.c_int32), ("KC", ctypes.c_int32),
        ("in_pix_stride", ctypes.c_int32), ("in_shift", ctypes.c_int32),
        ("OH", ctypes.c_int32), ("OW", ctypes.c_int32), ("stride", ctypes.c_int32),
        ("Cout", ctypes.c_int32), ("Npad", ctypes.c_int32), ("out_pix_stride", ctypes.c_int32),
        ("OHf", ctypes.c_int32), ("OWf", ctypes.c_int32),
        ("out_mul", ctypes.c_int32), ("out_offy", ctypes.c_int32), ("out_offx", ctypes.c_int32),
        ("add0_shift", ctypes.c_int32), ("add1_shift", ctypes.c_int32), ("add_pix_stride", ctypes.c_int32),
        ("ntaps", ctypes.c_int32),
        ("dy", ctypes.c_int8 * (I2R_MAX_TAPS + 3)), ("dx", ctypes.c_int8 * (I2R_MAX_TAPS + 3)),
        ("flags", ctypes.c_uint32),
        ("w_folded", ctypes.c_void_p), ("w_folded_copies", ctypes.c_int32), ("pair_lo_offset", ctypes.c_int32),
    ]


class I2RError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def load():
    """Load libi2r_sm100.so (building is the job of build.py / __graft_entry__.build())."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise I2RError(
                "libi2r_sm100.so not found at %s -- run `python __graft_entry__.py build` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
        lib.i2r_version.restype = ctypes.c_int
        lib.i2r_last_error.restype = ctypes.c_char_p
        lib.i2r_device_check.argtypes = [i32]
        lib.i2r_sm_count.argtypes = [i32]
        lib.i2r_conv_igemm.argtypes = [ctypes.POINTER(ConvProblem), i32, i32, vp]
        lib.i2r_conv_halo.argtypes = [ctypes.POINTER(ConvProblem), i32, vp]
        lib.i2r_conv_halo_supported.argtypes = [ctypes.POINTER(ConvProblem)]
        if "i2r_conv_halo_chain" in EXPORTS:
            lib.i2r_conv_halo_chain.argtypes = [ctypes.POINTER(ConvProblem), ctypes.POINTER(ctypes.c_int), i32, vp,
                                                ctypes.c_size_t, vp]
            lib.i2r_conv_halo_chain_workspace.argtypes = [ctypes.POINTER(ConvProblem), i32]
            lib.i2r_conv_halo_chain_workspace.restype = ctypes.c_size_t
        lib.i2r_debug_trace.argtypes = [vp, i32, i32]
        lib.i2r_debug_flags.argtypes = [i32]
        if "i2r_debug_chain_flags" in EXPORTS:
            lib.i2r_debug_chain_flags.argtypes = [i32]
        if "i2r_debug_hang_buffer" in EXPORTS:      # tools/hang_hunt.py drops it to load older builds (I2R_LIB)
            lib.i2r_debug_hang_buffer.argtypes = [vp]
        lib.i2r_stem_conv3x3s2.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
        lib.i2r_stem_conv3x3s2_tc.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
        lib.i2r_stem_tc_weight_bytes.restype = i64
        lib.i2r_maxpool3x3s2.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
        lib.i2r_attention_varlen.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, i32, f32, vp, i64,
                                             i32, i32, i32, i32, i32, vp]
        lib.i2r_attention_workspace_bytes.argtypes = [i32, i32, i32, i32]
        lib.i2r_attention_workspace_bytes.restype = i64
        lib.i2r_attention_tc.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, i32, f32, vp, i64,
                                         i32, i32, vp]
        lib.i2r_attention_tc_workspace_bytes.argtypes = [i32, i32, i32, i32]
        lib.i2r_attention_tc_workspace_bytes.restype = i64
        lib.i2r_encoder_tail.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, vp, vp, i32, i32, i32, f32, i32, vp]
        lib.i2r_encoder_tail_weight_bytes.argtypes = [i32]
        lib.i2r_encoder_tail_weight_bytes.restype = i64
        lib.i2r_dwconv3x3.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
        lib.i2r_upsum_bilinear.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]
        lib.i2r_layernorm_padded.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, i32, vp]
        lib.i2r_window_rows.argtypes = [i32, i32, i32, i32]
        lib.i2r_window_rows.restype = i64
        lib.i2r_ln_window_gather.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, i32, vp]
        lib.i2r_window_scatter_add.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
        lib.i2r_window_attention.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, f32, i32, i32, i32,
                                             i32, i32, vp]
        if "i2r_window_attention_tc" in EXPORTS:
            lib.i2r_window_attention_tc.argtypes = lib.i2r_window_attention.argtypes
        lib.i2r_layernorm.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, f32, i32, vp]
        lib.i2r_add_f16.argtypes = [vp, vp, vp, i64, i32, vp]
        lib.i2r_upsum.argtypes = [vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]
        lib.i2r_hflip_f32.argtypes = [vp, vp, i64, i32, vp]
        lib.i2r_flip_merge.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
        lib.i2r_mask_res_stem.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
        lib.i2r_crop_persons.argtypes = [vp, i32, i32, vp, i32, i32, i32, vp, vp, vp, vp]
        lib.i2r_box_masks.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
        lib.i2r_decode_heatmaps.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp]
        for name in EXPORTS:
            getattr(lib, name)  # raises AttributeError if the symbol is missing
        if lib.i2r_sizeof_conv_problem() != ctypes.sizeof(ConvProblem):
            raise I2RError("i2r_conv_problem layout mismatch: C %d vs ctypes %d" % (
                lib.i2r_sizeof_conv_problem(), ctypes.sizeof(ConvProblem)))
        if lib.i2r_version() != 4:
            raise I2RError("libi2r_sm100.so ABI version %d, binding expects 4" % lib.i2r_version())
        _lib = lib
        return lib


def check(rc, what):
    if rc != 0:
        msg = load().i2r_last_error().decode("utf-8", "replace")
        raise I2RError("%s failed (rc=%d): %s" % (what, rc, msg))
