"""ctypes binding of libvnfr_b200.so (include/vnfr_b200.h).  There is NO fallback: if the library is missing or a call
fails, this raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvnfr_b200.so")

MAX_LEVELS = 24


class VnfrError(RuntimeError):
    pass


class Pyramid(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_levels", C.c_int32),
        ("scale", C.c_float * MAX_LEVELS), ("scale_d", C.c_double * MAX_LEVELS),
        ("lh", C.c_int32 * MAX_LEVELS), ("lw", C.c_int32 * MAX_LEVELS),
        ("oh", C.c_int32 * MAX_LEVELS), ("ow", C.c_int32 * MAX_LEVELS),
        ("level_off", C.c_int64 * (MAX_LEVELS + 1)), ("map_off", C.c_int64 * (MAX_LEVELS + 1)),
        ("tiles_x", C.c_int32 * MAX_LEVELS), ("tiles_y", C.c_int32 * MAX_LEVELS),
        ("tile_off", C.c_int32 * (MAX_LEVELS + 1)), ("px_off", C.c_int64 * (MAX_LEVELS + 1)),
    ]


class ConvOp(C.Structure):
    _fields_ = [
        ("tmap_w", C.c_ubyte * 128), ("tmap_a", C.c_ubyte * 128),
        ("inp", C.c_void_p), ("weights", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("out0", C.c_void_p), ("out1", C.c_void_p), ("out_f32", C.c_void_p),
        ("n_img", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("cin", C.c_int32), ("in_pitch", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad_h", C.c_int32), ("pad_w", C.c_int32),
        ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("cout", C.c_int32), ("cout_pad", C.c_int32), ("k_pad", C.c_int32), ("block_n", C.c_int32),
        ("n_split", C.c_int32), ("out0_pitch", C.c_int32), ("out1_pitch", C.c_int32), ("res_pitch", C.c_int32),
        ("out_f32_pitch", C.c_int32), ("relu", C.c_int32), ("dtype", C.c_int32), ("a_mode", C.c_int32), ("epi_mode", C.c_int32),
        ("tmap_c", C.c_ubyte * 128), ("tmap_r", C.c_ubyte * 128),
        ("prelu_alpha", C.c_void_p), ("n_img_dev", C.c_void_p), ("split3", C.c_int32), ("reserved", C.c_int32 * 1),
    ]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("conv", ConvOp), ("ext", C.c_void_p)]


class Block17Op(C.Structure):
    _fields_ = [
        ("tmap", (C.c_ubyte * 128) * 5), ("x", C.c_void_p),
        ("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p), ("w4", C.c_void_p),
        ("b1", C.c_void_p), ("b2", C.c_void_p), ("b3", C.c_void_p), ("b4", C.c_void_p),
        ("n_img", C.c_int32), ("dtype", C.c_int32),
    ]


class HeadsBack(C.Structure):
    """VnfrHeadsBack (include/vnfr_b200.h): split-precision weights + workspace of the R-/O-Net layers after the last conv."""
    _fields_ = [("w", C.c_void_p * 3), ("bias", C.c_void_p * 3), ("alpha", C.c_void_p * 2), ("planes", C.c_void_p)]


class TailLayer(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("N", C.c_int32), ("N_pad", C.c_int32), ("split_k", C.c_int32), ("rowop", C.c_int32),
        ("out_vec_pitch", C.c_int32),
        ("weights", C.c_void_p), ("bias", C.c_void_p), ("partial", C.c_void_p), ("a_in", C.c_void_p), ("a_next", C.c_void_p),
        ("out_vec", C.c_void_p),
    ]


class TailOp(C.Structure):
    _fields_ = [
        ("tmap_a", (C.c_ubyte * 128) * 3), ("tmap_w", (C.c_ubyte * 128) * 3),
        ("n_layers", C.c_int32), ("n_pad", C.c_int32), ("in_mode", C.c_int32),
        ("hw", C.c_int32), ("x_pitch", C.c_int32), ("x_dtype", C.c_int32), ("x_f32_pitch", C.c_int32),
        ("x_f32_cols", C.c_int32), ("emb_half_dtype", C.c_int32), ("reserved", C.c_int32),
        ("x", C.c_void_p), ("x_f32", C.c_void_p),
        ("layer", TailLayer * 3),
        ("emb_half", C.c_void_p), ("label", C.c_void_p), ("prob", C.c_void_p), ("label_f", C.c_void_p), ("prob_f", C.c_void_p),
        ("lp_pitch", C.c_int32), ("n_classes", C.c_int32),
        ("thr_class", C.c_void_p), ("thr", C.c_float), ("count_value", C.c_int32),
        ("count_cell", C.c_void_p), ("grid_barrier", C.c_void_p),
    ]


_lib = None

# name -> argtypes (restype is always int unless listed in _SPECIAL)
_P, _I, _F, _D, _LL = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong
_SIGS = {
    "vnfr_pyramid_plan": [_I, _I, _I, _I, _D, C.POINTER(Pyramid)],
    "vnfr_pyramid_resize_norm": [C.POINTER(Pyramid), _P, _P, _P],
    "vnfr_pnet_pack_weights": [_P, _I, _P, _I],
    "vnfr_pnet_sweep_compact": [C.POINTER(Pyramid), _P, _P, _F, _I, _P, _P, _P, _P, _P, _P, _P],
    "vnfr_nms_segments": [_I, _I, _P, _P, _P, _F, _I, _P, _P, _P],
    "vnfr_stage1_boxes": [C.POINTER(Pyramid), _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P],
    "vnfr_rnet_forward": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P],
    "vnfr_rnet_forward_tc": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, C.POINTER(HeadsBack), _P],
    "vnfr_onet_forward": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P],
    "vnfr_onet_forward_tc": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P,
                             C.POINTER(HeadsBack), _P],
    "vnfr_stage2_boxes": [_I, _I, _I, _I, _P, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P],
    "vnfr_stage3_faces": [_I, _I, _P, _P, _P, _P, _P, _F, _I, _I, _P, _P, _P, _P, _P],
    "vnfr_face_crops": [_P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P, _I, _P],
    "vnfr_conv_prepare": [C.POINTER(ConvOp)],
    "vnfr_conv_run": [C.POINTER(ConvOp), _P],
    "vnfr_run_ops": [C.POINTER(Op), _I, _P],
    "vnfr_count_launches": [_LL],
    "vnfr_maxpool3s2_nhwc": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "vnfr_avgpool_nhwc": [_P, _I, _I, _I, _I, _P, _I, _P],
    "vnfr_nchw3_to_nhwc8": [_P, _I, _I, _I, _P, _I, _P],
    "vnfr_nchw3_to_s2d16": [_P, _I, _I, _I, _P, _I, _P],
    "vnfr_u8hwc_to_s2d16": [_P, _I, _I, _I, _P, _I, _P],
    "vnfr_l2norm_rows": [_P, _I, _I, _I, _P, _P, _I, _P],
    "vnfr_logsoftmax_argmax": [_P, _I, _I, _I, _P, _P, _P, _P],
    "vnfr_topk_rows": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "vnfr_swap_rb_u8": [_P, _P, _LL, _P],
    "vnfr_nv12_to_rgb_u8": [_P, _P, _I, _I, _I, _P],
    "vnfr_gallery_topk": [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "vnfr_block17_prepare": [C.POINTER(Block17Op)],
    "vnfr_block17_run": [C.POINTER(Block17Op), _P],
    "vnfr_tail_prepare": [C.POINTER(TailOp)],
    "vnfr_tail_run": [C.POINTER(TailOp), _I, _P],
}


def lib():
    """Loads the shared library (once).  Raises if it has not been built -- there is no CPU / PyTorch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VnfrError("%s not found: run `python -m vn_celeb_face_recognition_b200.build` (needs nvcc); there is "
                            "no CPU fallback for this package" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.vnfr_last_error.restype = C.c_char_p
        l.vnfr_launch_count.restype = C.c_longlong
        l.vnfr_rnet_weight_floats.restype = C.c_int
        l.vnfr_onet_weight_floats.restype = C.c_int
        l.vnfr_pnet_packed_bytes.restype = C.c_int
        l.vnfr_heads_back_workspace_bytes.restype = C.c_longlong
        l.vnfr_heads_back_workspace_bytes.argtypes = [C.c_int, C.c_int]
        for name, args in _SIGS.items():
            fn = getattr(l, name)      # AttributeError if the symbol is missing: fail loudly
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = l
    return _lib


def exported_symbols():
    return ["vnfr_last_error", "vnfr_version", "vnfr_launch_count", "vnfr_rnet_weight_floats",
            "vnfr_onet_weight_floats", "vnfr_pnet_packed_bytes", "vnfr_heads_back_workspace_bytes"] + sorted(_SIGS)


def check(rc):
    if rc != 0:
        raise VnfrError("libvnfr_b200 call failed (%d): %s" % (rc, lib().vnfr_last_error().decode()))


def call(name, *args):
    check(getattr(lib(), name)(*args))


def launch_count():
    return int(lib().vnfr_launch_count())


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on ``device`` (default: the current device).  The C library launches on the
    CURRENT device: callers that hold tensors of another device wrap their calls in ``torch.cuda.device(tensor.device)``."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def pack_pnet_weights(packed_host, device):
    """_pack_pnet floats (host) -> the P-Net kernel's weight block on ``device`` (caller-owned; see vnfr_pnet_pack_weights)."""
    import torch
    n = lib().vnfr_pnet_packed_bytes()
    out = torch.zeros(n, dtype=torch.uint8)
    call("vnfr_pnet_pack_weights", C.c_void_p(packed_host.data_ptr()), packed_host.numel(), C.c_void_p(out.data_ptr()), n)
    return out.to(device)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return C.c_void_p(0 if t is None else t.data_ptr())
