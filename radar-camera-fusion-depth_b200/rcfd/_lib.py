"""
ctypes binding of librcfd_b200.so (the C-ABI in include/rcfd.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  The library is built in-tree by build.py (nvcc, sm_100a).
"""
import ctypes
import os
from ctypes import c_int32, c_int64, c_float, c_void_p, c_char_p, POINTER, Structure

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_HERE, 'librcfd_b200.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_LEAKY, ACT_SIGMOID, ACT_DEPTH_HEAD = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TMA, ENGINE_STRIP = 0, 1, 2, 3, 4


class ConvDesc(Structure):
    _fields_ = [
        ('n', c_int32), ('ho', c_int32), ('wo', c_int32), ('cout', c_int32),
        ('kh', c_int32), ('kw', c_int32), ('stride', c_int32), ('pad', c_int32),
        ('in_dilation', c_int32), ('hin', c_int32), ('win', c_int32),
        ('src0', c_void_p), ('h0', c_int32), ('w0', c_int32), ('c0', c_int32),
        ('src1', c_void_p), ('c1', c_int32),
        ('weight', c_void_p), ('dst', c_void_p),
        ('scale', c_void_p), ('shift', c_void_p),
        ('act', c_int32), ('act_p0', c_float), ('act_p1', c_float),
        ('residual', c_void_p),
        ('stats_sum', c_void_p), ('stats_sqsum', c_void_p),
        ('accumulate', c_int32), ('dst_f32', c_int32), ('dtype', c_int32), ('engine', c_int32),
        ('weight_up2x', c_void_p),
    ]


class PackItem(Structure):
    """rcfd_pack_item (include/rcfd.h)."""
    _fields_ = [('src', c_void_p), ('dst', c_void_p), ('total', c_int64), ('kind', c_int32), ('dtype', c_int32),
                ('cout', c_int32), ('cin', c_int32), ('taps', c_int32), ('cin_off', c_int32), ('cin_cnt', c_int32),
                ('cpad', c_int32), ('col_off', c_int32), ('dst_cols', c_int32), ('block0', c_int32), ('nblocks', c_int32)]


# name -> argtypes (all return int32 status unless listed in _RESTYPE)
_P = c_void_p
_SIGS = {
    'rcfd_conv2d_fwd': [POINTER(ConvDesc), _P],
    'rcfd_conv2d_wgrad': [POINTER(ConvDesc), _P, _P, c_int64, _P],
    'rcfd_pack_conv_weight': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_pack_upconv2x_weight': [_P, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_unpack_conv_wgrad': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_pack_dgrad_s2_weight': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_pack_upconv2x_dgrad_weight': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_pack_batch': [_P, _P, c_int32, c_int32, _P],
    'rcfd_pack_item_blocks': [POINTER(PackItem)],
    'rcfd_bn_finalize': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int32, c_int64, c_float, c_float, _P],
    'rcfd_bn_fold': [_P, _P, _P, _P, _P, _P, c_int32, c_float, _P],
    'rcfd_bn_act_fwd': [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P],
    'rcfd_bn_train_act_fwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_float, c_float, c_int32, _P],
    'rcfd_bn_act_bwd_reduce': [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P],
    'rcfd_bn_act_bwd_reduce_acc': [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P],
    'rcfd_bn_act_bwd_apply': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P],
    'rcfd_bn_act_bwd_fused': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P],
    'rcfd_gate_fuse_fwd': [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P],
    'rcfd_gate_fuse_bwd': [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P],
    'rcfd_maxpool3x3s2_fwd': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_maxpool3x3s2_bwd': [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_maxpool3x3s2_fwd_idx': [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_maxpool3x3s2_bwd_idx': [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_upsample_nearest_bwd': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_leaky_bwd': [_P, _P, _P, c_int64, c_int32, _P],
    'rcfd_add_inplace': [_P, _P, c_int64, c_int32, _P],
    'rcfd_nchw_to_nhwc': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_nchw_to_s2d_nhwc': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_pack_stem_s2d_weight': [_P, _P, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_unpack_stem_s2d_wgrad': [_P, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_nhwc_to_nchw': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_depth_head_bwd': [_P, _P, _P, c_float, c_float, c_int64, c_int32, c_int32, _P],
    'rcfd_masked_l1_loss': [_P, _P, _P, c_float, _P, _P, _P, c_int64, _P],
    'rcfd_bilinear2x_fwd': [_P, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_bilinear2x_bwd': [_P, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_concat_logit': [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P],
    'rcfd_split_logit': [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P],
    'rcfd_smoothness_loss': [_P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P],
    'rcfd_sobel_smoothness_loss': [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P],
    'rcfd_outlier_removal': [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P],
    'rcfd_adam_step': [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_int32, _P],
    'rcfd_scatter_points_to_depth_map': [_P, _P, c_int32, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_scatter_tiles_argmax': [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P],
    'rcfd_stage1_to_stage2': [_P, _P, _P, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_roi_pool_fwd': [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_int32, _P],
    'rcfd_linear_leaky_fwd': [_P, _P, _P, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_split_bf16': [_P, _P, _P, _P, c_int64, _P],
    'rcfd_channel_stats': [_P, _P, _P, c_int64, c_int32, _P],
    'rcfd_epilogue_f32': [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_float, c_float, _P],
    'rcfd_transform_batch': [_P, _P, _P, _P, _P, c_int32, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, _P],
    'rcfd_roi_pool_bwd': [_P, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_int32, _P],
    'rcfd_cast_f32': [_P, _P, c_int64, c_int32, _P],
    'rcfd_linear_leaky_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, c_int32, c_int32, c_int32, _P],
    'rcfd_bce_logits_loss': [_P, _P, _P, c_float, _P, _P, _P, c_int64, _P],
    'rcfd_decode_crop': [_P, c_int32, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_int64, _P],
    'rcfd_encode_u16': [_P, _P, c_float, c_int64, _P],
    'rcfd_conv2d_wgrad_workspace': [POINTER(ConvDesc)],
    'rcfd_set_option': [c_char_p, c_int32],
    'rcfd_plan_row_chunks': [c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P],
    'rcfd_version': [], 'rcfd_arch': [], 'rcfd_last_error': [], 'rcfd_last_kernel': [],
}
_RESTYPE = {'rcfd_version': c_char_p, 'rcfd_arch': c_char_p, 'rcfd_last_error': c_char_p, 'rcfd_last_kernel': c_char_p,
            'rcfd_conv2d_wgrad_workspace': c_int64}

EXPORTED_SYMBOLS = sorted(_SIGS)

_lib = None
launch_count = 0      # number of C-ABI kernel-launching calls made (bench: gpu_launches evidence)
census = None         # list collecting (name, args, kernel) while rcfd.census.record() is active


class RcfdError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'librcfd_b200.so not found at %s -- build it with `python radar-camera-fusion-depth_b200/build.py` '
            '(nvcc, sm_100a).  There is no CPU / PyTorch fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, c_int32)
    _lib = lib
    # tuning knobs for experiments: RCFD_OPT="key=value,key=value" (rcfd_set_option)
    for kv in filter(None, os.environ.get('RCFD_OPT', '').split(',')):
        key, value = kv.split('=')
        if lib.rcfd_set_option(key.encode(), int(value)) != 0:
            raise RcfdError(lib.rcfd_last_error().decode())
    return lib


def call(name, *args):
    """Call an int-status entry point; raise RcfdError(message) on failure."""
    global launch_count
    lib = _lib if _lib is not None else load()
    rc = getattr(lib, name)(*args)
    launch_count += 1
    if census is not None:                # rcfd.census.record(): (entry point, arguments, kernel label) of every call
        conv = name in ('rcfd_conv2d_fwd', 'rcfd_conv2d_wgrad')
        census.append((name, args, lib.rcfd_last_kernel().decode() if conv else name))
    if rc != 0:
        raise RcfdError('%s failed (%d): %s' % (name, rc, lib.rcfd_last_error().decode()))
