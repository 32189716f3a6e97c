"""ctypes binding of libpgk.so (include/pgk.h).

There is no fallback: if the shared library is missing or the device is not
sm_100, importing the compute path raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C pggan-pytorch_b200/csrc``.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGK_LIB: another build of the same library (A/B runs, e.g. the programmatic-dependent-launch build libpgk_pdl.so)
LIB_PATH = os.environ.get('PGK_LIB') or os.path.join(_HERE, 'csrc', 'libpgk.so')

P, I, L, F = c_void_p, c_int, c_longlong, c_float

# name -> argument ctypes, in the order of include/pgk.h (the trailing stream argument is appended automatically)
SIGNATURES = {
    'pgk_prep_weight': [P, F, I, I, I, I, I, P, P],
    'pgk_unprep_grad': [P, F, I, I, I, I, I, P, I],
    'pgk_pack_operand': [P, I, I, P, L, I],
    'pgk_pack_thin': [P, I, I, P, L, I],
    'pgk_conv_thin': [P, I, I, L, I, I, I, I, I, P, L, P, I, P, L, F, P, L, P],
    'pgk_prep_multi': [P, I],
    'pgk_unprep_multi': [P, I],
    'pgk_conv': [P, I, I, L, I, I, I, I, I, I, I, P, P, L, P, P, P, I, P, L, F, P, L, P],
    'pgk_cvt_fp16x2': [P, L, I, L, P, L],
    'pgk_pack_operand_fp16': [P, I, I, P, L, I],
    'pgk_conv_fp16': [P, L, I, I, I, I, I, I, P, L, P, P, P, I, P, I, L],
    'pgk_wgrad': [P, L, P, L, I, I, I, I, I, I, I, I, I, I, P, P, P, P, ctypes.c_uint],
    'pgk_bias_grad': [P, L, I, I, I, I, I, P, F, P, I],
    'pgk_from_rgb': [P, I, I, I, I, I, P, F, P, I, P, L, P, I, L, P, L],
    'pgk_from_rgb_dgrad': [P, I, L, I, I, I, I, I, P, F, F, I, I, P],
    'pgk_to_rgb': [P, I, L, I, I, I, I, P, F, P, F, P, L, I, P, F, P, F, I, P, P, P],
    'pgk_to_rgb_dgrad': [P, I, I, I, I, I, P, F, F, I, P, I, L, P],
    'pgk_rgb_wgrad': [P, I, P, I, L, I, I, I, I, I, I, I, F, F, P, I, I, P, P, P],
    'pgk_pool2': [P, L, I, I, I, I, I, I, F, P, L, F, P, L, P, P, P, L],
    'pgk_mask_mul': [P, L, I, I, I, I, I, I, F, P, L, P, L, P, P, L],
    'pgk_axpby': [P, L, F, P, L, F, I, L, P, L],
    'pgk_pixelnorm': [P, L, I, L, I, P, L, P, P, L],
    'pgk_pixelnorm_bwd': [P, L, P, L, P, I, L, I, P, L],
    'pgk_latent_norm': [P, I, I, I, P, I, L],
    'pgk_stddev_stats': [P, L, I, I, L, P, P, I],
    'pgk_group_dot_pos': [P, L, I, I, I, I, I, P, P],
    'pgk_stddev_bwd': [P, L, P, P, I, I, L, P, L],
    'pgk_stddev_bwd2': [P, L, P, L, P, P, I, L, P, I, P, L, P],
    'pgk_posbias_wgrad': [P, L, I, I, I, I, I, P, F, I, I, P],
    'pgk_prep_posbias': [P, F, I, I, I, I, I, P],
    'pgk_linear_fwd': [P, L, I, I, I, P, P, P],
    'pgk_linear_bwd': [P, L, I, I, I, P, P, P, P, L, P, P],
    'pgk_colsum': [P, L, I, I, I, F, P],
    'pgk_interpolate': [P, P, P, I, L, P],
    'pgk_d_loss_seed': [P, I, F, P, P, P, P],
    'pgk_mean_scale': [P, I, F, P],
    'pgk_gp_penalty': [P, I, L, F, F, P, P, P, P, P, P],
    'pgk_fill': [P, L, F],
    'pgk_pool_img': [P, I, I, I, I, I, F, P],
    'pgk_unpool_img_add': [P, I, I, I, I, F, I, P],
    'pgk_real_prep': [P, I, I, I, I, I, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                      ctypes.c_double, P],
    'pgk_adam_multi': [P, I, L, F, F, F],
}



class PrepLayer(ctypes.Structure):
    """PgkPrepLayer of include/pgk.h (one conv layer of pgk_prep_multi's table)."""
    _fields_ = [('w', c_void_p), ('c', c_float), ('kind', c_int), ('cin', c_int), ('cin_stride', c_int),
                ('cout', c_int), ('ks', c_int), ('planes', c_int), ('wf', c_void_p), ('wb', c_void_p),
                ('F', c_void_p), ('F_ps', c_longlong), ('B', c_void_p), ('B_ps', c_longlong), ('F16', c_void_p),
                ('F16_ps', c_longlong), ('thinF', c_int), ('thinB', c_int)]


class UnprepLayer(ctypes.Structure):
    """PgkUnprepLayer of include/pgk.h."""
    _fields_ = [('dwp', c_void_p), ('c', c_float), ('kind', c_int), ('cin', c_int), ('cin_stride', c_int),
                ('cout', c_int), ('ks', c_int), ('dw', c_void_p), ('accumulate', c_int)]


_lib = None


class PgkError(RuntimeError):
    pass


def load():
    """Load libpgk.so once; raise loudly if it is absent (no CPU / PyTorch fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PgkError('libpgk.so not found at %s -- build it first (__graft_entry__.build() or '
                       '`make -C pggan-pytorch_b200/csrc`); there is no fallback path' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.pgk_last_error.restype = c_char_p
    lib.pgk_last_error.argtypes = []
    lib.pgk_version.restype = c_int
    lib.pgk_arch_check.argtypes = [c_int]
    lib.pgk_arch_check.restype = c_int
    lib.pgk_launch_count.restype = c_longlong
    lib.pgk_reset_launch_count.restype = None
    lib.pgk_count_launch.argtypes = [c_int]
    lib.pgk_count_launch.restype = None
    lib.pgk_pdl_state.argtypes = [c_int]
    lib.pgk_pdl_state.restype = c_int
    lib.pgk_prof_enable.argtypes = [c_int]
    lib.pgk_prof_enable.restype = None
    lib.pgk_prof_read.argtypes = [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_longlong)]
    lib.pgk_prof_read.restype = c_int
    lib.pgk_prof_read_products.argtypes = [c_int, ctypes.POINTER(ctypes.c_double)]
    lib.pgk_prof_read_products.restype = c_int
    lib.pgk_prof_reset.restype = None
    lib.pgk_pack_thin_plane_elems.argtypes = [c_int, c_int]
    lib.pgk_pack_thin_plane_elems.restype = c_longlong
    lib.pgk_set_tc.argtypes = [c_int]
    lib.pgk_set_tc.restype = None
    lib.pgk_conv_tc_supported.argtypes = [c_int] * 7
    lib.pgk_conv_tc_supported.restype = c_int
    lib.pgk_adam_chunk.argtypes = []
    lib.pgk_adam_chunk.restype = c_int
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = list(args) + [c_void_p]
        fn.restype = c_int
    _lib = lib
    return lib


def exported_symbols():
    return ['pgk_version', 'pgk_last_error', 'pgk_arch_check', 'pgk_launch_count', 'pgk_reset_launch_count',
            'pgk_count_launch', 'pgk_pdl_state',
            'pgk_prof_enable', 'pgk_prof_read', 'pgk_prof_read_products', 'pgk_prof_reset', 'pgk_set_tc', 'pgk_pack_thin_plane_elems',
            'pgk_conv_tc_supported', 'pgk_adam_chunk'] + list(SIGNATURES)


_checked_devices = set()


def check_device(device):
    lib = load()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _checked_devices:
        return
    rc = lib.pgk_arch_check(idx)
    if rc != 0:
        raise PgkError(lib.pgk_last_error().decode())
    _checked_devices.add(idx)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def call(name, *args):
    """Enqueue one libpgk call on torch's current CUDA stream."""
    lib = load()
    rc = getattr(lib, name)(*args, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise PgkError('%s failed (%d): %s' % (name, rc, lib.pgk_last_error().decode()))


def launch_count():
    return int(load().pgk_launch_count())


def add_launches(n):
    """Account for kernels launched by a replayed CUDA graph (they do not pass through the C entry points)."""
    load().pgk_count_launch(int(n))


def adam_chunk():
    """Elements one block of pgk_adam_multi updates (include/pgk.h: PGK_ADAM_CHUNK)."""
    return int(load().pgk_adam_chunk())


def pdl_state(on=None):
    """Programmatic dependent launch (include/pgk.h): query (None) or switch; returns the state."""
    return bool(load().pgk_pdl_state(-1 if on is None else int(bool(on))))


def reset_launch_count():
    load().pgk_reset_launch_count()


def prof_enable(on):
    load().pgk_prof_enable(1 if on else 0)


def prof_read(family):
    """(algorithmic FLOPs, algorithmic bytes, device ms, launches) of one kernel family since the last prof_reset()."""
    lib = load()
    f, b, t, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), c_longlong()
    rc = lib.pgk_prof_read(family, ctypes.byref(f), ctypes.byref(b), ctypes.byref(t), ctypes.byref(n))
    if rc != 0:
        raise PgkError('pgk_prof_read failed: %s' % lib.pgk_last_error().decode())
    return f.value, b.value, t.value, n.value


def prof_read_products(family):
    """bf16 tensor-core FLOPs issued by one kernel family since the last prof_reset() (algorithmic FLOPs x the
    1 / 3 / 6 plane products of each launch)."""
    lib = load()
    f = ctypes.c_double()
    rc = lib.pgk_prof_read_products(family, ctypes.byref(f))
    if rc != 0:
        raise PgkError('pgk_prof_read_products failed: %s' % lib.pgk_last_error().decode())
    return f.value


def prof_reset():
    load().pgk_prof_reset()
