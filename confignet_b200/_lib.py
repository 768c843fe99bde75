"""ctypes binding of libconfignet_b200.so (the C ABI declared in include/confignet_b200.h).

The product path has no fallback: if the shared library is missing, importing this module raises;
if a call fails, ``CnError`` carries the library's message.  Build with ``__graft_entry__.build()``
(nvcc -gencode arch=compute_100a,code=sm_100a).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# The product library carries no test hooks.  CN_TEST_HOOKS=1 loads the hooks build of the same sources (cn_debug_* entry
# points, in-kernel role timers) for the measurement scripts; CN_LIB names an experimental build (A/B runs).
HOOKS_LIB_PATH = os.path.join(_HERE, "lib", "libconfignet_b200_hooks.so")
LIB_PATH = os.environ.get("CN_LIB") or (HOOKS_LIB_PATH if os.environ.get("CN_TEST_HOOKS") == "1" else
                                        os.path.join(_HERE, "lib", "libconfignet_b200.so"))
CSRC_DIR = os.path.join(_HERE, "csrc")


class CnError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """cn_conv_desc (include/confignet_b200.h)."""
    _fields_ = [("nd", ctypes.c_int), ("batch", ctypes.c_int), ("in_dims", ctypes.c_int * 3),
                ("cin", ctypes.c_int), ("cout", ctypes.c_int), ("ksize", ctypes.c_int * 3),
                ("stride", ctypes.c_int), ("upsample", ctypes.c_int), ("pad", ctypes.c_int)]

    def key(self):
        k = (self.nd, self.batch, tuple(self.in_dims), self.cin, self.cout, tuple(self.ksize),
             self.stride, self.upsample)
        return k if self.pad < 0 else k + (self.pad,)


def make_conv_desc(nd, batch, in_dims, cin, cout, ksize, stride=1, upsample=1, pad=-1):
    """pad = -1: TF "SAME"; pad >= 0: ZeroPadding(pad) + "VALID" (ResNet50 stem)."""
    in_dims = list(in_dims) + [1] * (3 - len(in_dims))
    ksize = list(ksize) + [1] * (3 - len(ksize))
    return ConvDesc(nd, batch, (ctypes.c_int * 3)(*in_dims), cin, cout, (ctypes.c_int * 3)(*ksize),
                    stride, upsample, pad)


ACT_NONE, ACT_LRELU, ACT_RELU, ACT_TANH, ACT_RELU6, ACT_SIGMOID = 0, 1, 2, 3, 4, 5
POOL_MAX, POOL_AVG_VALID = 0, 1
IMPL_AUTO, IMPL_FFMA, IMPL_TC = 0, 1, 2

_F = ctypes.POINTER(ctypes.c_float)
_V = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_f = ctypes.c_float
_D = ctypes.POINTER(ConvDesc)

# name -> argtypes; every function returns int (0 = ok).  Must list every symbol of the header.
SIGNATURES = {
    "cn_conv_out_dims": [_D, ctypes.POINTER(ctypes.c_int)],
    "cn_conv_fwd": [_D, _V, _V, _V, _I, _f, _V, _I, _V],
    "cn_conv_dgrad": [_D, _V, _V, _V, _I, _V],
    "cn_conv_wgrad": [_D, _V, _V, _V, _V, _I, _V],
    "cn_chan_sums": [_V, _V, _V, _I, _I, _I, _I, _f, _V, _V],
    "cn_chan_sums_dual": [_V, _I, _I, _I, _f, _V, _V, _V],
    "cn_chan_affine": [_V, _V, _V, _V, _I, _I, _I, _I, _f, _V, _V],
    "cn_chan_affine2": [_V, _V, _V, _V, _V, _I, _I, _I, _I, _f, _V, _V],
    "cn_chan_affine_pair": [_V, _V, _V, _V, _V, _I, _I, _I, _f, _V, _V, _V],
    "cn_chan_sums_splits": [_I, _I, _I],
    "cn_register_params": [_V, ctypes.c_size_t],
    "cn_unregister_params": [_V],
    "cn_weights_changed": [],
    "cn_params_changed": [_V],
    "cn_set_params_frozen": [_V, _I],
    "cn_graphs_captured": [],
    "cn_norm_coef": [_I, _V, _I, _V, _V, _I, _I, _I, _f, _V, _V, _V, _V, _V],
    "cn_lrelu_fwd": [_V, _f, _V, _L, _V],
    "cn_act_bwd": [_V, _V, _I, _f, _V, _L, _V],
    "cn_act_bwd_sqdiff": [_V, _V, _V, _V, _f, _I, _f, _V, _L, _V],
    "cn_axpby": [_V, _V, _f, _f, _V, _L, _V],
    "cn_maxpool2_fwd": [_V, _I, _I, _I, _I, _V, _V],
    "cn_maxpool2_bwd": [_V, _V, _V, _I, _I, _I, _I, _V, _V],
    "cn_rotate3d_fwd": [_V, _V, _I, _I, _I, _V, _V],
    "cn_rotate3d_bwd_grid": [_V, _V, _I, _I, _I, _V, _V],
    "cn_reduce": [_V, _V, _V, _I, _L, _I, _f, _f, _V, _V, _V],
    "cn_reduce_bwd": [_V, _V, _V, _I, _L, _I, _f, _f, _V, _V, _V],
    "cn_to_uint8": [_V, _V, _L, _V],
    "cn_from_uint8": [_V, _V, _L, _V],
    "cn_vgg_preprocess": [_V, _V, _L, _I, _V],
    "cn_adam_ema_step": [_V, _V, _V, _V, _V, _L, _f, _f, _f, _f, _f, _f, _V],
    "cn_adam_ema_step_dev": [_V, _V, _V, _V, _V, _L, _V, _f, _f, _f, _f, _f, _V],
    "cn_ema": [_V, _V, _L, _f, _V],
    "cn_multi_copy": [_I, _V, _V, _V, _V, _V],
    "cn_bn_fold": [_V, _V, _V, _V, _f, _I, _V, _V, _V],
    "cn_bn_act_fwd": [_V, _V, _V, _V, _I, _V, _L, _I, _V],
    "cn_bn_act_bwd": [_V, _V, _V, _V, _V, _V, _f, _I, _V, _V, _V, _V, _L, _I, _V],
    "cn_maxpool3s2_fwd": [_V, _I, _I, _I, _I, _V, _V],
    "cn_maxpool3s2_bwd": [_V, _V, _V, _I, _I, _I, _I, _V, _V],
    "cn_avgpool_fwd": [_V, _I, _I, _I, _V, _V],
    "cn_avgpool_bwd": [_V, _I, _I, _I, _V, _V],
    "cn_col_scale": [_V, _V, _I, _I, _V, _V],
    "cn_euler_fwd": [_V, _I, _V, _V],
    "cn_euler_bwd": [_V, _V, _I, _V, _V],
    "cn_rotate3d_bwd_rot": [_V, _V, _V, _I, _I, _I, _V, _V],
    "cn_norm_latent_loss_fwd": [_V, _V, _I, _I, _I, _f, _V, _V],
    "cn_norm_latent_loss_bwd": [_V, _V, _I, _I, _I, _f, _V, _V, _V, _V],
    "cn_pool2d_fwd": [_V, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _V, _I, _V],
    "cn_dwconv3x3_fwd": [_V, _V, _V, _I, _I, _I, _I, _I, _I, _f, _V, _V],
    "cn_resize_bilinear": [_V, _I, _I, _I, _I, _I, _I, _I, _V, _V],
    "cn_u8_to_f32": [_V, _V, _L, _V],
    "cn_pixel_map": [_V, _V, _L, _I, _V],
    "cn_act_ext": [_V, _V, _L, _I, _V],
}
NO_STATUS = {"cn_last_error": ctypes.c_char_p, "cn_version": ctypes.c_int, "cn_reduce_ws_floats": ctypes.c_int,
             "cn_launch_count": ctypes.c_longlong, "cn_params_epoch": ctypes.c_longlong, "cn_launch_count_add": ctypes.c_longlong, "cn_last_conv_impl": ctypes.c_int}

_lib = None


def load():
    """Loads the shared library (once) and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CnError("libconfignet_b200.so not built (%s); run __graft_entry__.build()" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        if not hasattr(lib, name) and os.environ.get("CN_ALLOW_PARTIAL") == "1":
            continue        # bring-up only: lets a partially built library be probed
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    if hasattr(lib, "cn_launch_count_add"):
        lib.cn_launch_count_add.argtypes = [ctypes.c_longlong]
    lib.cn_launch_count.argtypes = [ctypes.c_int]
    lib.cn_params_epoch.argtypes = [ctypes.c_void_p]
    for name, restype in NO_STATUS.items():
        if not hasattr(lib, name) and os.environ.get("CN_ALLOW_PARTIAL") == "1":
            continue
        getattr(lib, name).restype = restype
    # measurement knobs of the hooks build (ignored - loudly - by the product library, which has no cn_debug_* symbols)
    knobs = [("CN_CHUNK_KB", "cn_debug_set_chunk"), ("CN_PERSISTENT", "cn_debug_set_persistent"), ("CN_WCACHE", "cn_debug_set_wcache"),
             ("CN_COAL", "cn_debug_set_coal"), ("CN_DBG", "cn_debug_set"), ("CN_CLUSTER", "cn_debug_set_cluster")]
    for env, sym in knobs:
        if os.environ.get(env):
            if not hasattr(lib, sym):
                raise CnError("%s needs the hooks build: set CN_TEST_HOOKS=1 (%s is not in the product library)" % (env, sym))
            getattr(lib, sym)(int(os.environ[env]))
    if os.environ.get("CN_FOLD"):               # "fold,s2all" e.g. "0,1": folded upsample+conv plans / merged stride-2 dgrad phases
        if not hasattr(lib, "cn_debug_set_fold"):
            raise CnError("CN_FOLD needs the hooks build: set CN_TEST_HOOKS=1")
        a, b = (os.environ["CN_FOLD"].split(",") + ["1"])[:2]
        lib.cn_debug_set_fold(int(a), int(b))
    _lib = lib
    return lib


def load_hooks():
    """The hooks build of the library as a SEPARATE handle (tests of the host-evaluated plan geometry, cn_debug_conv_host):
    it shares no state with the product library the ops go through."""
    if not os.path.exists(HOOKS_LIB_PATH):
        raise CnError("libconfignet_b200_hooks.so not built (%s); run __graft_entry__.build()" % HOOKS_LIB_PATH)
    lib = ctypes.CDLL(HOOKS_LIB_PATH)
    lib.cn_debug_conv_host.restype = ctypes.c_int
    return lib


def check(rc):
    if rc != 0:
        raise CnError("libconfignet_b200 error %d: %s" % (rc, load().cn_last_error().decode()))


def call(name, *args):
    check(getattr(load(), name)(*args))
