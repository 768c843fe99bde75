"""Differentiable operators of the ConfigNet hot path, each a thin launch of libconfignet_b200.so.

torch is the host container only: tensors own the device memory, ``torch.autograd`` plays the role
of ``tf.GradientTape`` (confignet_first_stage.py:469-557, losses.py:26,57).  Every arithmetic kernel
is ours.  The discriminator operators are closed under differentiation (their backward passes are
themselves Functions with hand-derived backward passes) because the R1 penalty differentiates the
input gradient (losses.py:42-43,75-82).

There is no CPU implementation: tensors must live on a CUDA device.
"""
import ctypes
import contextlib
import os
import torch
from . import _lib as L

_PARAM_GRADS = [True]


@contextlib.contextmanager
def input_grad_only():
    """Inside this context the backward passes skip parameter gradients (weights, biases, gamma/beta).
    Used for ``tape.gradient(out, real_imgs)`` (losses.py:79), which only needs d/d(input)."""
    _PARAM_GRADS.append(False)
    try:
        yield
    finally:
        _PARAM_GRADS.pop()


def _want_param_grads():
    return _PARAM_GRADS[-1]


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise L.CnError("confignet_b200 operators need CUDA tensors (there is no CPU path)")
    if t.dtype != torch.float32:
        raise L.CnError("confignet_b200 operators are fp32 (got %s)" % t.dtype)
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# convolution / dense
# ------------------------------------------------------------------------------------------------
class ConvGeom:
    """Geometry of one Keras Conv2D / Conv3D / Dense layer call (cn_conv_desc + cached shapes)."""
    _cache = {}

    def __init__(self, nd, batch, in_dims, cin, cout, ksize, stride, upsample, pad=-1):
        self.desc = L.make_conv_desc(nd, batch, in_dims, cin, cout, ksize, stride, upsample, pad)
        self.nd, self.batch, self.cin, self.cout = nd, batch, cin, cout
        self.in_dims, self.ksize = tuple(in_dims), tuple(ksize)
        od = (ctypes.c_int * 3)()
        L.call("cn_conv_out_dims", ctypes.byref(self.desc), od)
        self.out_dims = tuple(od[:nd])
        self.x_shape = (batch,) + self.in_dims + (cin,)
        self.y_shape = (batch,) + self.out_dims + (cout,)
        self.w_shape = self.ksize + (cin, cout)
        self.ref = ctypes.byref(self.desc)

    @classmethod
    def get(cls, x_shape, w_shape, stride=1, upsample=1, pad=-1):
        key = (tuple(x_shape), tuple(w_shape), stride, upsample, pad)
        g = cls._cache.get(key)
        if g is None:
            nd = len(w_shape) - 2
            if len(x_shape) != nd + 2:
                raise L.CnError("input rank %d does not match kernel rank %d" % (len(x_shape), len(w_shape)))
            if x_shape[-1] != w_shape[-2]:
                raise L.CnError("channel mismatch: x %s kernel %s" % (tuple(x_shape), tuple(w_shape)))
            g = cls(nd, x_shape[0], x_shape[1:-1], w_shape[-2], w_shape[-1], w_shape[:nd], stride, upsample, pad)
            cls._cache[key] = g
        return g


IMPL = [L.IMPL_AUTO]     # global kernel selection (tests switch it to force one family)
PROFILE = [None]         # bench.py: list collecting (op, algorithmic flops, start event, end event, impl)


def _flops(g):
    m = g.batch
    for d in g.out_dims:
        m *= d
    k = g.cin
    for d in g.ksize:
        k *= d
    return 2.0 * m * k * g.cout


def _exec_flops(g):
    """FLOPs actually executed: the sub-pixel folded plans of UpSampling(2)+conv (csrc/conv.cu, csrc/skinny.cu) sum the
    taps that land on the same low-resolution pixel, so a k-tap axis costs (k+1)/2 taps per output instead of k."""
    f = _flops(g)
    if g.desc.upsample == 2:
        for k in g.ksize:
            f *= (k + 1) / (2.0 * k)
    return f


def _timed(op, g, fn):
    prof = PROFILE[0]
    if prof is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    prof.append((op, _flops(g), e0, e1, L.load().cn_last_conv_impl(), g.desc.key(), _exec_flops(g)))


def _conv_fwd_raw(g, x, w, bias, act=L.ACT_NONE, alpha=0.0):
    y = torch.empty(g.y_shape, device=x.device, dtype=torch.float32)
    _timed("fwd", g, lambda: L.call("cn_conv_fwd", g.ref, _p(x), _p(w), _p(bias), act, alpha, _p(y), IMPL[0], _stream()))
    return y


def _conv_dgrad_raw(g, gy, w):
    gx = torch.empty(g.x_shape, device=gy.device, dtype=torch.float32)
    _timed("dgrad", g, lambda: L.call("cn_conv_dgrad", g.ref, _p(gy), _p(w), _p(gx), IMPL[0], _stream()))
    return gx


def _conv_wgrad_raw(g, x, gy, want_bias):
    gw = torch.empty(g.w_shape, device=x.device, dtype=torch.float32)
    gb = torch.empty(g.cout, device=x.device, dtype=torch.float32) if want_bias else None
    _timed("wgrad", g, lambda: L.call("cn_conv_wgrad", g.ref, _p(x), _p(gy), _p(gw), _p(gb), IMPL[0], _stream()))
    return gw, gb


class ConvFwd(torch.autograd.Function):
    """y = conv_same(x, w) + bias, no activation; differentiable to any order."""

    @staticmethod
    def forward(ctx, x, w, bias, g):
        x, w, bias = _chk(x), _chk(w), _chk(bias)
        ctx.g = g
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w)
        return _conv_fwd_raw(g, x, w, bias)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        g = ctx.g
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ConvDgrad.apply(gy, w, g)
        if _want_param_grads() and (ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])):
            gw, gb = ConvWgrad.apply(x, gy, g, ctx.has_bias)
        return gx, gw, gb, None


class ConvDgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, w, g):
        gy, w = _chk(gy), _chk(w)
        ctx.g = g
        ctx.save_for_backward(gy, w)
        return _conv_dgrad_raw(g, gy, w)

    @staticmethod
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        g = ctx.g
        d_gy = d_w = None
        if ctx.needs_input_grad[0]:
            d_gy = ConvFwd.apply(ggx, w, None, g)
        if ctx.needs_input_grad[1] and _want_param_grads():
            d_w, _ = ConvWgrad.apply(ggx, gy, g, False)
        return d_gy, d_w, None


class ConvWgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gy, g, want_bias):
        x, gy = _chk(x), _chk(gy)
        ctx.g = g
        ctx.save_for_backward(x, gy)
        return _conv_wgrad_raw(g, x, gy, want_bias)

    @staticmethod
    def backward(ctx, ggw, ggb):
        x, gy = ctx.saved_tensors
        g = ctx.g
        d_x = d_gy = None
        if ctx.needs_input_grad[0]:
            d_x = ConvDgrad.apply(gy, ggw, g)
        if ctx.needs_input_grad[1]:
            d_gy = ConvFwd.apply(x, ggw, ggb, g)
        return d_x, d_gy, None, None


def conv(x, w, bias=None, stride=1):
    """Keras Conv2D/Conv3D padding='same' (or Dense when w is 2-D); any-order differentiable."""
    return ConvFwd.apply(x, w, bias, ConvGeom.get(x.shape, w.shape, stride, 1))


def dense(x, w, bias=None):
    return conv(x, w, bias)


class ConvActFwd(torch.autograd.Function):
    """y = act(conv_same(upsample(x), w) + bias) with the activation in the kernel epilogue.
    First-order only (generator, VGG, latent regressor).  With ``grad_is_preact`` the incoming
    gradient is already d/d(pre-activation) (the AdaIN backward kernel applied lrelu')."""

    @staticmethod
    def forward(ctx, x, w, bias, g, act, alpha, grad_is_preact):
        x, w, bias = _chk(x), _chk(w), _chk(bias)
        y = _conv_fwd_raw(g, x, w, bias, act, alpha)
        ctx.g, ctx.act, ctx.alpha, ctx.preact = g, act, alpha, grad_is_preact
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w, y if (act != L.ACT_NONE and not grad_is_preact) else None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        g = ctx.g
        gy = _chk(gy)
        if ctx.act != L.ACT_NONE and not ctx.preact:
            gpre = torch.empty_like(gy)
            L.call("cn_act_bwd", _p(gy), _p(y), ctx.act, ctx.alpha, _p(gpre), gy.numel(), _stream())
            gy = gpre
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _conv_dgrad_raw(g, gy, w)
        if _want_param_grads() and (ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])):
            gw, gb = _conv_wgrad_raw(g, x, gy, ctx.has_bias)
        return gx, gw, gb, None, None, None, None


class ConvActSqdiff(torch.autograd.Function):
    """y = act(conv_same(x, w) + bias) AND loss = scale * sum((y - target)^2) as one node (a tapped VGG layer of the
    perceptual loss, perceptual_loss.py:61-82).  The backward folds the loss gradient, its accumulation onto the gradient
    coming from the next layer and the activation derivative into ONE pass (cn_act_bwd_sqdiff) - separately they are
    reduce_bwd + an add by the autograd engine + act_bwd, 9 tensor passes instead of 4 - with the same arithmetic, so the
    gradients are bit-identical to the unfused graph.  First order only (the VGG path is never differentiated twice)."""

    @staticmethod
    def forward(ctx, x, w, bias, g, act, alpha, target, scale):
        x, w, bias, target = _chk(x), _chk(w), _chk(bias), _chk(target)
        y = _conv_fwd_raw(g, x, w, bias, act, alpha)
        res = torch.empty(1, device=x.device, dtype=torch.float32)
        L.call("cn_reduce", _p(y), _p(target), None, 1, y.numel(), RED_SQDIFF, 1.0, scale, _p(_ws(x.device)), _p(res), _stream())
        ctx.g, ctx.act, ctx.alpha, ctx.scale = g, act, alpha, scale
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w, y, target)
        ctx.set_materialize_grads(False)
        return y, res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy, gres):
        x, w, y, target = ctx.saved_tensors
        g = ctx.g
        if gy is None and gres is None:
            return (None,) * 8
        gpre = torch.empty_like(y)
        if gres is None:
            L.call("cn_act_bwd", _p(_chk(gy)), _p(y), ctx.act, ctx.alpha, _p(gpre), y.numel(), _stream())
        else:
            L.call("cn_act_bwd_sqdiff", _p(_chk(gy)), _p(y), _p(target), _p(_chk(gres)), ctx.scale, ctx.act, ctx.alpha, _p(gpre),
                   y.numel(), _stream())
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _conv_dgrad_raw(g, gpre, w)
        if _want_param_grads() and (ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])):
            gw, gb = _conv_wgrad_raw(g, x, gpre, ctx.has_bias)
        return gx, gw, gb, None, None, None, None, None


def conv_act_sqdiff(x, w, bias, target, scale, act=L.ACT_RELU, alpha=0.0):
    """(activation, scale * sum((activation - target)^2)); needs numel % 4 == 0 (every VGG activation)."""
    g = ConvGeom.get(x.shape, w.shape, 1, 1, -1)
    return ConvActSqdiff.apply(x, w, bias, g, act, alpha, target, scale)


def conv_act(x, w, bias=None, stride=1, upsample=1, act=L.ACT_NONE, alpha=0.0, grad_is_preact=False, pad=-1):
    """pad = -1: TF "SAME"; pad >= 0: ZeroPadding(pad) + "VALID" (ResNet50 stem)."""
    g = ConvGeom.get(x.shape, w.shape, stride, upsample, pad)
    return ConvActFwd.apply(x, w, bias, g, act, alpha, grad_is_preact)


# ------------------------------------------------------------------------------------------------
# per-(sample, channel) statistics family
# ------------------------------------------------------------------------------------------------
FLAG_LRELU_A, FLAG_MASK_OUT, FLAG_MASK_C = 1, 2, 4
(COEF_IN_FWD, COEF_IN_BWD, COEF_IN_BWDBWD, COEF_STYLE_FWD, COEF_STYLE_BWD, COEF_STYLE_BWDBWD,
 COEF_ADAIN_FWD, COEF_ADAIN_BWD) = range(8)


_SPLITS = {}     # (n, p, ch) -> pixel slices cn_chan_sums writes


def _npc(x):
    n, c = x.shape[0], x.shape[-1]
    return n, x.numel() // (n * c), c


def _sums(a, b=None, c=None, flags=0, alpha=0.0):
    n, p, ch = _npc(a)
    ns = _SPLITS.get((n, p, ch))
    if ns is None:
        ns = _SPLITS[(n, p, ch)] = int(L.load().cn_chan_sums_splits(n, p, ch))
    s = torch.empty((ns, n, ch, 8), device=a.device, dtype=torch.float32)       # 7 sums + 1 pad per record (csrc/norm.cu)
    L.call("cn_chan_sums", _p(a), _p(b), _p(c), n, p, ch, flags, alpha, _p(s), _stream())
    return s


def _sums_dual(a, alpha):
    """(sums of lrelu(a), sums of a) in one pass over a (cn_chan_sums_dual); two passes where the one-pass kernel does not
    take the channel count.  Same records, bit for bit, as _sums(a, flags=FLAG_LRELU_A) and _sums(a)."""
    n, p, ch = _npc(a)
    if ch % 4 != 0 or ch > 1024:
        return _sums(a, flags=FLAG_LRELU_A, alpha=alpha), _sums(a)
    ns = _SPLITS.get((n, p, ch))
    if ns is None:
        ns = _SPLITS[(n, p, ch)] = int(L.load().cn_chan_sums_splits(n, p, ch))
    s_act = torch.empty((ns, n, ch, 8), device=a.device, dtype=torch.float32)
    s_raw = torch.empty((ns, n, ch, 8), device=a.device, dtype=torch.float32)
    L.call("cn_chan_sums_dual", _p(a), n, p, ch, alpha, _p(s_act), _p(s_raw), _stream())
    return s_act, s_raw


def _affine(a, b, c, coef, flags=0, alpha=0.0):
    n, p, ch = _npc(a)
    out = torch.empty_like(a)
    L.call("cn_chan_affine", _p(a), _p(b), _p(c), _p(coef), n, p, ch, flags, alpha, _p(out), _stream())
    return out


def _coef(kind, sums, p0, p1, n, ch, npix, eps, ncoef=1, out0_shape=None, out1_shape=None):
    dev = sums.device
    coefs = [torch.empty((n, ch, 4), device=dev, dtype=torch.float32) for _ in range(ncoef)]
    out0 = torch.empty(out0_shape, device=dev, dtype=torch.float32) if out0_shape else None
    out1 = torch.empty(out1_shape, device=dev, dtype=torch.float32) if out1_shape else None
    L.call("cn_norm_coef", kind, _p(sums), sums.shape[0], _p(p0), _p(p1), n, ch, npix, eps,
           _p(coefs[0]) if ncoef > 0 else None, _p(coefs[1]) if ncoef > 1 else None, _p(out0), _p(out1), _stream())
    return coefs, out0, out1


IN_EPS = 1e-3       # instance_normalization.py:46 (added to the std)
STYLE_EPS = 1e-6    # confignet_utils.py:154
ADAIN_EPS = 1e-3    # keras LayerNormalization default epsilon [TF-2.1]


def _in_bwd_bwd(c, gamma, gy, h, alpha, need_c, need_gy):
    """Second-order terms of InstanceNorm(LeakyReLU(c)): given the cotangent h of the input gradient, the gradients wrt c,
    gamma and gy (shared by LReluInstanceNormBwd and the fused DiscrNormBwd)."""
    h = _chk(h)
    n, p, ch = _npc(c)
    fl = FLAG_LRELU_A | FLAG_MASK_C
    s = _sums(c, gy, h, flags=fl, alpha=alpha)
    (coef_a, coef_g), dgamma, _ = _coef(COEF_IN_BWDBWD, s, gamma, None, n, ch, p, IN_EPS, 2, (ch,))
    d_c = d_gy = None
    if need_c and need_gy and ch % 4 == 0 and ch <= 1024 and n <= 65535 and os.environ.get("CN_AFFINE_ROWS", "1") != "0":
        # both results share (c, gy, h): one pass, 3 reads + 2 writes (same arithmetic as the two calls below)
        d_c, d_gy = torch.empty_like(c), torch.empty_like(c)
        L.call("cn_chan_affine_pair", _p(c), _p(gy), _p(h), _p(coef_a), _p(coef_g), n, p, ch, alpha, _p(d_c), _p(d_gy), _stream())
    else:
        if need_c:
            d_c = _affine(c, gy, h, coef_a, fl | FLAG_MASK_OUT, alpha)
        if need_gy:
            d_gy = _affine(c, None, h, coef_g, fl, alpha)
    return d_c, (dgamma if _want_param_grads() else None), d_gy


def _style_bwd_bwd(c, gstyle, h, need_c):
    """Second-order terms of get_layer_style: gradients wrt c and gstyle given the cotangent h of the input gradient."""
    h = _chk(h)
    n, p, ch = _npc(c)
    s = _sums(c, h)
    (coef,), d_gstyle, _ = _coef(COEF_STYLE_BWDBWD, s, gstyle, None, n, ch, p, STYLE_EPS, 1, (n, 2 * ch))
    d_c = _affine(c, h, None, coef) if need_c else None
    return d_c, d_gstyle


def _affine2(a, b, coef, coef2, flags=0, alpha=0.0):
    """[affine(a, b, coef, flags)] + coef2.x * a + coef2.w in one pass (cn_chan_affine2); channel counts the one-pass
    kernel does not take fall back to two passes and an add (same values, same rounding)."""
    n, p, ch = _npc(a)
    if ch % 4 == 0 and ch <= 1024 and n <= 65535:
        out = torch.empty_like(a)
        try:
            L.call("cn_chan_affine2", _p(a), _p(b), None, _p(coef), _p(coef2), n, p, ch, flags, alpha, _p(out), _stream())
            return out
        except L.CnError as e:                     # CN_AFFINE_ROWS=0 (A/B runs): the row-walking kernel is switched off
            if "two-pass" not in str(e):
                raise
    return _affine(a, b, None, coef, flags, alpha) + _affine(a, None, None, coef2)


class LReluInstanceNorm(torch.autograd.Function):
    """y = InstanceNormalization(LeakyReLU(alpha)(c)) (building_blocks.py:104-106,
    instance_normalization.py:108-131); any-order differentiable."""

    @staticmethod
    def forward(ctx, c, gamma, beta, alpha):
        c, gamma, beta = _chk(c), _chk(gamma), _chk(beta)
        n, p, ch = _npc(c)
        s = _sums(c, flags=FLAG_LRELU_A, alpha=alpha)
        (coef,), _, _ = _coef(COEF_IN_FWD, s, gamma, beta, n, ch, p, IN_EPS)
        ctx.alpha = alpha
        ctx.save_for_backward(c, gamma)
        return _affine(c, None, None, coef, FLAG_LRELU_A, alpha)

    @staticmethod
    def backward(ctx, gy):
        c, gamma = ctx.saved_tensors
        gc, ggamma, gbeta = LReluInstanceNormBwd.apply(c, gamma, gy, ctx.alpha)
        if not _want_param_grads():
            ggamma = gbeta = None
        return gc, ggamma, gbeta, None


class LReluInstanceNormBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c, gamma, gy, alpha):
        c, gamma, gy = _chk(c), _chk(gamma), _chk(gy)
        n, p, ch = _npc(c)
        s = _sums(c, gy, flags=FLAG_LRELU_A, alpha=alpha)
        (coef,), dgamma, dbeta = _coef(COEF_IN_BWD, s, gamma, None, n, ch, p, IN_EPS, 1, (ch,), (ch,))
        ctx.alpha = alpha
        ctx.save_for_backward(c, gamma, gy)
        ctx.set_materialize_grads(False)
        return _affine(c, gy, None, coef, FLAG_LRELU_A | FLAG_MASK_OUT, alpha), dgamma, dbeta

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, h, h_dgamma, h_dbeta):
        if h_dgamma is not None or h_dbeta is not None:
            raise NotImplementedError("second-order terms through d(gamma)/d(beta) are not on ConfigNet's path")
        c, gamma, gy = ctx.saved_tensors
        if h is None:
            return None, None, None, None
        d_c, dgamma, d_gy = _in_bwd_bwd(c, gamma, gy, h, ctx.alpha, ctx.needs_input_grad[0], ctx.needs_input_grad[2])
        return d_c, dgamma, d_gy, None


def lrelu_instance_norm(c, gamma, beta, alpha=0.3):
    return LReluInstanceNorm.apply(c, gamma, beta, alpha)


class LayerStyle(torch.autograd.Function):
    """get_layer_style (confignet_utils.py:147-159): (n, 2C) = concat(mean, sqrt(var + 1e-6))."""

    @staticmethod
    def forward(ctx, c):
        c = _chk(c)
        n, p, ch = _npc(c)
        s = _sums(c)
        _, style, _ = _coef(COEF_STYLE_FWD, s, None, None, n, ch, p, STYLE_EPS, 0, (n, 2 * ch))
        ctx.save_for_backward(c)
        return style

    @staticmethod
    def backward(ctx, gstyle):
        (c,) = ctx.saved_tensors
        return LayerStyleBwd.apply(c, gstyle)


class LayerStyleBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c, gstyle):
        c, gstyle = _chk(c), _chk(gstyle)
        n, p, ch = _npc(c)
        s = _sums(c)
        (coef,), _, _ = _coef(COEF_STYLE_BWD, s, gstyle, None, n, ch, p, STYLE_EPS)
        ctx.save_for_backward(c, gstyle)
        return _affine(c, None, None, coef)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, h):
        c, gstyle = ctx.saved_tensors
        return _style_bwd_bwd(c, gstyle, h, ctx.needs_input_grad[0])


def layer_style(c):
    return LayerStyle.apply(c)


class DiscrNorm(torch.autograd.Function):
    """The two consumers of a DiscrBlock's conv output c (building_blocks.py:100-106): style = get_layer_style(c) and
    y = InstanceNormalization(LeakyReLU(c)), as ONE autograd node.  Values are those of layer_style(c) and
    lrelu_instance_norm(c, ...); the point is the backward: when both outputs carry a gradient the two input gradients
    leave in one pass (DiscrNormBwd) instead of statistics + affine per consumer and an add by the autograd engine
    (11 -> 5 tensor passes), and the raw statistics of the forward pass are reused.  Any-order differentiable."""

    @staticmethod
    def forward(ctx, c, gamma, beta, alpha):
        c, gamma, beta = _chk(c), _chk(gamma), _chk(beta)
        n, p, ch = _npc(c)
        s, s_raw = _sums_dual(c, alpha)
        _, style, _ = _coef(COEF_STYLE_FWD, s_raw, None, None, n, ch, p, STYLE_EPS, 0, (n, 2 * ch))
        (coef,), _, _ = _coef(COEF_IN_FWD, s, gamma, beta, n, ch, p, IN_EPS)
        ctx.alpha = alpha
        ctx.save_for_backward(c, gamma, s_raw)
        ctx.set_materialize_grads(False)
        return _affine(c, None, None, coef, FLAG_LRELU_A, alpha), style

    @staticmethod
    def backward(ctx, gy, gstyle):
        c, gamma, s_raw = ctx.saved_tensors
        gc = dgamma = dbeta = None
        if gy is not None and gstyle is not None:
            gc, dgamma, dbeta = DiscrNormBwd.apply(c, gamma, gy, gstyle, s_raw, ctx.alpha)
        elif gy is not None:                               # R1 passes of a later head: only the trunk carries a gradient
            gc, dgamma, dbeta = LReluInstanceNormBwd.apply(c, gamma, gy, ctx.alpha)
        elif gstyle is not None:                           # R1 pass of this block's own style head
            gc = LayerStyleBwd.apply(c, gstyle)
        if not _want_param_grads():
            dgamma = dbeta = None
        return gc, dgamma, dbeta, None


class DiscrNormBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c, gamma, gy, gstyle, s_raw, alpha):
        c, gamma, gy, gstyle = _chk(c), _chk(gamma), _chk(gy), _chk(gstyle)
        n, p, ch = _npc(c)
        s = _sums(c, gy, flags=FLAG_LRELU_A, alpha=alpha)
        (coef_in,), dgamma, dbeta = _coef(COEF_IN_BWD, s, gamma, None, n, ch, p, IN_EPS, 1, (ch,), (ch,))
        (coef_st,), _, _ = _coef(COEF_STYLE_BWD, s_raw, gstyle, None, n, ch, p, STYLE_EPS)
        ctx.alpha = alpha
        ctx.save_for_backward(c, gamma, gy, gstyle)
        ctx.set_materialize_grads(False)
        return _affine2(c, gy, coef_in, coef_st, FLAG_LRELU_A | FLAG_MASK_OUT, alpha), dgamma, dbeta

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, h, h_dgamma, h_dbeta):
        if h_dgamma is not None or h_dbeta is not None:
            raise NotImplementedError("second-order terms through d(gamma)/d(beta) are not on ConfigNet's path")
        c, gamma, gy, gstyle = ctx.saved_tensors
        if h is None:
            return None, None, None, None, None, None
        need_c = ctx.needs_input_grad[0]
        d_c1, dgamma, d_gy = _in_bwd_bwd(c, gamma, gy, h, ctx.alpha, need_c, ctx.needs_input_grad[2])
        d_c2, d_gstyle = _style_bwd_bwd(c, gstyle, h, need_c)
        d_c = (d_c1 + d_c2) if need_c else None
        return d_c, dgamma, d_gy, (d_gstyle if ctx.needs_input_grad[3] else None), None, None


def discr_norm(c, gamma, beta, alpha=0.3):
    """(InstanceNormalization(LeakyReLU(alpha)(c)), get_layer_style(c)) - see DiscrNorm."""
    return DiscrNorm.apply(c, gamma, beta, alpha)


class AdaIN(torch.autograd.Function):
    """AdaIn.call (building_blocks.py:135-149) on a = LeakyReLU(conv) computed by the conv epilogue.
    sb = MLP(z) (n, 2C): first C = scale, last C = bias.  With ``mask_alpha`` the returned input
    gradient is multiplied by lrelu'(a), i.e. it is the gradient wrt the conv pre-activation."""

    @staticmethod
    def forward(ctx, a, sb, mask_alpha):
        a, sb = _chk(a), _chk(sb)
        n, p, ch = _npc(a)
        s = _sums(a)
        (coef,), _, _ = _coef(COEF_ADAIN_FWD, s, sb, None, n, ch, p, ADAIN_EPS)
        ctx.mask_alpha = mask_alpha
        ctx.save_for_backward(a, sb)
        return _affine(a, None, None, coef)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        a, sb = ctx.saved_tensors
        gy = _chk(gy)
        n, p, ch = _npc(a)
        s = _sums(a, gy)
        (coef,), dsb, _ = _coef(COEF_ADAIN_BWD, s, sb, None, n, ch, p, ADAIN_EPS, 1, (n, 2 * ch))
        if ctx.mask_alpha is None:
            ga = _affine(a, gy, None, coef)
        else:
            ga = _affine(a, gy, None, coef, FLAG_MASK_OUT, ctx.mask_alpha)
        return ga, dsb, None


def adain(a, sb, mask_alpha=None):
    return AdaIN.apply(a, sb, mask_alpha)


# ------------------------------------------------------------------------------------------------
# elementwise
# ------------------------------------------------------------------------------------------------
class LRelu(torch.autograd.Function):
    """LeakyReLU; any-order differentiable (latent discriminator MLP under R1, losses.py:49-73)."""

    @staticmethod
    def forward(ctx, x, alpha):
        x = _chk(x)
        y = torch.empty_like(x)
        L.call("cn_lrelu_fwd", _p(x), alpha, _p(y), x.numel(), _stream())
        ctx.alpha = alpha
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return LReluBwd.apply(gy, x, ctx.alpha), None


class LReluBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, x, alpha):
        gy, x = _chk(gy), _chk(x)
        gx = torch.empty_like(gy)
        L.call("cn_act_bwd", _p(gy), _p(x), L.ACT_LRELU, alpha, _p(gx), gy.numel(), _stream())
        ctx.alpha = alpha
        ctx.save_for_backward(x)
        return gx

    @staticmethod
    def backward(ctx, h):
        (x,) = ctx.saved_tensors
        return LReluBwd.apply(h, x, ctx.alpha), None, None


def lrelu(x, alpha):
    return LRelu.apply(x, alpha)


# ------------------------------------------------------------------------------------------------
# VGG pieces, rotation, reductions, image conversion
# ------------------------------------------------------------------------------------------------
class MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x)
        n, h, w, c = x.shape
        y = torch.empty((n, h // 2, w // 2, c), device=x.device, dtype=torch.float32)
        L.call("cn_maxpool2_fwd", _p(x), n, h, w, c, _p(y), _stream())
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, y = ctx.saved_tensors
        gy = _chk(gy)
        n, h, w, c = x.shape
        gx = torch.empty_like(x)
        L.call("cn_maxpool2_bwd", _p(x), _p(y), _p(gy), n, h, w, c, _p(gx), _stream())
        return gx


def maxpool2(x):
    return MaxPool2.apply(x)


class VggPreprocess(torch.autograd.Function):
    """mode 0: keras 'caffe' preprocessing (VGG19 / ResNet50); mode 2: VGGFace means, no channel flip
    (perceptual_loss.py:50-59)."""

    @staticmethod
    def forward(ctx, x, mode):
        x = _chk(x)
        out = torch.empty_like(x)
        ctx.mode = mode
        L.call("cn_vgg_preprocess", _p(x), _p(out), x.numel() // 3, mode, _stream())
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        g = _chk(g)
        out = torch.empty_like(g)
        L.call("cn_vgg_preprocess", _p(g), _p(out), g.numel() // 3, ctx.mode + 1, _stream())
        return out, None


def vgg_preprocess(x, face=False):
    return VggPreprocess.apply(x, 2 if face else 0)


class Rotate3D(torch.autograd.Function):
    """transform_3d_grid_tf (confignet_utils.py:63-120); gradients wrt the volume and, where the rotation is
    predicted or optimised (confignet_second_stage.py:169-170,348,392), wrt the 3x3 matrix."""

    @staticmethod
    def forward(ctx, grid, rot):
        grid, rot = _chk(grid), _chk(rot.reshape(rot.shape[0], 9))
        b, s, c = grid.shape[0], grid.shape[1], grid.shape[-1]
        out = torch.empty_like(grid)
        L.call("cn_rotate3d_fwd", _p(grid), _p(rot), b, s, c, _p(out), _stream())
        ctx.save_for_backward(rot, grid if ctx.needs_input_grad[1] else None)
        ctx.dims = (b, s, c)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        rot, grid = ctx.saved_tensors
        gout = _chk(gout)
        b, s, c = ctx.dims
        gg = grot = None
        if ctx.needs_input_grad[0]:
            gg = torch.empty_like(gout)      # written whole (fixed-point scatter + conversion, csrc/misc.cu)
            L.call("cn_rotate3d_bwd_grid", _p(gout), _p(rot), b, s, c, _p(gg), _stream())
        if ctx.needs_input_grad[1]:
            grot = torch.empty((b, 9), device=gout.device, dtype=torch.float32)
            L.call("cn_rotate3d_bwd_rot", _p(grid), _p(gout), _p(rot), b, s, c, _p(grot), _stream())
        return gg, grot


def rotate3d(grid, rot):
    return Rotate3D.apply(grid, rot)


RED_SOFTPLUS, RED_SQDIFF, RED_SQ, RED_SUM = 0, 1, 2, 3
_ws_cache = {}


def _ws(dev):
    w = _ws_cache.get(dev)
    if w is None:
        w = torch.empty(L.load().cn_reduce_ws_floats(), device=dev, dtype=torch.float32)
        _ws_cache[dev] = w
    return w


class Reduce(torch.autograd.Function):
    """scale * sum_i wgt[i//wdiv] * term(x_i, y_i): softplus(sign*x), (x-y)^2, x^2 or x (first order)."""

    @staticmethod
    def forward(ctx, x, y, wgt, wdiv, kind, sign, scale):
        x, y, wgt = _chk(x), _chk(y), _chk(wgt)
        res = torch.empty(1, device=x.device, dtype=torch.float32)
        L.call("cn_reduce", _p(x), _p(y), _p(wgt), wdiv, x.numel(), kind, sign, scale, _p(_ws(x.device)), _p(res), _stream())
        ctx.save_for_backward(x, y, wgt)
        ctx.args = (wdiv, kind, sign, scale)
        return res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, y, wgt = ctx.saved_tensors
        wdiv, kind, sign, scale = ctx.args
        g = _chk(g)
        gx = torch.empty_like(x)
        L.call("cn_reduce_bwd", _p(x), _p(y), _p(wgt), wdiv, x.numel(), kind, sign, scale, _p(g), _p(gx), _stream())
        gy = None
        if y is not None and ctx.needs_input_grad[1]:
            if kind != RED_SQDIFF:
                raise NotImplementedError("gradient wrt the second operand only exists for the squared difference")
            gy = torch.empty_like(gx)
            L.call("cn_axpby", _p(gx), None, -1.0, 0.0, _p(gy), gx.numel(), _stream())
        return (gx if ctx.needs_input_grad[0] else None), gy, None, None, None, None, None


def reduce_sum(x, kind, y=None, wgt=None, wdiv=1, sign=1.0, scale=1.0):
    return Reduce.apply(x, y, wgt, wdiv, kind, sign, scale)


def to_uint8(x):
    x = _chk(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    L.call("cn_to_uint8", _p(x), _p(out), x.numel(), _stream())
    return out


def from_uint8(x):
    if not x.is_cuda or x.dtype != torch.uint8:
        raise L.CnError("from_uint8 needs a CUDA uint8 tensor")
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    L.call("cn_from_uint8", _p(x), _p(out), x.numel(), _stream())
    return out


def adam_ema_step(p, g, m, v, ema, lr_t, b1, b2, eps, ema_alpha=0.999, gscale=1.0):
    L.call("cn_adam_ema_step", _p(p), _p(g), _p(m), _p(v), _p(ema), p.numel(), lr_t, b1, b2, eps, ema_alpha, gscale, _stream())


def adam_ema_step_dev(p, g, m, v, ema, lr_dev, b1, b2, eps, ema_alpha=0.999, gscale=1.0):
    L.call("cn_adam_ema_step_dev", _p(p), _p(g), _p(m), _p(v), _p(ema), p.numel(), _p(lr_dev), b1, b2, eps, ema_alpha, gscale, _stream())


def ema_update(ema, p, alpha):
    L.call("cn_ema", _p(ema), _p(p), p.numel(), alpha, _stream())


# ------------------------------------------------------------------------------------------------
# second stage / fine-tuning operators (csrc/stage2.cu)
# ------------------------------------------------------------------------------------------------
class EulerToMatrix(torch.autograd.Function):
    """euler_angles_to_matrix (confignet_utils.py:122-145): (B,3) radians -> (B,9), differentiable."""

    @staticmethod
    def forward(ctx, angles):
        angles = _chk(angles)
        b = angles.shape[0]
        rot = torch.empty((b, 9), device=angles.device, dtype=torch.float32)
        L.call("cn_euler_fwd", _p(angles), b, _p(rot), _stream())
        ctx.save_for_backward(angles)
        return rot

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grot):
        (angles,) = ctx.saved_tensors
        grot = _chk(grot)
        ga = torch.empty_like(angles)
        L.call("cn_euler_bwd", _p(angles), _p(grot), angles.shape[0], _p(ga), _stream())
        return ga


def euler_to_matrix(angles):
    return EulerToMatrix.apply(angles)


BN_EPS = 1.001e-5       # keras-applications ResNet50 BatchNormalization epsilon [TF-2.1]


class BnAct(torch.autograd.Function):
    """relu?(BatchNorm_inference(x; gamma, beta, moving_mean, moving_var) + residual): the BN / Add / Activation
    tail of a ResNet50 block.  gamma and beta are trainable, the moving statistics are constants (keras
    BatchNormalization called without training=True, real_encoder.py:27)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mean, var, residual, relu):
        x, gamma, beta, mean, var, residual = [_chk(t) for t in (x, gamma, beta, mean, var, residual)]
        c = x.shape[-1]
        npix = x.numel() // c
        scale = torch.empty(c, device=x.device, dtype=torch.float32)
        shift = torch.empty_like(scale)
        L.call("cn_bn_fold", _p(gamma), _p(beta), _p(mean), _p(var), BN_EPS, c, _p(scale), _p(shift), _stream())
        out = torch.empty_like(x)
        L.call("cn_bn_act_fwd", _p(x), _p(scale), _p(shift), _p(residual), int(relu), _p(out), npix, c, _stream())
        ctx.relu, ctx.has_res = relu, residual is not None
        ctx.save_for_backward(x, out if relu else None, scale, mean, var)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        x, out, scale, mean, var = ctx.saved_tensors
        gout = _chk(gout)
        c = x.shape[-1]
        npix = x.numel() // c
        gx = torch.empty_like(x)
        gres = torch.empty_like(x) if ctx.has_res else None
        dgamma = torch.empty(c, device=x.device, dtype=torch.float32)
        dbeta = torch.empty_like(dgamma)
        L.call("cn_bn_act_bwd", _p(gout), _p(out), _p(x), _p(scale), _p(mean), _p(var), BN_EPS, int(ctx.relu),
               _p(gx), _p(gres), _p(dgamma), _p(dbeta), npix, c, _stream())
        if not _want_param_grads():
            dgamma = dbeta = None
        return gx, dgamma, dbeta, None, None, gres, None


def bn_act(x, gamma, beta, mean, var, residual=None, relu=True):
    return BnAct.apply(x, gamma, beta, mean, var, residual, relu)


class MaxPool3s2(torch.autograd.Function):
    """ZeroPadding2D(1) + MaxPooling2D(3, strides=2) of the ResNet50 stem."""

    @staticmethod
    def forward(ctx, x):
        x = _chk(x)
        n, h, w, c = x.shape
        y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), device=x.device, dtype=torch.float32)
        L.call("cn_maxpool3s2_fwd", _p(x), n, h, w, c, _p(y), _stream())
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, y = ctx.saved_tensors
        gy = _chk(gy)
        n, h, w, c = x.shape
        gx = torch.empty_like(x)
        L.call("cn_maxpool3s2_bwd", _p(x), _p(y), _p(gy), n, h, w, c, _p(gx), _stream())
        return gx


def maxpool3s2(x):
    return MaxPool3s2.apply(x)


class GlobalAvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x)
        n, c = x.shape[0], x.shape[-1]
        p = x.numel() // (n * c)
        y = torch.empty((n, c), device=x.device, dtype=torch.float32)
        L.call("cn_avgpool_fwd", _p(x), n, p, c, _p(y), _stream())
        ctx.shape = tuple(x.shape)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        gy = _chk(gy)
        shape = ctx.shape
        n, c = shape[0], shape[-1]
        gx = torch.empty(shape, device=gy.device, dtype=torch.float32)
        L.call("cn_avgpool_bwd", _p(gy), n, gx.numel() // (n * c), c, _p(gx), _stream())
        return gx


def global_avg_pool(x):
    return GlobalAvgPool.apply(x)


class ColScale(torch.autograd.Function):
    """x * scale[column] with a constant scale (rotation_range_multiplier, real_encoder.py:20-21,30)."""

    @staticmethod
    def forward(ctx, x, scale):
        x, scale = _chk(x), _chk(scale)
        out = torch.empty_like(x)
        L.call("cn_col_scale", _p(x), _p(scale), x.shape[0], x.shape[1], _p(out), _stream())
        ctx.save_for_backward(scale)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (scale,) = ctx.saved_tensors
        g = _chk(g)
        out = torch.empty_like(g)
        L.call("cn_col_scale", _p(g), _p(scale), g.shape[0], g.shape[1], _p(out), _stream())
        return out, None


def col_scale(x, scale):
    return ColScale.apply(x, scale)


class NormLatentLoss(torch.autograd.Function):
    """compute_normalized_latent_regression_loss (confignet_second_stage.py:93-107) on the regressor output and the
    labels (both (B, latent+3)); gradients flow to both."""

    @staticmethod
    def forward(ctx, out, labels, weight, nrot):
        out, labels = _chk(out), _chk(labels)
        b, j = out.shape
        res = torch.empty(1, device=out.device, dtype=torch.float32)
        L.call("cn_norm_latent_loss_fwd", _p(out), _p(labels), b, j, nrot, weight, _p(res), _stream())
        ctx.save_for_backward(out, labels)
        ctx.args = (weight, nrot)
        return res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        out, labels = ctx.saved_tensors
        weight, nrot = ctx.args
        g = _chk(g)
        b, j = out.shape
        go, gl = torch.empty_like(out), torch.empty_like(labels)
        L.call("cn_norm_latent_loss_bwd", _p(out), _p(labels), b, j, nrot, weight, _p(g), _p(go), _p(gl), _stream())
        return (go if ctx.needs_input_grad[0] else None), (gl if ctx.needs_input_grad[1] else None), None, None


def norm_latent_loss(out, labels, weight, nrot=3):
    return NormLatentLoss.apply(out, labels, weight, nrot)
