"""-m gpu: every CUDA operator against the CPU oracle (oracle/confignet_oracle.py) on the same seeded
inputs, through the C ABI.  Tolerances: 1e-4 for fp32 CUDA-core kernels, 1e-3 (the north-star parity
bar is 1e-3, measured as max|a-b|/max|b|) 2e-4 for the tcgen05 kernels (3xTF32: tf32 big/small split, 3 products)."""
import numpy as np
import pytest
import torch

from oracle import confignet_oracle as O
from parity_utils import nerr

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4
TOL_TC = 2e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _conv_case(dev, impl, tol, nd, B, dims, cin, cout, k, s, up, seed=0):
    from confignet_b200 import ops, _lib as L
    torch.manual_seed(seed)
    x = torch.randn(B, *dims, cin)
    w = torch.randn(*([k] * nd), cin, cout) / np.sqrt(cin * k ** nd)
    b = torch.randn(cout)
    xr, wr, br = [t.double().requires_grad_(True) for t in (x, w, b)]
    xu = O.upsample_nearest2(xr) if up == 2 else xr
    yr = (xr @ wr + br) if nd == 0 else O.conv_same(xu, wr, br, s)
    gy = torch.randn(*yr.shape)
    gxr, gwr, gbr = torch.autograd.grad(yr, (xr, wr, br), gy.double())
    old = ops.IMPL[0]
    ops.IMPL[0] = impl
    try:
        xg, wg, bg = [t.to(dev).requires_grad_(True) for t in (x, w, b)]
        y = ops.conv_act(xg, wg, bg, stride=s, upsample=up)
        gx, gw, gb = torch.autograd.grad(y, (xg, wg, bg), gy.to(dev))
    finally:
        ops.IMPL[0] = old
    assert nerr(y, yr) <= tol and nerr(gx, gxr) <= tol and nerr(gw, gwr) <= tol and nerr(gb, gbr) <= tol, \
        (nerr(y, yr), nerr(gx, gxr), nerr(gw, gwr), nerr(gb, gbr))


SMALL = [(2, 2, (8, 8), 3, 5, 4, 1, 1), (2, 2, (8, 8), 3, 5, 3, 2, 1), (2, 1, (7, 9), 2, 3, 3, 2, 1),
         (2, 2, (4, 6), 3, 4, 4, 1, 2), (3, 1, (4, 4, 4), 2, 3, 3, 1, 2), (0, 5, (), 7, 3, 1, 1, 1),
         (2, 2, (16, 16), 3, 48, 3, 2, 1), (2, 2, (16, 16), 32, 3, 4, 1, 2), (0, 8, (), 2048, 148, 1, 1, 1),
         (2, 1, (5, 5), 4, 4, 1, 1, 1),
         # skinny-layer kernels (small-N pixel kernel, skinny wgrad, narrow column sums)
         (2, 2, (64, 64), 32, 3, 4, 1, 2), (2, 4, (128, 128), 3, 48, 3, 2, 1), (2, 2, (128, 128), 3, 3, 1, 1, 1),
         (2, 2, (64, 64), 3, 64, 3, 1, 1), (0, 32, (), 32768, 1, 1, 1, 1),
         # skinny.cu: odd sizes, rows longer than one 128-pixel segment, partial segments
         (2, 1, (9, 11), 3, 48, 3, 2, 1), (2, 1, (6, 300), 3, 48, 3, 2, 1), (2, 1, (4, 300), 3, 64, 3, 1, 1),
         (2, 3, (5, 7), 3, 48, 3, 1, 1), (2, 1, (6, 130), 32, 3, 4, 1, 2), (2, 2, (3, 5), 32, 3, 4, 1, 2),
         (2, 1, (258, 256), 3, 48, 3, 2, 1),
         # fromRGB 1x1 3 -> 3 as a flat stream (skinny.cu p3_*): pixel counts that are / are not multiples of 4
         (2, 1, (5, 7), 3, 3, 1, 1, 1), (2, 3, (64, 64), 3, 3, 1, 1, 1), (2, 1, (1, 1), 3, 3, 1, 1, 1)]
TCS = [(2, 2, (16, 16), 64, 64, 3, 1, 1), (2, 2, (16, 16), 32, 32, 4, 1, 1), (2, 4, (16, 16), 48, 96, 3, 2, 1),
       (2, 2, (8, 8), 64, 32, 4, 1, 2), (3, 2, (4, 4, 4), 64, 32, 3, 1, 2), (3, 1, (8, 8, 8), 32, 64, 3, 1, 1),
       (2, 1, (16, 16), 128, 256, 1, 1, 1), (2, 2, (32, 32), 96, 192, 3, 2, 1), (2, 4, (15, 17), 64, 48, 3, 2, 1),
       (2, 4, (16, 16), 512, 256, 4, 1, 1), (0, 256, (), 256, 512, 1, 1, 1),
       # folded upsample+conv (phased forward, folded dgrad, folded wgrad + unfold) at sizes the tensor-core path takes
       (2, 4, (16, 16), 64, 32, 4, 1, 2), (3, 4, (4, 4, 4), 64, 32, 3, 1, 2), (3, 2, (8, 8, 8), 32, 64, 3, 1, 2),
       (2, 3, (16, 32), 32, 32, 4, 1, 2), (2, 2, (32, 32), 192, 384, 3, 2, 1),
       # the generator's folded 3-D convs at generate_images batch sizes: split-K over the equal-length sub-pixel phases
       (3, 2, (8, 8, 8), 256, 128, 3, 1, 2)]


@pytest.mark.parametrize("cfg", SMALL)
def test_conv_cuda_core(dev, cfg):
    from confignet_b200 import _lib as L
    _conv_case(dev, L.IMPL_FFMA, TOL_FP32, *cfg)


@pytest.mark.parametrize("cfg", TCS)
def test_conv_tcgen05(dev, cfg):
    from confignet_b200 import _lib as L
    _conv_case(dev, L.IMPL_TC, TOL_TC, *cfg)


@pytest.mark.parametrize("cfg", [(3, 1, (4, 4, 4), 512, 256, 3, 1, 2), (3, 1, (8, 8, 8), 256, 128, 3, 1, 2), (3, 1, (16, 16, 16), 128, 64, 3, 1, 1),
                                 (2, 1, (16, 16), 1024, 512, 1, 1, 1), (2, 1, (16, 16), 512, 256, 4, 1, 1), (2, 2, (16, 16), 256, 64, 4, 1, 2)])
@pytest.mark.parametrize("act", ["lrelu", "none"])
def test_conv_small_batch_split_k_with_activation(dev, cfg, act):
    """generate_images at batch 1-2: a handful of tiles with hundreds of k-blocks each.  These launches split K across the SMs
    (slabs + ordered reduction) also when the layer carries a fused LeakyReLU - the reduction kernel applies bias and
    activation - and, for the folded upsample + conv, over the equal-length sub-pixel phases of the one phased launch."""
    from confignet_b200 import ops, _lib as L
    nd, B, dims, cin, cout, k, s, up = cfg
    torch.manual_seed(3)
    x = torch.randn(B, *dims, cin)
    w = torch.randn(*([k] * nd), cin, cout) / np.sqrt(cin * k ** nd)
    b = torch.randn(cout)
    xu = O.upsample_nearest2(x.double()) if up == 2 else x.double()
    yr = O.conv_same(xu, w.double(), b.double(), s)
    if act == "lrelu":
        yr = O.lrelu(yr, 0.3)
    with torch.no_grad():
        y = ops.conv_act(x.to(dev), w.to(dev), b.to(dev), stride=s, upsample=up, act=L.ACT_LRELU if act == "lrelu" else L.ACT_NONE, alpha=0.3)
    assert L.load().cn_last_conv_impl() == 2
    assert nerr(y, yr) <= TOL_TC, nerr(y, yr)


def test_conv_double_backward(dev):
    """conv -> dgrad -> (differentiated again): the closure the R1 penalty relies on."""
    from confignet_b200 import ops, _lib as L
    torch.manual_seed(1)
    x = torch.randn(2, 8, 8, 4); w = torch.randn(3, 3, 4, 6) * 0.2; b = torch.randn(6)
    xr, wr, br = [t.double().requires_grad_(True) for t in (x, w, b)]
    yr = O.conv_same(xr, wr, br, 2)
    gr, = torch.autograd.grad((yr ** 2).sum(), xr, create_graph=True)
    ref = torch.autograd.grad((gr ** 2).sum(), (wr, xr))
    ops.IMPL[0] = L.IMPL_FFMA
    try:
        xg, wg, bg = [t.to(dev).requires_grad_(True) for t in (x, w, b)]
        y = ops.conv(xg, wg, bg, stride=2)
        g, = torch.autograd.grad((y ** 2).sum(), xg, create_graph=True)
        got = torch.autograd.grad((g ** 2).sum(), (wg, xg))
    finally:
        ops.IMPL[0] = L.IMPL_AUTO
    assert nerr(got[0], ref[0]) <= TOL_FP32 and nerr(got[1], ref[1]) <= TOL_FP32


@pytest.mark.parametrize("shape", [(3, 5, 4, 6), (2, 16, 16, 48), (2, 8, 8, 7),
                                   # sample groups x slice groups of the coefficient kernel (batch 16 / 33 / 9), one row group
                                   # per block in the float4 kernels (192 channels), > 1024 channels (grid-stride fallback)
                                   (16, 8, 8, 192), (33, 6, 6, 96), (9, 40, 40, 24), (2, 3, 3, 1028)])
def test_instance_norm_all_orders(dev, shape):
    from confignet_b200 import ops
    torch.manual_seed(2)
    c = torch.randn(*shape); gam = torch.randn(shape[-1]); bet = torch.randn(shape[-1])
    cr, gr, br = [t.double().requires_grad_(True) for t in (c, gam, bet)]
    yr = O.instance_norm_std(O.lrelu(cr, 0.3), gr, br)
    gy = torch.randn(*shape); h = torch.randn(*shape)
    gyr = gy.double().requires_grad_(True)
    g1 = torch.autograd.grad(yr, (cr, gr, br), gyr, create_graph=True)
    g2 = torch.autograd.grad(g1[0], (cr, gr, gyr), h.double())
    cg, gg, bg = [t.to(dev).requires_grad_(True) for t in (c, gam, bet)]
    y = ops.lrelu_instance_norm(cg, gg, bg, 0.3)
    gyg = gy.to(dev).requires_grad_(True)
    q1 = torch.autograd.grad(y, (cg, gg, bg), gyg, create_graph=True)
    q2 = torch.autograd.grad(q1[0], (cg, gg, gyg), h.to(dev))
    assert nerr(y, yr) <= TOL_FP32
    for a, b in zip(q1, g1):
        assert nerr(a, b) <= TOL_FP32
    for a, b in zip(q2, g2):
        assert nerr(a, b) <= 5 * TOL_FP32


@pytest.mark.parametrize("shape", [(3, 5, 4, 6), (2, 16, 16, 48), (16, 8, 8, 192), (33, 6, 6, 96)])
def test_layer_style_all_orders(dev, shape):
    from confignet_b200 import ops
    torch.manual_seed(3)
    c = torch.randn(*shape)
    cr = c.double().requires_grad_(True)
    sr = O.layer_style(cr)
    gs = torch.randn(*sr.shape); h = torch.randn(*shape)
    gsr = gs.double().requires_grad_(True)
    g1, = torch.autograd.grad(sr, cr, gsr, create_graph=True)
    g2 = torch.autograd.grad(g1, (cr, gsr), h.double())
    cg = c.to(dev).requires_grad_(True)
    s = ops.layer_style(cg)
    gsg = gs.to(dev).requires_grad_(True)
    q1, = torch.autograd.grad(s, cg, gsg, create_graph=True)
    q2 = torch.autograd.grad(q1, (cg, gsg), h.to(dev))
    assert nerr(s, sr) <= TOL_FP32 and nerr(q1, g1) <= TOL_FP32
    assert nerr(q2[0], g2[0]) <= 5 * TOL_FP32 and nerr(q2[1], g2[1]) <= 5 * TOL_FP32


@pytest.mark.parametrize("shape", [(3, 5, 4, 6), (2, 16, 16, 48), (16, 8, 8, 192), (9, 12, 12, 24)])
def test_discr_norm_fused_node_all_orders(dev, shape):
    """ops.discr_norm = (InstanceNorm(LeakyReLU(c)), layer_style(c)) as one node (DiscrBlock, building_blocks.py:100-106):
    outputs, the one-pass input gradient when both outputs carry a cotangent (cn_chan_affine2), the single-consumer
    branches, and the second-order terms through the fused backward - against the fp64 oracle; the fused input gradient
    must also equal the sum of the two unfused nodes' gradients bit for bit."""
    from confignet_b200 import ops
    torch.manual_seed(11)
    c = torch.randn(*shape); gam = torch.randn(shape[-1]); bet = torch.randn(shape[-1])
    cr, gr, br = [t.double().requires_grad_(True) for t in (c, gam, bet)]
    yr = O.instance_norm_std(O.lrelu(cr, 0.3), gr, br)
    sr = O.layer_style(cr)
    gy = torch.randn(*shape); gs = torch.randn(*sr.shape); h = torch.randn(*shape)
    gyr, gsr = gy.double().requires_grad_(True), gs.double().requires_grad_(True)
    g1 = torch.autograd.grad((yr, sr), (cr, gr, br), (gyr, gsr), create_graph=True)
    g2 = torch.autograd.grad(g1[0], (cr, gr, gyr, gsr), h.double())
    cg, gg, bg = [t.to(dev).requires_grad_(True) for t in (c, gam, bet)]
    y, st = ops.discr_norm(cg, gg, bg, 0.3)
    gyg, gsg = gy.to(dev).requires_grad_(True), gs.to(dev).requires_grad_(True)
    q1 = torch.autograd.grad((y, st), (cg, gg, bg), (gyg, gsg), create_graph=True)
    q2 = torch.autograd.grad(q1[0], (cg, gg, gyg, gsg), h.to(dev))
    assert nerr(y, yr) <= TOL_FP32 and nerr(st, sr) <= TOL_FP32
    for a, b in zip(q1, g1):
        assert nerr(a, b) <= TOL_FP32
    for a, b in zip(q2, g2):
        assert nerr(a, b) <= 5 * TOL_FP32
    # the one-pass dual statistics write the same records as the two one-operand passes
    s_act, s_raw = ops._sums_dual(cg.detach(), 0.3)
    ref_act, ref_raw = ops._sums(cg.detach(), flags=ops.FLAG_LRELU_A, alpha=0.3), ops._sums(cg.detach())
    for got, want, what in ((s_act, ref_act, "lrelu sums"), (s_raw, ref_raw, "raw sums")):
        assert torch.equal(got[..., 0], want[..., 0]) and torch.equal(got[..., 3], want[..., 3]), what
    # the unfused nodes on the same inputs: same outputs, and gc(fused) == gc(InstanceNorm) + gc(style) exactly
    c2, g2_, b2 = [t.to(dev).requires_grad_(True) for t in (c, gam, bet)]
    y2 = ops.lrelu_instance_norm(c2, g2_, b2, 0.3); st2 = ops.layer_style(c2)
    assert torch.equal(y, y2) and torch.equal(st, st2)
    ga, = torch.autograd.grad(y2, c2, gy.to(dev), retain_graph=True)
    gb, = torch.autograd.grad(st2, c2, gs.to(dev))
    assert torch.equal(q1[0].detach(), ga + gb)
    # single-consumer branches of the fused node (the R1 passes): only y, only the style
    only_y, = torch.autograd.grad(y, cg, gy.to(dev), retain_graph=True)
    only_s, = torch.autograd.grad(st, cg, gs.to(dev), retain_graph=True)
    assert torch.equal(only_y, ga) and torch.equal(only_s, gb)


def test_conv_act_sqdiff_tapped_layer_equals_unfused_graph(dev):
    """ops.conv_act_sqdiff (a tapped VGG layer of the perceptual loss, perceptual_loss.py:61-82): activation, loss term and
    the input gradient - with and without a gradient arriving from the next layer - equal the unfused graph
    (conv_act -> reduce_sum + next layer) bit for bit, and the oracle within the operator tolerance."""
    from confignet_b200 import ops, _lib as L
    torch.manual_seed(21)
    x = torch.randn(2, 12, 10, 8); w = torch.randn(3, 3, 8, 16) * 0.2; b = torch.randn(16) * 0.1
    t = torch.randn(2, 12, 10, 16).abs(); gnext = torch.randn(2, 12, 10, 16); gl = torch.tensor([0.7])
    scale = 1.0 / t.numel()
    xr = x.double().requires_grad_(True)
    yr = torch.relu(O.conv_same(xr, w.double(), b.double(), 1))
    lr = ((yr - t.double()) ** 2).sum() * scale
    for with_next in (True, False):
        ref, = torch.autograd.grad((yr, lr) if with_next else (lr,), xr,
                                   (gnext.double(), gl.double()[0]) if with_next else (gl.double()[0],), retain_graph=True)
        xa = x.to(dev).requires_grad_(True)
        ya, la = ops.conv_act_sqdiff(xa, w.to(dev), b.to(dev), t.to(dev), scale)
        outs, cot = ((ya, la), (gnext.to(dev), gl.to(dev))) if with_next else ((la,), (gl.to(dev),))
        ga, = torch.autograd.grad(outs, xa, cot)
        xb = x.to(dev).requires_grad_(True)
        yb = ops.conv_act(xb, w.to(dev), b.to(dev), act=L.ACT_RELU)
        lb = ops.reduce_sum(yb, ops.RED_SQDIFF, y=t.to(dev), scale=scale)
        outs, cot = ((yb, lb), (gnext.to(dev), gl.to(dev))) if with_next else ((lb,), (gl.to(dev),))
        gb, = torch.autograd.grad(outs, xb, cot)
        assert torch.equal(ya, yb) and torch.equal(la, lb) and torch.equal(ga, gb)
        assert nerr(ya, yr) <= TOL_TC and nerr(la, lr) <= TOL_TC and nerr(ga, ref) <= TOL_TC


@pytest.mark.parametrize("n,ch,side", [(3, 8, 4), (16, 128, 6), (5, 512, 3)])
def test_adain(dev, n, ch, side):
    from confignet_b200 import ops
    torch.manual_seed(4)
    a0 = torch.randn(n, side, side, side, ch); sb = torch.randn(n, 2 * ch)
    pre = a0.double().requires_grad_(True); sbr = sb.double().requires_grad_(True)
    a = O.lrelu(pre, 0.3)
    mean = a.mean((1, 2, 3), keepdim=True); var = ((a - mean) ** 2).mean((1, 2, 3), keepdim=True)
    yr = (a - mean) * torch.rsqrt(var + 1e-3) * (sbr[:, :ch].reshape(n, 1, 1, 1, ch) + 1) + sbr[:, ch:].reshape(n, 1, 1, 1, ch)
    gy = torch.randn(*yr.shape)
    gr = torch.autograd.grad(yr, (pre, sbr), gy.double())
    ag = O.lrelu(a0, 0.3).to(dev).requires_grad_(True); sbg = sb.to(dev).requires_grad_(True)
    y = ops.adain(ag, sbg, mask_alpha=0.3)
    q = torch.autograd.grad(y, (ag, sbg), gy.to(dev))
    assert nerr(y, yr) <= TOL_FP32 and nerr(q[0], gr[0]) <= TOL_FP32 and nerr(q[1], gr[1]) <= TOL_FP32


def test_rotate3d(dev):
    from confignet_b200 import ops, networks
    torch.manual_seed(5)
    B, S, C = 3, 16, 8
    grid = torch.randn(B, S, S, S, C)
    rot = np.array([[0.3, -0.1, 0.0], [0.0, 0.0, 0.0], [-0.5, 0.17, 0.05]], np.float32)
    gr = grid.double().requires_grad_(True)
    outr = O.transform_3d_grid(gr, O.euler_angles_to_matrix(torch.tensor(rot).double()))
    go = torch.randn(*outr.shape)
    ggr, = torch.autograd.grad(outr, gr, go.double())
    gg = grid.to(dev).requires_grad_(True)
    R = torch.from_numpy(networks.euler_angles_to_matrix_np(rot)).to(dev)
    out = ops.rotate3d(gg, R)
    gq, = torch.autograd.grad(out, gg, go.to(dev))
    assert nerr(out, outr) <= TOL_FP32 and nerr(gq, ggr) <= TOL_FP32
    # identity rotation is an exact copy (diffs = 0): bit-exact
    assert torch.equal(out[1].cpu(), grid[1])


def test_maxpool_vggpre_reduce_uint8(dev):
    from confignet_b200 import ops
    torch.manual_seed(6)
    # scalar kernels (5 channels) and the 16-byte kernels (8 / 64 channels); ReLU outputs tie at 0 inside a window: the
    # gradient goes to the first maximum in (dy, dx) order, as in the oracle
    for shape, relu in [((2, 8, 6, 5), False), ((3, 6, 10, 8), True), ((2, 16, 12, 64), True)]:
        x = torch.randn(*shape)
        if relu:
            x = torch.relu(x)
        xr = x.double().requires_grad_(True)
        yr = torch.nn.functional.max_pool2d(xr.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        gy = torch.randn(*yr.shape)
        gxr, = torch.autograd.grad(yr, xr, gy.double())
        xg = x.to(dev).requires_grad_(True)
        y = ops.maxpool2(xg)
        gx, = torch.autograd.grad(y, xg, gy.to(dev))
        assert torch.equal(y.cpu().double(), yr.detach()) and nerr(gx, gxr) == 0.0
    img = torch.rand(2, 4, 4, 3) * 2 - 1
    ir = img.double().requires_grad_(True)
    pr = O.vgg19_preprocess(ir)
    g = torch.randn(*pr.shape)
    gir, = torch.autograd.grad(pr, ir, g.double())
    ig = img.to(dev).requires_grad_(True)
    pg = ops.vgg_preprocess(ig)
    gig, = torch.autograd.grad(pg, ig, g.to(dev))
    assert nerr(pg, pr) <= 1e-6 and nerr(gig, gir) <= 1e-6
    s = torch.randn(37, 1)
    sg = s.to(dev).requires_grad_(True)
    sr = s.double().requires_grad_(True)
    for sign in (-1.0, 1.0):
        lr = torch.nn.functional.softplus(sign * sr).mean()
        lg = ops.reduce_sum(sg, ops.RED_SOFTPLUS, sign=sign, scale=1.0 / 37)
        assert nerr(lg, lr) <= 1e-6
        assert nerr(torch.autograd.grad(lg, sg)[0], torch.autograd.grad(lr, sr)[0]) <= 1e-6
    big = torch.randn(3, 50, 40, 3) * 1.2
    u = ops.to_uint8(big.to(dev)).cpu().numpy()
    assert np.array_equal(u, O.to_uint8_images(big.numpy()))
    raw = torch.randint(0, 256, (2, 5, 5, 3), dtype=torch.uint8)
    f = ops.from_uint8(raw.to(dev)).cpu().numpy()
    assert np.array_equal(f, raw.numpy().astype(np.float32) / 127.5 - 1.0)


def test_adam_ema_matches_keras_formula(dev):
    from confignet_b200.runtime import ParamGroup, KerasAdam
    from collections import OrderedDict
    rng = np.random.RandomState(0)
    arrays = OrderedDict(a=rng.randn(7, 5).astype(np.float32), b=rng.randn(11).astype(np.float32))
    grp = ParamGroup(arrays, dev)
    ref = O.to_torch(arrays)
    opt_ref, opt = O.KerasAdam(), KerasAdam()
    for it in range(3):
        gs = [torch.tensor(rng.randn(*v.shape).astype(np.float32)) for v in arrays.values()]
        opt_ref.apply_gradients(zip(gs, ref.values()))
        grp.pack_grads([g.to(dev) for g in gs])
        opt.apply_flat([grp])
    for got, want in zip(grp.get_weights(), ref.values()):
        assert nerr(got, want) <= 1e-6
