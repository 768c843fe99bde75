"""The keras-applications networks on the hot path (VGG19 / VGG16 of the perceptual losses, ResNet50 of the real encoder)
are not reference code and cannot be executed here (no TensorFlow): the oracle restates them.  These
tests pin those restatements to torchvision's implementations of the same published architectures, executed in this
container with the same weights - an independent third party, not Keras itself (DESIGN.md section 4 says which is which)."""
import numpy as np
import pytest
import torch

from confignet_b200 import netspec
from oracle import confignet_oracle as O
from oracle import confignet_oracle_stage2 as O2

tv = pytest.importorskip("torchvision")


def _w(k):
    return torch.nn.Parameter(torch.tensor(np.transpose(k, (3, 2, 0, 1)).copy()))


@pytest.mark.parametrize("which", ["vgg19", "vgg16"])
def test_oracle_vgg_trunks_equal_torchvision(which):
    """Keras VGG19 / VGG16 up to block4_conv2: 3x3 SAME convs with bias + ReLU, 2x2 max-pools; the four tapped layers
    (perceptual_loss.py:21,35) are compared."""
    from torchvision.models import vgg
    if which == "vgg19":
        spec, model, acts, used, layers = netspec.vgg19_spec(), vgg.vgg19(weights=None), O.vgg19_activations, O.VGG19_USED, O.VGG19_LAYERS
    else:
        spec, model, acts, used, layers = netspec.vgg16_spec(), vgg.vgg16(weights=None), O2.vgg16_activations, O2.VGG16_USED, O2.VGG16_LAYERS
    raw = netspec.init_params(spec, 3, vgg_like=True)
    model = model.double().eval()
    convs = [m for m in model.features if isinstance(m, torch.nn.Conv2d)]
    names = [name for kind, name in layers if kind == "conv"]
    with torch.no_grad():
        for conv, name in zip(convs, names):
            conv.weight = _w(raw[name + "/kernel"].astype(np.float64))
            conv.bias = torch.nn.Parameter(torch.tensor(raw[name + "/bias"].astype(np.float64)))
    x = torch.tensor(np.random.RandomState(0).uniform(-120, 130, (1, 32, 40, 3)))
    got = acts(O.to_torch(raw, dtype=torch.float64), x)
    # torchvision's features: conv, relu, (pool) ...; walk it and tap after the ReLU / pool that closes each Keras layer
    want, t, keras_idx = [], x.permute(0, 3, 1, 2), 0
    with torch.no_grad():
        for m in model.features:
            if isinstance(m, torch.nn.Conv2d):
                t = m(t)
                continue
            t = m(t)                                   # ReLU closes a conv layer, MaxPool2d is a layer of its own
            keras_idx += 1
            if keras_idx in used:
                want.append(t.permute(0, 2, 3, 1))
            if keras_idx == len(layers):
                break
    assert len(got) == len(want) == 4
    for a, b in zip(got, want):
        assert tuple(a.shape) == tuple(b.shape) and float((a - b).abs().max() / b.abs().max()) < 1e-10


def test_oracle_resnet50_v1_equals_torchvision_with_v1_strides():
    """keras-applications ResNet50 is the ORIGINAL v1 (stride 2 on the first 1x1 conv of a stage's first block, convs with
    bias, BatchNorm eps 1.001e-5); torchvision ships v1.5 (stride on the 3x3).  Moving the stride back, giving the convs
    their biases and setting eps makes torchvision's network the Keras one: stem 7x7/s2 behind 3 pixels of zero padding,
    3x3/s2 max-pool behind 1 (zeros vs -inf are the same after a ReLU), [3, 4, 6, 3] bottlenecks, projection shortcuts."""
    from torchvision.models import resnet
    raw = netspec.init_real_encoder_params(145, 5)
    p = O.to_torch(raw, dtype=torch.float64)
    model = resnet.resnet50(weights=None).double().eval()

    def load(conv, bn, name):
        conv.weight = _w(raw["resnet/%s_conv/kernel" % name].astype(np.float64))
        conv.bias = torch.nn.Parameter(torch.tensor(raw["resnet/%s_conv/bias" % name].astype(np.float64)))
        q = "resnet/%s_bn" % name
        bn.eps = O2.BN_EPS
        bn.weight = torch.nn.Parameter(torch.tensor(raw[q + "/gamma"].astype(np.float64)))
        bn.bias = torch.nn.Parameter(torch.tensor(raw[q + "/beta"].astype(np.float64)))
        bn.running_mean.copy_(torch.tensor(raw[q + "/moving_mean"])); bn.running_var.copy_(torch.tensor(raw[q + "/moving_variance"]))

    with torch.no_grad():
        load(model.conv1, model.bn1, "conv1")
        for si, layer in enumerate([model.layer1, model.layer2, model.layer3, model.layer4], start=2):
            for b, block in enumerate(layer, start=1):
                q = "conv%d_block%d" % (si, b)
                load(block.conv1, block.bn1, q + "_1"); load(block.conv2, block.bn2, q + "_2"); load(block.conv3, block.bn3, q + "_3")
                block.conv1.stride, block.conv2.stride = block.conv2.stride, (1, 1)          # v1.5 -> v1
                if block.downsample is not None:
                    load(block.downsample[0], block.downsample[1], q + "_0")
        x = torch.tensor(np.random.RandomState(1).uniform(-120, 130, (1, 64, 96, 3)))
        t = model.maxpool(model.relu(model.bn1(model.conv1(x.permute(0, 3, 1, 2)))))
        want = torch.flatten(model.avgpool(model.layer4(model.layer3(model.layer2(model.layer1(t))))), 1)
    got = O2.resnet50_forward(p, x)
    assert tuple(got.shape) == tuple(want.shape) == (1, 2048) and float(want.std()) > 1e-6
    assert float((got - want).abs().max() / want.abs().max()) < 1e-9
