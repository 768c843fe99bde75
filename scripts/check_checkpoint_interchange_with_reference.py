"""Checkpoint interchange with the reference, both directions, by EXECUTING the reference's own
ConfigNetFirstStage.save / load / get_weights / set_weights / initialize_network (confignet_first_stage.py:129-149,
:173-207, :250-289) from /root/reference on the torch-backed TensorFlow stand-in (scripts/tf_torch_shim.py, whose layer
tracking follows [TF-2.1] Layer.__setattr__ / Network.get_weights):

  1. the reference model, its networks filled BY NAME with seeded values, writes <tmp>/ref.{npz,json,_facemodel_distr.pck}
     with its own save(); confignet_b200.ConfigNetFirstStage.load reads them (CPU tensors, no kernel involved) and every
     variable of every network must carry the value that was put under the same name;
  2. the product, re-seeded, writes <tmp>/prod.* with its save(); the reference's own load() classmethod (constructor ->
     initialize_network -> set_weights -> pickle) reads them and every reference layer must carry the product's value.

The one NumPy-version shim: the reference calls np.savez(**{name: [ragged list of arrays]}), which NumPy < 1.24 (the
reference pins TensorFlow 2.1-era NumPy) turns into an object array implicitly and NumPy 2 rejects - np.savez is wrapped to
do that conversion explicitly, which is also what the product's save() does.

Run in the build container only:  python scripts/check_checkpoint_interchange_with_reference.py
(tests/test_host_cpu.py::test_checkpoint_interchange_with_reference runs it when /root/reference exists.)
"""
import importlib
import json
import os
import sys
import tempfile
import types
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
REF = "/root/reference"

import tf_torch_shim as S                                         # noqa: E402
tf = S.install()


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return sys.modules.get(self.__name__ + "." + name, _Anything)


for m in ["matplotlib", "matplotlib.pyplot", "transformations", "azureml", "azureml.core", "azureml.core.run", "cv2",
          "tensorflow.keras.utils", "tensorflow.keras.applications"]:
    parts = m.split(".")
    for i in range(1, len(parts) + 1):
        sys.modules.setdefault(".".join(parts[:i]), _Stub(".".join(parts[:i])))
sys.path.insert(0, REF)
for name, sub in (("confignet", "confignet"), ("confignet.dnn_models", "confignet/dnn_models")):
    pkg = types.ModuleType(name)
    pkg.__path__ = [os.path.join(REF, sub)]
    sys.modules[name] = pkg
for heavy in ["confignet.perceptual_loss", "confignet.metrics", "confignet.metrics.metrics"]:
    sys.modules[heavy] = _Stub(heavy)                             # keras.applications VGG19 + ImageNet weights: not on this path
sys.modules["confignet"].confignet_utils = importlib.import_module("confignet.confignet_utils")
fs = importlib.import_module("confignet.confignet_first_stage")
sys.modules["confignet"].ConfigNetFirstStage = fs.ConfigNetFirstStage          # what confignet/__init__.py exports

from confignet_b200 import netspec                               # noqa: E402
from confignet_b200 import ConfigNetFirstStage, ConfigNet, LatentGAN, load_confignet      # noqa: E402
from oracle import confignet_oracle_stage2 as O2                 # noqa: E402


class ResNet50StandIn(S.Model):
    """keras-applications ResNet50 (library code, not under /root/reference): variables laid out in the functional model's
    layer order (netspec.resnet50_spec), BatchNormalization moving statistics non-trainable; forward = the oracle's."""
    output = types.SimpleNamespace(shape=(None, 2048))

    def __init__(self, weights=None, include_top=True, input_shape=None, pooling=None):
        S.Model.__init__(self)
        assert include_top is False and pooling == "avg"
        self.vars = OrderedDict(("resnet/" + k, torch.zeros(shape, dtype=S.DT)) for k, (shape, _) in netspec.resnet50_spec().items())
        for k, v in self.vars.items():
            if k.endswith("moving_variance") or k.endswith("gamma"):
                v.fill_(1.0)
        self.keras_trainable = [v for k, v in self.vars.items() if netspec.is_trainable(k)]
        self.keras_non_trainable = [v for k, v in self.vars.items() if not netspec.is_trainable(k)]

    def call(self, x):
        return O2.resnet50_forward(self.vars, x)


apps = types.ModuleType("tensorflow.keras.applications")
apps.resnet50 = types.SimpleNamespace(ResNet50=ResNet50StandIn, preprocess_input=lambda x: S.T(x).flip(-1))
sys.modules["tensorflow.keras.applications"] = apps
sys.modules["tensorflow"].keras.applications = apps
second = importlib.import_module("confignet.confignet_second_stage")
lgan = importlib.import_module("confignet.latent_gan")

FM = netspec.default_facemodel_inputs()
RES = 256
NETS = ["generator", "generator_smoothed", "discriminator", "synth_discriminator", "latent_discriminator", "latent_regressor",
        "synthetic_encoder"]


def dense_layers(mlp_simple):
    return [l for l in mlp_simple.map.layers if type(l).__name__ == "Dense"]


def named_tensors(net_name, m):
    """product variable name -> the reference layer variable that plays it (by attribute path, never by position)"""
    t = OrderedDict()

    def kb(prefix, layer, names=("kernel", "bias")):
        for n in names:
            t[prefix + "/" + n] = getattr(layer, n)

    if net_name.startswith("generator"):
        kb("learned_input", m.learned_input_layer)
        for blk in ("map_3d_0", "map_3d_1", "map_2d_0", "map_2d_1", "map_2d_2", "map_2d_2b"):
            b = getattr(m, blk)
            kb(blk + "/conv", getattr(b, "map_3d" if "3d" in blk else "map_2d").layers[0])
            d0, d1 = dense_layers(b.adain.adain_mlp)
            kb(blk + "/adain/dense0", d0); kb(blk + "/adain/dense1", d1)
        kb("map_3d_post/conv0", m.map_3d_post.layers[0]); kb("map_3d_post/conv1", m.map_3d_post.layers[2])
        kb("projection_conv", m.projection_conv); kb("map_final", m.map_final)
    elif net_name in ("discriminator", "synth_discriminator", "latent_regressor"):
        kb("initial_1x1_conv", m.initial_1x1_conv)
        for i, blk in enumerate(m.conv_blocks):
            kb("block%d/conv" % i, blk.map_2d)
            kb("block%d/in" % i, blk.instance_norm, names=("gamma", "beta"))
        if net_name == "latent_regressor":
            kb("latent_predictor", m.latent_predictor)
        else:
            for i, sc in enumerate(m.style_classifiers):
                kb("style%d" % i, sc)
            kb("disc_map", m.disc_map)
    elif net_name == "latent_discriminator":
        for j, d in enumerate(dense_layers(m)):
            kb("mlp/dense%d" % j, d)
    elif net_name == "synthetic_encoder":
        for name in FM:
            d0, d1 = dense_layers(m.per_facemodel_input_mlps[name])
            kb("mlp_%s/dense0" % name, d0); kb("mlp_%s/dense1" % name, d1)
    elif net_name == "encoder":
        t.update(m.resnet.vars)
        kb("rotation_regressor", m.rotation_regressor); kb("feature_to_latent_mlp", m.feature_to_latent_mlp)
    elif net_name.startswith("gan_"):
        for j, d in enumerate(dense_layers(m)):
            kb("mlp/dense%d" % j, d)
    return t


def seeded(shape, seed):
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32)


_savez = np.savez


def savez_ragged(path, **kw):
    """NumPy < 1.24 behaviour for ragged lists (see the module docstring)"""
    conv = {}
    for k, v in kw.items():
        if isinstance(v, list):
            arr = np.empty(len(v), dtype=object)
            arr[:] = v
            v = arr
        conv[k] = v
    return _savez(path, **conv)


def fill_reference(ref, nets, seed):
    n = 0
    for net in nets:
        for name, w in named_tensors(net, getattr(ref, net.replace("gan_", ""))).items():
            seed += 1
            with torch.no_grad():
                w.copy_(torch.from_numpy(seeded(tuple(w.shape), seed)).to(w.dtype))
            n += 1
    return n


def product_matches(prod, ref, nets, seed):
    for net in nets:
        attr = net.replace("gan_", "")
        params = getattr(prod, attr).group.params
        names = list(named_tensors(net, getattr(ref, attr)).keys())
        assert sorted(names) == sorted(params.keys()), net
        for name in names:
            seed += 1
            assert np.array_equal(params[name].detach().numpy(), seeded(tuple(params[name].shape), seed)), (net, name)


def fill_product(prod, nets, seed):
    for net in nets:
        netw = getattr(prod, net.replace("gan_", ""))
        new = []
        for name in netw.group.names:
            seed += 1
            new.append(seeded(tuple(netw.group.params[name].shape), seed))
        netw.group.set_weights(new)                                  # flat-layout order; compared BY NAME below


def reference_matches(ref, prod, nets, seed):
    for net in nets:
        attr = net.replace("gan_", "")
        tensors = named_tensors(net, getattr(ref, attr))
        for name in getattr(prod, attr).group.names:
            seed += 1
            got = tensors[name].detach().numpy().astype(np.float32)
            assert np.array_equal(got, seeded(tuple(tensors[name].shape), seed)), (net, name)


def ref_save(ref, tmp, name):
    np.savez = savez_ragged
    try:
        ref.save(tmp, name)
    finally:
        np.savez = _savez


def round_trip(tag, ref_cls, prod_load, cfg, nets, tmp, with_distributions=True):
    # ---- 1. reference save() -> product load()
    ref = ref_cls(cfg)                                              # the reference's own constructor + initialize_network
    n_vars = fill_reference(ref, nets, 1000)
    if with_distributions:
        ref.facemodel_param_distributions = {"eye_color": {"mean": [0.1, 0.2], "std": 0.5}}
    ref_save(ref, tmp, "ref")
    prod = prod_load(os.path.join(tmp, "ref.json"))
    assert type(prod).__name__ == ref_cls.__name__
    product_matches(prod, ref, nets, 1000)
    norm = lambda v: json.loads(json.dumps(v))                     # tuples come back from .json as lists on both sides
    for k in ref.config:
        assert norm(prod.config[k]) == norm(ref.config[k]), k
    if with_distributions:
        assert prod.facemodel_param_distributions == ref.facemodel_param_distributions
        assert list(prod.config["facemodel_inputs"].keys()) == list(ref.config["facemodel_inputs"].keys())
    print("%-20s reference save() -> product load(): %d variables of %d networks arrive under the right names" % (tag, n_vars, len(nets)))
    # ---- 2. product save() -> reference load()
    fill_product(prod, nets, 5000)
    if with_distributions:
        prod.facemodel_param_distributions = {"hdri_embedding": [1, 2, 3]}
    prod.save(tmp, "prod")
    ref2 = ref_cls.load(os.path.join(tmp, "prod.json"))
    reference_matches(ref2, prod, nets, 5000)
    if with_distributions:
        assert ref2.facemodel_param_distributions == prod.facemodel_param_distributions
    assert ref2.config["latent_dim"] == prod.config["latent_dim"]
    print("%-20s product save() -> reference load(): %d variables arrive under the right names" % (tag, n_vars))
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))


def main():
    tmp = tempfile.mkdtemp(prefix="cn_ckpt_")
    cfg = {"output_shape": (RES, RES, 3), "batch_size": 2, "facemodel_inputs": {k: tuple(v) for k, v in FM.items()}}
    round_trip("ConfigNetFirstStage", fs.ConfigNetFirstStage, lambda p: ConfigNetFirstStage.load(p, device="cpu"), cfg, NETS, tmp)
    # second stage: + real_encoder_weights, whose nested ResNet50 lists trainable before non-trainable variables
    # (netspec.real_encoder_keras_order); load_confignet dispatches on config["model_type"] like confignet/__init__.py
    round_trip("ConfigNet", second.ConfigNet, lambda p: load_confignet(p, device="cpu"), cfg, NETS + ["encoder"], tmp)
    round_trip("LatentGAN", lgan.LatentGAN, lambda p: LatentGAN.load(p, device="cpu"), {"latent_dim": 145},
               ["gan_generator", "gan_generator_smoothed", "gan_discriminator"], tmp, with_distributions=False)
    os.rmdir(tmp)
    print("checkpoint interchange OK")


if __name__ == "__main__":
    main()
