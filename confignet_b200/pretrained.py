"""Pretrained third-party networks the reference downloads at construction time and this offline build cannot:

  VGG19, ImageNet        perceptual_loss.py:19-24 (keras.applications.VGG19)        -> model.perceptual_loss
  VGG16, VGGFace         perceptual_loss.py:26-41 (rcmalli_vggface_tf_notop_vgg16)  -> model.perceptual_loss_face_reco
  ResNet50, ImageNet     dnn_models/real_encoder.py:13 (keras.applications.ResNet50) -> the "resnet/" part of model.encoder

Until real weights are loaded the three networks hold SEEDED STAND-INS (netspec.init_params): every kernel runs and
every parity test holds (the oracle gets the same arrays), but the perceptual / face-recognition losses then measure
distances between random features, and a fresh encoder does not start from ImageNet features.  They are not part of
the ConfigNet checkpoint (the reference does not save them either), so ``load()`` of a trained model does NOT bring them.
The classes warn once, at the first training or fine-tuning call, while a stand-in is still in place.

Loading.  On a machine with TensorFlow / Keras (this one has neither, nor h5py for the .h5 files) export once:

    np.savez("vgg19.npz", *tf.keras.applications.VGG19(weights="imagenet", include_top=False).get_weights())
    np.savez("resnet50.npz", *tf.keras.applications.ResNet50(weights="imagenet", include_top=False, pooling="avg").get_weights())
    np.savez("vggface_vgg16.npz", *vggface_vgg16_notop_model.get_weights())

and pass the files: ``model.load_pretrained_weights(vgg19=..., vggface=..., resnet50=...)``, the config keys
``vgg19_weights`` / ``vggface_weights`` / ``resnet50_weights`` or the environment variables CONFIGNET_VGG19_WEIGHTS /
CONFIGNET_VGGFACE_WEIGHTS / CONFIGNET_RESNET50_WEIGHTS.  An .npz written by ``np.savez(path, *get_weights())`` holds
arr_0, arr_1, ... in Keras' get_weights() order; an .npz keyed by this package's variable names (netspec) works too.
"""
import os
import warnings
import numpy as np

from . import netspec

ENV = {"vgg19": "CONFIGNET_VGG19_WEIGHTS", "vggface": "CONFIGNET_VGGFACE_WEIGHTS", "resnet50": "CONFIGNET_RESNET50_WEIGHTS"}


def _arrays(path_or_arrays):
    """-> (list in file order, dict by name or None)"""
    if isinstance(path_or_arrays, (list, tuple)):
        return [np.asarray(a) for a in path_or_arrays], None
    if isinstance(path_or_arrays, dict):
        return None, {k: np.asarray(v) for k, v in path_or_arrays.items()}
    z = np.load(path_or_arrays, allow_pickle=True)
    if all(k.startswith("arr_") for k in z.files):
        return [z["arr_%d" % i] for i in range(len(z.files))], None
    return None, {k: z[k] for k in z.files}


def _fill(group, names, ordered, by_name, what):
    """names: the group's variables in Keras get_weights() order.  Extra trailing arrays (the layers the reference
    truncates away: VGG19 beyond block4_conv2, VGG16 beyond layer 12) are ignored; shapes are checked."""
    cur = dict(zip(group.names, group.get_weights()))
    if by_name is not None:
        missing = [n for n in names if n not in by_name]
        if missing:
            raise ValueError("%s weights: missing variables %s ..." % (what, missing[:3]))
        src = [by_name[n] for n in names]
    else:
        if len(ordered) < len(names):
            raise ValueError("%s weights: expected at least %d arrays in get_weights() order, got %d" % (what, len(names), len(ordered)))
        src = ordered[:len(names)]
    for n, a in zip(names, src):
        if tuple(a.shape) != tuple(cur[n].shape):
            raise ValueError("%s weights: %s has shape %s, expected %s" % (what, n, tuple(a.shape), tuple(cur[n].shape)))
        cur[n] = np.asarray(a, np.float32)
    group.set_weights([cur[n] for n in group.names])
    group.pretrained = True


def load_vgg(group, path_or_arrays, what):
    """VGG19 / VGG16 trunk: Keras lists kernel, bias per conv layer in layer order - the order of the group itself."""
    ordered, by_name = _arrays(path_or_arrays)
    _fill(group, list(group.names), ordered, by_name, what)


def load_resnet50(group, path_or_arrays):
    """The nested ResNet50 of RealEncoder; the two Dense heads keep their values.  Keras lists a (not nested) ResNet50's
    weights layer by layer: conv kernel, bias, then gamma, beta, moving_mean, moving_variance of its BatchNormalization."""
    ordered, by_name = _arrays(path_or_arrays)
    names = [n for n in group.names if n.startswith("resnet/")]          # real_encoder_spec keeps Keras' per-layer order
    if by_name is not None and not any(k.startswith("resnet/") for k in by_name):
        by_name = {"resnet/" + k: v for k, v in by_name.items()}
    _fill(group, names, ordered, by_name, "ResNet50")


def warn_if_standin(model, names, where):
    """one warning per model and network while a seeded stand-in is in place"""
    told = model.__dict__.setdefault("_standin_warned", set())
    for attr, what in names:
        net = getattr(model, attr, None)
        if net is None or getattr(net.group, "pretrained", False) or attr in told:
            continue
        told.add(attr)
        warnings.warn("%s: %s still holds seeded stand-in weights (the pretrained ones cannot be downloaded offline); "
                      "load them with load_pretrained_weights() - see confignet_b200/pretrained.py" % (where, what), stacklevel=3)


def from_config_or_env(model):
    kw = {}
    for key in ("vgg19", "vggface", "resnet50"):
        p = model.config.get(key + "_weights") or os.environ.get(ENV[key])
        if p:
            kw[key] = p
    if kw:
        model.load_pretrained_weights(**kw)
