"""CelebA attribute classifier (reference metrics/celeba_attribute_prediction.py): MobileNetV2 + pooled-feature BatchNorm +
sigmoid Dense head, used by the controllability metrics as a frozen predictor.  ``predict_attributes`` and the
``save`` / ``load`` file pair (.json metadata + .npy object array in the Keras get_weights() order) are the reference's;
training the classifier (``train``, a Keras fit loop on CelebA) is a separate offline job of the reference and is not part
of this path."""
import json
import os
import numpy as np
import torch

from .. import ops
from ..runtime import ParamGroup
from . import nets

DEFAULT_CONFIG = {"input_shape": None, "predicted_attributes": None, "optimizer": {"lr": 0.001}, "batch_size": 32}
DEVICE_CHUNK = 256


class CelebaAttributeClassifier:
    def __init__(self, config, device=None, seed=2468):
        self.config = config
        self.logs = {}
        self.device = torch.device(device if device is not None else "cuda:0")
        self._seed = seed
        self.classifier = None
        self.group = None               # folded kernels in HBM, created at the first prediction (the host half runs without a GPU)
        self.initialize_dnn()

    # ---- the Keras model's weights, in classifier.get_weights() order
    def initialize_dnn(self):
        """celeba_attribute_prediction.py:54-62; freshly initialised (seeded) until set_weights / load"""
        n_attr = len(self.config["predicted_attributes"])
        self._order = nets.attribute_classifier_keras_order(n_attr)
        self.set_weights_by_name(nets.init_stand_in(nets.attribute_classifier_spec(n_attr), self._seed))
        self.classifier = self          # reference code reaches the Keras model as .classifier (get_weights / set_weights / predict)

    def set_weights_by_name(self, raw):
        n_attr = len(self.config["predicted_attributes"])
        self._raw = {k: np.asarray(raw[k], np.float32) for k in nets.attribute_classifier_spec(n_attr)}
        if self.group is not None:
            folded = nets.fold_attribute_classifier_params(self._raw, n_attr)
            self.group.set_weights([folded[k] for k in self.group.names])

    def _params(self):
        if self.group is None:
            folded = nets.fold_attribute_classifier_params(self._raw, len(self.config["predicted_attributes"]))
            self.group = ParamGroup(folded, self.device)
            self.group.set_frozen()
        return self.group.params

    def get_weights(self):
        return [self._raw[k].copy() for k in self._order]

    def set_weights(self, weights):
        weights = list(weights)
        if len(weights) != len(self._order):
            raise ValueError("expected %d weight arrays, got %d" % (len(self._order), len(weights)))
        spec = nets.attribute_classifier_spec(len(self.config["predicted_attributes"]))
        for k, w in zip(self._order, weights):
            if tuple(np.shape(w)) != tuple(spec[k][0]):
                raise ValueError("weight %s: expected shape %s, got %s" % (k, tuple(spec[k][0]), tuple(np.shape(w))))
        self.set_weights_by_name(dict(zip(self._order, weights)))

    def save(self, output_dir, output_filename):
        metadata = {"logs": self.logs, "config": self.config}
        with open(os.path.join(output_dir, output_filename + ".json"), "w") as fp:
            json.dump(metadata, fp, indent=4)
        weights = np.empty(len(self._order), dtype=object)
        for i, w in enumerate(self.get_weights()):
            weights[i] = w
        np.save(os.path.join(output_dir, output_filename + ".npy"), weights, allow_pickle=True)

    @classmethod
    def load(cls, file_path, device=None):
        with open(file_path, "r") as fp:
            metadata = json.load(fp)
        weight_file_path = os.path.splitext(file_path)[0] + ".npy"
        weights = np.load(weight_file_path, allow_pickle=True)
        classifier = cls(metadata["config"], device=device)
        classifier.logs = metadata["logs"]
        classifier.classifier.set_weights(weights)
        return classifier

    def train(self, *a, **k):
        raise NotImplementedError("training the CelebA attribute classifier is a separate offline job of the reference "
                                  "(a Keras fit loop on CelebA); load a trained one with CelebaAttributeClassifier.load")

    # ---- inference
    def predict_device(self, preprocessed_images):
        """(B,h,w,3) float32 device tensor in [-1,1] -> (B, n_attributes) probabilities (device)"""
        return nets.attribute_classifier_forward(self._params(), preprocessed_images)

    def predict(self, preprocessed_images):
        x = torch.as_tensor(np.asarray(preprocessed_images, np.float32)).to(self.device)
        return self.predict_device(x).cpu().numpy()

    def predict_attributes(self, input_images):
        """celeba_attribute_prediction.py:128-141: float32 images in [-1,1] are mapped back to [0,255]; images of another
        size go through cv2.resize's bilinear interpolation (uint8: OpenCV's fixed-point form); mobilenet_v2.preprocess_input."""
        in_h, in_w = self.config["input_shape"][:2]
        out = []
        for b in range(0, input_images.shape[0], DEVICE_CHUNK):
            x = input_images[b:b + DEVICE_CHUNK]
            x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            x = x.to(self.device)
            if x.dtype == torch.float32:
                x = nets.pixel_map(x, 0)                                    # (x + 1) * 127.5
            elif x.dtype != torch.uint8:
                x = x.float()
            if tuple(x.shape[1:3]) != (in_h, in_w):
                x = nets.resize_images(x, in_h, in_w)
            x = ops.from_uint8(x) if x.dtype == torch.uint8 else nets.pixel_map(x, 1)   # astype(float32) / 127.5 - 1
            out.append(self.predict_device(x))
        if not out:
            return np.zeros((0, len(self.config["predicted_attributes"])), np.float32)
        return torch.cat(out, dim=0).cpu().numpy()
