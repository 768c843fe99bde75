"""Synthetic stand-in for NeuralRendererDataset (confignet/neural_renderer_dataset.py) with exactly the
attributes the training steps read: ``imgs`` (N,H,W,3) uint8, ``eye_masks`` (N,H,W) uint8,
``metadata_inputs`` (per facemodel parameter + "rotations"), ``metadata_input_distributions``.
Shapes follow SURVEY.md section 8d (test-dataset input dims); values are seeded random."""
import numpy as np
import torch
from . import netspec


class SyntheticDataset:
    def __init__(self, n_images=64, res=256, seed=0, facemodel_inputs=None):
        rng = np.random.RandomState(seed)
        fm = facemodel_inputs or netspec.default_facemodel_inputs()
        self.imgs = rng.randint(0, 256, (n_images, res, res, 3), dtype=np.uint8)
        self.eye_masks = (rng.rand(n_images, res, res) < 0.01).astype(np.uint8)
        self.metadata_inputs = {}
        for name, (n_in, _) in fm.items():
            if "embedding" in name or "params" in name:
                a = rng.standard_normal((n_images, n_in))
            else:
                a = rng.uniform(0, 1, (n_images, n_in))
            self.metadata_inputs[name] = a.astype(np.float32)
        rot = np.zeros((n_images, 3), np.float32)
        rot[:, 0] = np.pi * rng.uniform(-30, 30, n_images) / 180
        rot[:, 1] = np.pi * rng.uniform(-10, 10, n_images) / 180
        self.metadata_inputs["rotations"] = rot
        self.metadata_input_distributions = None

    def to_device(self, device):
        """Moves the image store and masks into HBM (the per-sample metadata stays on the host: it is
        indexed with the host RNG draw exactly as the reference does)."""
        self.imgs = torch.from_numpy(self.imgs).to(device)
        self.eye_masks = torch.from_numpy(self.eye_masks).to(device)
        return self
