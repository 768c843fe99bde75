"""Deterministic stand-ins for the ConfigNet model and the attribute classifier, shared by
scripts/make_golden_metrics_from_reference.py (which drives the REFERENCE's ControllabilityMetrics / InceptionMetrics with
them) and tests/test_metrics_cpu.py (which drives the product's): the host logic of the metric classes is then compared
call for call, without any network in the loop."""
import numpy as np

FACEMODEL_INPUTS = {"beard_style_embedding": (9, 3), "blendshape_values": (62, 30), "head_hair_color": (3, 3),
                    "texture_embedding": (50, 30)}
ATTRIBUTES = ["Black_Hair", "Blond_Hair", "Brown_Hair", "Gray_Hair", "Mouth_Slightly_Open", "Smiling", "Narrow_Eyes",
              "Mustache", "No_Beard", "Goatee", "Sideburns", "Young", "Male"]


class FakeModel:
    """the part of the ConfigNet class surface the metric classes touch (metrics/metrics.py:30-102)"""

    def __init__(self, seed=5):
        self.config = {"facemodel_inputs": dict(FACEMODEL_INPUTS), "output_shape": (8, 8, 3)}
        self.rng = np.random.RandomState(seed)
        self.latent_dim = sum(v[1] for v in FACEMODEL_INPUTS.values())
        self.mix = [np.random.RandomState(100 + i).standard_normal((v[0], v[1])) for i, v in enumerate(FACEMODEL_INPUTS.values())]
        self.log = []

    def sample_facemodel_params(self, n_samples):
        return [self.rng.standard_normal((n_samples, v[0])) for v in FACEMODEL_INPUTS.values()]

    class _Enc:
        def __init__(self, outer):
            self.outer = outer

        def __call__(self, facemodel_params):
            self.outer.log.append(["synthetic_encoder"] + [np.asarray(p, np.float64).round(12).tolist() for p in facemodel_params])
            return np.hstack([np.tanh(np.asarray(p, np.float64) @ m) for p, m in zip(facemodel_params, self.outer.mix)])

        predict = __call__

    @property
    def synthetic_encoder(self):
        return FakeModel._Enc(self)

    def encode_images(self, imgs):
        flat = np.asarray(imgs, np.float64).reshape(len(imgs), -1)
        lat = np.sin(flat[:, :1] * 0.01 + np.arange(self.latent_dim)[None, :] * 0.1)
        return lat, np.cos(flat[:, :3] * 0.02)

    def fine_tune_on_img(self, img, n_iters):
        lat, rot = self.encode_images(img)
        return lat + 0.001 * n_iters, rot

    def generate_images(self, latents, rotations):
        self.log.append(["generate_images", np.asarray(latents, np.float64).round(12).tolist()])
        lat, rot = np.asarray(latents, np.float64), np.asarray(rotations, np.float64)
        v = np.concatenate([lat, rot], axis=1)
        img = 127.5 + 127.5 * np.sin(v @ np.random.RandomState(9).standard_normal((v.shape[1], 8 * 8 * 3)))
        return img.reshape(-1, 8, 8, 3).astype(np.uint8)


def fake_predict_attributes(images):
    flat = np.asarray(images, np.float64).reshape(len(images), -1) / 255.0
    w = np.random.RandomState(17).standard_normal((flat.shape[1], len(ATTRIBUTES)))
    return (1.0 / (1.0 + np.exp(-(flat - 0.5) @ w))).astype(np.float32)


class FakeDataset:
    def __init__(self, n=37, nfeat=24, seed=3):
        r = np.random.RandomState(seed)
        self.imgs = r.randint(0, 256, (n, 8, 8, 3)).astype(np.uint8)
        self.inception_features = r.standard_normal((n, nfeat)).astype(np.float32)


def fake_inception_features(images, nfeat=24):
    flat = np.asarray(images, np.float64).reshape(len(images), -1) / 255.0
    w = np.random.RandomState(23).standard_normal((flat.shape[1], nfeat))
    return ((flat - 0.5) @ w).astype(np.float32)
