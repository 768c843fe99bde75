"""confignet_b200 - B200-native implementation of ConfigNet's G+D training hot path.

Public surface mirrors the reference package (confignet/__init__.py:3-14) for the components in scope."""
import json
import sys

from .confignet_first_stage import ConfigNetFirstStage, DEFAULT_CONFIG, merge_configs  # noqa: F401
from .confignet_second_stage import ConfigNet  # noqa: F401
from .latent_gan import LatentGAN  # noqa: F401
from .metrics.inception_distance import InceptionFeatureExtractor, compute_FID, compute_KID  # noqa: F401
from .metrics.metrics import InceptionMetrics, ControllabilityMetrics  # noqa: F401
from .metrics.celeba_attribute_prediction import CelebaAttributeClassifier  # noqa: F401


def load_confignet(model_path, **kw):
    """confignet_utils.py:14-21: dispatch on config["model_type"]."""
    with open(model_path, "r") as fp:
        config = json.load(fp)
    return getattr(sys.modules[__name__], config["model_type"]).load(model_path, **kw)
