"""Training-time metrics of the reference (confignet/metrics/) on the B200 kernels: KID / FID on
InceptionV3 features (inception_distance.py), the CelebA attribute classifier on MobileNetV2 (celeba_attribute_prediction.py)
and the controllability metrics built on it (metrics.py).  Same module, class and method names as the reference package."""
from .inception_distance import InceptionFeatureExtractor, compute_FID, compute_KID      # noqa: F401
from .celeba_attribute_prediction import CelebaAttributeClassifier                       # noqa: F401
from .controllability_metric_configs import ControllabilityMetricConfigs                 # noqa: F401
from .metrics import ControllabilityMetrics, InceptionMetrics                            # noqa: F401
