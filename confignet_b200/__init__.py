"""confignet_b200 - B200-native implementation of ConfigNet's G+D training hot path.

Public surface mirrors the reference package (confignet/__init__.py:3-14) for the components in scope."""
from .confignet_first_stage import ConfigNetFirstStage, DEFAULT_CONFIG, merge_configs  # noqa: F401
