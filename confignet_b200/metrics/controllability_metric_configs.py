"""The attribute configurations of the controllability metric (reference metrics/controllability_metric_configs.py): which
CelebA attribute a face-model parameter value is expected to drive, which attributes may move with it, and the two
parameter values that are compared.  The values are data (controllability_tables.json, written from the reference's
table by scripts/make_metric_tables_from_reference.py)."""
from collections import namedtuple
import json
import os

ControllableAttributeConfig = namedtuple(
    "ControllableAttributeConfig",
    "driven_attribute ignored_attributes facemodel_param_name facemodel_param_value facemodel_param_value_other")

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "controllability_tables.json")) as _fp:
    _TABLE = json.load(_fp)["configs"]


class ControllabilityMetricConfigs:
    @staticmethod
    def all_configs():
        """[(config name, ControllableAttributeConfig)] sorted by name - the order inspect.getmembers gives the reference"""
        return sorted(((name, ControllableAttributeConfig(**fields)) for name, fields in _TABLE), key=lambda t: t[0])


for _name, _fields in _TABLE:
    setattr(ControllabilityMetricConfigs, _name, ControllableAttributeConfig(**_fields))
