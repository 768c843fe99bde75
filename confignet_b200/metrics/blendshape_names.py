"""Blendshape order of the face model (reference metrics/blendshape_names.py); the list is data, kept in
controllability_tables.json (scripts/make_metric_tables_from_reference.py)."""
import json
import os

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "controllability_tables.json")) as _fp:
    blendshape_names = json.load(_fp)["blendshape_names"]
