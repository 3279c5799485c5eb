"""YAML configs -> namespaces, same contract as the reference's busca/option.py:6-39.

The ten shipped YAMLs (config/<Tracker>/<MOT17|MOT20>/*.yml) are read unchanged: four sections
``transformer / tracker / trainer / dataset``; the transformer namespace is attached to the tracker
and trainer namespaces, the dataset namespace to the trainer namespace.
"""
from __future__ import annotations

import copy
from types import SimpleNamespace

import yaml

_SECTIONS = ("tracker", "trainer", "transformer", "dataset")


def load_args_from_config(config_file):
    with open(config_file, "r") as fh:
        doc = yaml.safe_load(fh)
    ns = {name: SimpleNamespace(**doc[name]) for name in _SECTIONS}   # KeyError on a missing section, like the reference
    ns["tracker"].transformer = ns["transformer"]
    ns["trainer"].transformer = ns["transformer"]
    ns["trainer"].dataset = ns["dataset"]
    return ns["tracker"], ns["trainer"]


def merge_args(base_args, new_args, verbose=True):
    """CLI values override YAML values; a CLI ``None`` never overrides an existing key, but a key the
    YAML does not have is always added (option.py:23-39)."""
    merged = copy.deepcopy(base_args)
    for key, value in vars(new_args).items():
        present = hasattr(merged, key)
        if present and value is None:
            continue
        if verbose:
            if present:
                print("Overriding {} from {} to {}".format(key, getattr(merged, key), value), flush=True)
            else:
                print("Setting {} to {}".format(key, value), flush=True)
        setattr(merged, key, value)
    return merged
