"""Loaders for the committed golden fixtures (made by tools/make_golden.py from the unmodified reference)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_cache = {}


def small_cases():
    if "small" not in _cache:
        _cache["small"] = (np.load(os.path.join(GOLD, "small_cases.npz")), json.load(open(os.path.join(GOLD, "small_cases.json"))))
    return _cache["small"]


def small_case_names():
    return sorted(json.load(open(os.path.join(GOLD, "small_cases.json"))).keys())


def kodak():
    if "kodak" not in _cache:
        _cache["kodak"] = (np.load(os.path.join(GOLD, "kodak_gray.npz")), json.load(open(os.path.join(GOLD, "kodak_manifest.json"))))
    return _cache["kodak"]
