"""Layered config loading with the reference's semantics (``annotator/utils/load.py:16-84``):
the first file is the base, every later file overlays it, and a dotted key
``a.b.c: v`` creates/overwrites nested entries."""
import json
import os
import pickle

import yaml


def load_config(path):
    if isinstance(path, str):
        return load_config([path])
    assert isinstance(path, (tuple, list)) and path
    configs = [_load_single(p) for p in path]
    config = configs[0]
    for extra in configs[1:]:
        config = _overlay(config, extra)
    return config


def _overlay(base, extra):
    def put(target, dotted, value):
        head, _, rest = dotted.partition('.')
        if not rest:
            target[head] = value
        else:
            put(target.setdefault(head, dict()), rest, value)
    for key, val in extra.items():
        put(base, key, val)
    return base


def _load_single(path):
    ext = os.path.splitext(path)[1][1:]
    if ext == 'json':
        with open(path) as f:
            return json.load(f)
    if ext == 'yaml':
        with open(path) as f:
            return yaml.safe_load(f)
    if ext == 'pickle':
        with open(path, 'rb') as f:
            return pickle.load(f)
    raise NotImplementedError(f'Unexpected extension {ext}')
