"""Host-side helpers with the reference's names (utils.py): config loading and checkpoint discovery."""
from __future__ import annotations

import os

import yaml


def get_config(config):
    """YAML -> plain dict (reference utils.py:183-185)."""
    with open(config, 'r') as stream:
        return yaml.safe_load(stream)


def get_model_list(dirname, key):
    """Lexicographically last ``*.pt`` file in ``dirname`` whose name contains ``key`` (utils.py:336-348)."""
    if not os.path.isdir(dirname):
        return None
    found = sorted(os.path.join(dirname, f) for f in os.listdir(dirname)
                   if '.pt' in f and key in f and os.path.isfile(os.path.join(dirname, f)))
    return found[-1] if found else None
