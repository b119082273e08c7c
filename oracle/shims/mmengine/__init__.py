"""Minimal stand-in for mmengine so the read-only reference tree can be imported offline.
Test infrastructure only (oracle pinning); never imported by the product package."""
import torch.nn as nn


def print_log(msg, logger=None, level=None):
    pass


class Registry:
    def __init__(self, name, parent=None, locations=None, **kw):
        self.name = name
        self._mods = {}

    def register_module(self, name=None, module=None, force=False):
        def deco(cls):
            self._mods[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        typ = cfg.pop('type')
        return self._mods[typ](**cfg)

    def get(self, k):
        return self._mods.get(k)
