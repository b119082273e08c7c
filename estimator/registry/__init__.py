"""Minimal `MODELS` registry with the reference's call surface (`estimator/registry/registry.py:7`,
`estimator/models/builder.py:6-8`): `MODELS.build(dict(type='PatchFusion', config=...))`."""


class _Registry:
    def __init__(self, name):
        self.name, self._mods = name, {}

    def register_module(self, name=None):
        def deco(cls):
            self._mods[name or cls.__name__] = cls
            return cls
        return deco

    def get(self, k):
        return self._mods.get(k)

    def build(self, cfg):
        cfg = dict(cfg)
        typ = cfg.pop('type')
        if typ not in self._mods:
            raise KeyError('%s is not in the %s registry (only the hot-path model is built)' % (typ, self.name))
        return self._mods[typ](**cfg)


MODELS = _Registry('model')
