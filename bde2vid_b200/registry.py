"""mmengine-lite: just enough ``Registry`` / ``Config`` for the reference driver's loading path
(eval_models_seq.py:52-60: ``Config.fromstring(checkpoint['meta']['cfg'], '.py').model`` ->
``MODELS.build(model_cfg)``) to work without mmengine installed."""


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError("%s is already registered in %s" % (key, self.name))
            self._modules[key] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        cls = self._modules.get(typ)
        if cls is None:
            raise KeyError("%s is not registered in %s" % (typ, self.name))
        return cls(**cfg)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v


class Config(ConfigDict):
    @staticmethod
    def fromstring(text, file_format=".py"):
        if file_format not in (".py", "py"):
            raise NotImplementedError("only python-syntax configs are supported")
        ns = {}
        exec(compile(text, "<bde2vid-cfg>", "exec"), ns)  # noqa: S102 - same trust model as mmengine's Config
        return Config({k: v for k, v in ns.items() if not k.startswith("__")})


MODELS = Registry("model")
ACTIVATION = Registry("activation")
