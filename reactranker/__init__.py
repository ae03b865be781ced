"""Reference-compatible import paths (``reactranker.models.base_model.build_model`` ...) for the
B200-native implementation in ``reactranker_b200``.  Every sub-module is a thin alias."""
import importlib
import sys

_ALIASES = {
    "features": "reactranker_b200.features", "features.featurization": "reactranker_b200.features.featurization",
    "models": "reactranker_b200.models", "models.mpn": "reactranker_b200.models.mpn",
    "models.base_model": "reactranker_b200.models.base_model",
    "train": "reactranker_b200.train", "train.loss": "reactranker_b200.train.loss", "train.utils": "reactranker_b200.train.utils",
    "train.train_listwise": "reactranker_b200.train.train_listwise", "train.train_pairwise": "reactranker_b200.train.train_pairwise",
    "train.run_train_pairwise": "reactranker_b200.train.run_train_pairwise", "train.eval": "reactranker_b200.train.eval",
    "train.test_listwise": "reactranker_b200.train.test_listwise", "train.test_ranknet": "reactranker_b200.train.test_ranknet",
    "data": "reactranker_b200.data", "data.load_reactions": "reactranker_b200.data.load_reactions",
    "utils": "reactranker_b200.utils",
}


class _AliasFinder:
    @staticmethod
    def find_spec(name, path=None, target=None):
        if not name.startswith("reactranker."):
            return None
        real = _ALIASES.get(name[len("reactranker."):])
        if real is None:
            return None
        mod = importlib.import_module(real)
        sys.modules[name] = mod
        return importlib.util.spec_from_loader(name, loader=_Loader(mod))


class _Loader:
    def __init__(self, mod):
        self.mod = mod

    def create_module(self, spec):
        return self.mod

    def exec_module(self, module):
        pass


import importlib.util  # noqa: E402

sys.meta_path.insert(0, _AliasFinder)
