"""Import stub (test infrastructure only): bin/train_utils.py:9 imports OmegaConf for
load_config/save_config, which the hot path never calls."""


class OmegaConf:
    @staticmethod
    def load(*a, **k):
        raise NotImplementedError

    @staticmethod
    def merge(*a, **k):
        raise NotImplementedError
