"""yacs-compatible configuration for the entry points (reference: config/defaults.py:1-97, read by
train_clip2.py:496-497 / test_clip2.py:421-422 through ``cfg.merge_from_file`` / ``cfg.merge_from_list``).

yacs is not installed in this image, so ``CfgNode`` below implements the subset the reference uses:
attribute access, nested nodes, ``merge_from_file`` (YAML), ``merge_from_list`` (``KEY.SUB value`` pairs with
yacs' literal parsing and type checks), ``clone``, ``freeze``/``defrost`` and the YAML-like ``str()`` that the
reference writes to ``config.yaml``.  Key names and defaults are the reference's.
"""
import ast
import copy

import yaml


class CfgNode(dict):
    _FROZEN = "__frozen__"

    def __init__(self, init=None):
        super().__init__()
        object.__setattr__(self, CfgNode._FROZEN, False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if object.__getattribute__(self, CfgNode._FROZEN):
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def _set_frozen(self, flag):
        object.__setattr__(self, CfgNode._FROZEN, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        return out

    @staticmethod
    def _decode(value):
        """yacs' _decode_cfg_value: strings are parsed as Python literals when possible ("(300, 375)" -> tuple)."""
        if isinstance(value, dict):
            return CfgNode(value)
        if not isinstance(value, str):
            return value
        try:
            return ast.literal_eval(value)
        except (ValueError, SyntaxError):
            return value

    @staticmethod
    def _coerce(new, old, key):
        """yacs' _check_and_coerce_cfg_value_type."""
        if old is None or type(new) is type(old):
            return new
        for a, b in ((tuple, list), (list, tuple)):
            if isinstance(new, a) and isinstance(old, b):
                return b(new)
        if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
            return float(new)
        if isinstance(old, str) and new is None:
            return new
        raise ValueError(f"Type mismatch ({type(old)} vs. {type(new)}) with values ({old} vs. {new}) for config key: {key}")

    def _merge(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError(f"Non-existent config key: {full}")
            v = self._decode(v)
            if isinstance(self[k], CfgNode):
                if not isinstance(v, dict):
                    raise ValueError(f"config key {full} is a node, got {type(v)}")
                self[k]._merge(v, path + [k])
            else:
                self[k] = self._coerce(v, self[k], full)

    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self._merge(loaded, [])

    def merge_from_list(self, cfg_list):
        cfg_list = list(cfg_list or [])
        if len(cfg_list) % 2:
            raise AssertionError(f"Override list has odd length: {cfg_list}; it must be a list of pairs")
        for full, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node, keys = self, full.split(".")
            for k in keys[:-1]:
                if k not in node:
                    raise AssertionError(f"Non-existent key: {full}")
                node = node[k]
            if keys[-1] not in node:
                raise AssertionError(f"Non-existent key: {full}")
            node[keys[-1]] = self._coerce(self._decode(v), node[keys[-1]], full)

    def __str__(self):
        def rec(node, indent):
            lines = []
            for k in sorted(node):
                v = node[k]
                if isinstance(v, CfgNode):
                    lines.append(" " * indent + f"{k}:")
                    lines.extend(rec(v, indent + 2))
                else:
                    lines.append(" " * indent + f"{k}: {v}")
            return lines
        return "\n".join(rec(self, 0))

    __repr__ = __str__


CN = CfgNode


def get_defaults():
    """Same keys and default values as the reference's config/defaults.py."""
    c = CN()
    c.DIR = "ckpt/ade20k-resnet50dilated-ppm_deepsup"
    c.DATASET = CN(dict(root_dataset="./data/", list_train="./data/training.odgt", list_val="./data/validation.odgt",
                        num_class=150, imgSizes=(300, 375, 450, 525, 600), imgMaxSize=1000, padding_constant=8,
                        segm_downsampling_rate=8, random_flip=True))
    c.MODEL = CN(dict(arch_encoder="resnet50dilated", arch_decoder="ppm_deepsup", weights_encoder="", weights_decoder="",
                      fc_dim=2048))
    c.TRAIN = CN(dict(batch_size_per_gpu=2, num_epoch=20, start_epoch=0, epoch_iters=5000, optim="SGD", lr_encoder=0.02,
                      lr_decoder=0.02, lr_pow=0.9, beta1=0.9, weight_decay=1e-4, deep_sup_scale=0.4, fix_bn=False,
                      workers=16, disp_iter=20, seed=304))
    c.VAL = CN(dict(batch_size=1, visualize=False, checkpoint="epoch_20.pth"))
    c.TEST = CN(dict(batch_size=1, checkpoint="epoch_20.pth", result="./"))
    return c


cfg = get_defaults()
