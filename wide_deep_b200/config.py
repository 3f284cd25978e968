"""YAML configuration surface (conf/{schema,feature,cross_feature,model,train}.yaml).

Host-side mirror of the reference's ``Config`` object (reference python/lib/read_conf.py:21-279):
same constructor arguments, same accessor names (``read_schema``, ``read_feature_conf``,
``read_cross_feature_conf``, ``get_feature_name``, ``.train/.model/.distribution/.runconfig``) and the
same error classes for malformed files (ValueError / TypeError / AssertionError), so the rest of the
host code — and a user of the reference — can switch without touching their conf directory.

Differences that are deliberate:
  * files are parsed once per ``Config`` instance and cached (the reference re-reads the YAML on every
    property access and at every module import, read_conf.py:41-47,234-257);
  * ``yaml.safe_load`` (PyYAML >= 6 refuses ``yaml.load`` without a Loader);
  * cross ``hash_bucket_size`` is returned as an ``int`` (the reference yields ``100.0`` for ``0.1``,
    read_conf.py:151, which TensorFlow's op attr would reject).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import yaml

_DEFAULT_CONF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conf")


def _load(path):
    with open(path) as fh:
        return yaml.safe_load(fh)


class Config(object):
    """Parsed view of a conf directory.  ``conf_dir`` defaults to ``<repo>/conf``."""

    def __init__(self,
                 schema_conf_file="schema.yaml",
                 data_process_conf_file="data_process.yaml",
                 feature_conf_file="feature.yaml",
                 cross_feature_conf_file="cross_feature.yaml",
                 model_conf_file="model.yaml",
                 train_conf_file="train.yaml",
                 serving_conf_file="serving.yaml",
                 conf_dir=None):
        base = conf_dir or os.environ.get("WD_CONF_DIR") or _DEFAULT_CONF_DIR
        self.conf_dir = base
        self._paths = dict(
            schema=os.path.join(base, schema_conf_file),
            feature=os.path.join(base, feature_conf_file),
            cross=os.path.join(base, cross_feature_conf_file),
            model=os.path.join(base, model_conf_file),
            train=os.path.join(base, train_conf_file),
            serving=os.path.join(base, serving_conf_file),
            data_process=os.path.join(base, data_process_conf_file),
        )
        self._cache = {}

    # ------------------------------------------------------------------ schema
    def read_schema(self):
        """{1-based column index: lower-cased column name} (reference read_conf.py:41-43)."""
        if "schema" not in self._cache:
            raw = _load(self._paths["schema"])
            self._cache["schema"] = OrderedDict((k, str(raw[k]).lower()) for k in sorted(raw))
        return self._cache["schema"]

    # ----------------------------------------------------------------- feature
    @staticmethod
    def _check_feature_conf(feature, valid_names, type=None, transform=None, parameter=None, **_):
        # same checks, same exception classes as reference read_conf.py:49-112
        if type is None:
            raise ValueError("Type are required in feature conf, found empty value for feature `{}`".format(feature))
        if feature not in valid_names:
            raise ValueError("Invalid feature name `{}` in feature conf, must be consistent with schema conf".format(feature))
        assert type in ("category", "continuous"), (
            "Invalid type `{}` for feature `{}` in feature conf, must be 'category' or 'continuous'".format(type, feature))
        if type == "category":
            assert transform in ("hash_bucket", "identity", "vocab"), (
                "Invalid transform `{}` for feature `{}` in feature conf, "
                "must be one of `hash_bucket`, `vocab`, `identity`.".format(transform, feature))
            if transform in ("hash_bucket", "identity"):
                if not isinstance(parameter, int) or isinstance(parameter, bool):
                    raise TypeError("Invalid parameter `{}` for feature `{}` in feature conf, "
                                    "{} parameter must be an integer.".format(parameter, feature, transform))
            elif not isinstance(parameter, (tuple, list)):
                raise TypeError("Invalid parameter `{}` for feature `{}` in feature conf, "
                                "vocab parameter must be a list.".format(parameter, feature))
            return
        norm, bounds = parameter["normalization"], parameter["boundaries"]
        if transform:
            assert transform in ("min_max", "log", "standard"), (
                "Invalid transform `{}` for feature `{}` in feature conf, continuous feature transform "
                "must be `min_max` or `log` or `standard`.".format(transform, feature))
            # (the reference's `if trans == 'min_max' or 'standard'` is always true: the 2-element
            #  check applies to `log` too, read_conf.py:82)
            if not isinstance(norm, (list, tuple)) or len(norm) != 2:
                raise TypeError("Invalid normalization parameter `{}` for feature `{}` in feature conf, "
                                "must be 2 elements list for `min_max` or `standard` scaler.".format(norm, feature))
            if transform == "min_max":
                lo, hi = norm
                if not isinstance(lo, (float, int)) or not isinstance(hi, (float, int)):
                    raise TypeError("Invalid normalization parameter `{}` for feature `{}` in feature conf, "
                                    "list elements must be int or float.".format(norm, feature))
                assert lo < hi, ("Invalid normalization parameter `{}` for feature `{}` in feature conf, "
                                 "[min, max] list elements must be min<max".format(norm, feature))
            elif transform == "standard":
                mean, std = norm
                if not isinstance(mean, (float, int)):
                    raise TypeError("Invalid normalization parameter `{}` for feature `{}` in feature conf, "
                                    "parameter mean must be int or float.".format(mean, feature))
                if not isinstance(std, (float, int)) or std <= 0:
                    raise TypeError("Invalid normalization parameter `{}` for feature `{}` in feature conf, "
                                    "parameter std must be a positive number.".format(std, feature))
        if bounds:
            if not isinstance(bounds, (tuple, list)):
                raise TypeError("Invalid parameter `{}` for feature `{}` in feature conf, "
                                "discretize parameter must be a list.".format(bounds, feature))
            for v in bounds:
                assert isinstance(v, (int, float)), (
                    "Invalid parameter `{}` for feature `{}` in feature conf, "
                    "discretize parameter element must be integer or float.".format(bounds, feature))

    def read_feature_conf(self):
        """Ordered {feature: {type, transform, parameter}} (reference read_conf.py:135-141)."""
        if "feature" not in self._cache:
            raw = _load(self._paths["feature"]) or {}
            valid = set(self.read_schema().values())
            for name, conf in raw.items():
                self._check_feature_conf(name.lower(), valid, **conf)
            self._cache["feature"] = OrderedDict(raw)
        return self._cache["feature"]

    # ------------------------------------------------------------------- cross
    @staticmethod
    def _check_cross_feature_conf(features, feature_conf, hash_bucket_size=None, is_deep=None, **_):
        names = [f.strip() for f in features.split("&")]
        assert len(names) > 1, ("Invalid cross feature name `{}` in cross feature conf,"
                                "at least 2 features".format(features))
        for f in names:
            if f not in feature_conf:
                raise ValueError("Invalid cross feature name `{}` in cross feature conf, "
                                 "must be consistent with feature conf".format(features))
            if feature_conf[f]["type"] == "continuous":
                assert feature_conf[f]["parameter"]["boundaries"] is not None, (
                    "Continuous feature must be set bounaries to be bucketized in feature conf as cross feature")
        if hash_bucket_size:
            assert isinstance(hash_bucket_size, (int, float)), (
                "Invalid hash_bucket_size `{}` for features `{}` in cross feature conf, "
                "expected int or float".format(hash_bucket_size, features))
        if is_deep:
            assert is_deep in (0, 1), ("Invalid is_deep `{}` for features `{}`, expected 0 or 1.".format(is_deep, features))

    def read_cross_feature_conf(self):
        """[(feature list, bucket count, is_deep)]; size = 1000*conf or 10000, is_deep defaults to 1
        (reference read_conf.py:143-154)."""
        if "cross" not in self._cache:
            raw = _load(self._paths["cross"]) or {}
            fconf = self.read_feature_conf()
            out = []
            for key, conf in raw.items():
                conf = conf or {}
                conf = {"hash_bucket_size": conf.get("hash_bucket_size"), "is_deep": conf.get("is_deep")}
                self._check_cross_feature_conf(key, fconf, **conf)
                names = [f.strip() for f in key.split("&")]
                size = int(round(1000 * conf["hash_bucket_size"])) if conf["hash_bucket_size"] else 10000
                deep = conf["is_deep"] if conf["is_deep"] is not None else 1
                out.append((names, size, int(deep)))
            self._cache["cross"] = out
        return self._cache["cross"]

    # ------------------------------------------------------------ model / train
    @staticmethod
    def _check_numeric(key, value):
        if isinstance(value, bool) or not isinstance(value, (int, float)):
            raise ValueError("Numeric type is required for key `{}`, found `{}`.".format(key, value))

    @staticmethod
    def _check_string(key, value):
        if not isinstance(value, str):
            raise ValueError("String type is required for key `{}`, found `{}`.".format(key, value))

    @staticmethod
    def _check_bool(key, value):
        if value not in (True, False, 1, 0):
            raise ValueError("Bool type is required for key `{}`, found `{}`.".format(key, value))

    @staticmethod
    def _check_list(key, value):
        if not isinstance(value, (list, tuple)):
            raise ValueError("List type is required for key `{}`, found `{}`.".format(key, value))

    @staticmethod
    def _check_required(key, value):
        if value is None:
            raise ValueError("Required type for key `{}`, found None.".format(key))

    def _read_model_conf(self):
        if "model" in self._cache:
            return self._cache["model"]
        # reference read_conf.py:181-211 (its `req_str_keys` list glues the last two names together by
        # a missing comma, so only the first three are really enforced; we enforce the intended four)
        req_str = ("linear_optimizer", "dnn_optimizer", "dnn_activation_function")
        opt_num = ("linear_initial_learning_rate", "linear_decay_rate", "dnn_initial_learning_rate",
                   "dnn_decay_rate", "dnn_l1", "dnn_l2")
        opt_bool = ("dnn_batch_normalization", "cnn_use_flag")
        conf = _load(self._paths["model"])
        for k, v in conf.items():
            if k in req_str:
                self._check_required(k, v)
                self._check_string(k, v)
            elif k in opt_num:
                if v:
                    self._check_numeric(k, v)
            elif k in opt_bool:
                if v:
                    self._check_bool(k, v)
            elif k == "dnn_hidden_units":
                self._check_required(k, v)
                self._check_list(k, v)
        self._check_required("dnn_connected_mode", conf.get("dnn_connected_mode"))
        self._cache["model"] = conf
        return conf

    def _read_train_conf(self):
        if "train" in self._cache:
            return self._cache["train"]
        req_str = ("model_dir", "model_type", "train_data", "test_data")
        req_num = ("train_epochs", "epochs_per_eval", "batch_size", "num_examples")
        opt_num = ("pos_sample_loss_weight", "neg_sample_loss_weight", "num_parallel_calls")
        req_bool = ("keep_train", "multivalue", "dynamic_train")
        conf = _load(self._paths["train"])
        for k, v in conf["train"].items():
            if k in req_str:
                self._check_required(k, v)
                self._check_string(k, v)
            elif k in req_num:
                self._check_required(k, v)
                self._check_numeric(k, v)
            elif k in opt_num:
                if v:
                    self._check_numeric(k, v)
            elif k in req_bool:
                self._check_required(k, v)
                self._check_bool(k, v)
        self._cache["train"] = conf
        return conf

    @property
    def config(self):
        return self._read_train_conf()

    @property
    def train(self):
        return self._read_train_conf()["train"]

    @property
    def distribution(self):
        return self._read_train_conf().get("distribution", {"is_distribution": 0})

    @property
    def runconfig(self):
        return self._read_train_conf().get("runconfig", {})

    @property
    def model(self):
        return self._read_model_conf()

    @property
    def serving(self):
        return _load(self._paths["serving"]) if os.path.exists(self._paths["serving"]) else {}

    def read_data_process_conf(self):
        return _load(self._paths["data_process"]) if os.path.exists(self._paths["data_process"]) else {}

    # ------------------------------------------------------------- name lists
    def get_feature_name(self, feature_type="all"):
        """Feature names by kind (reference read_conf.py:259-279).  'all' is every schema column but the
        label, in schema order."""
        fconf = self.read_feature_conf()
        schema_names = [v for v in self.read_schema().values() if v != "clk"]
        if feature_type == "all":
            return schema_names
        if feature_type == "used":
            return list(fconf.keys())
        if feature_type == "unused":
            return [n for n in schema_names if n not in fconf]
        if feature_type in ("category", "continuous"):
            return [f for f, c in fconf.items() if c["type"] == feature_type]
        raise ValueError("Invalid parameter, must be one of 'all', 'used', 'category, 'continuous")
