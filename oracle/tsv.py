"""ORACLE (test infrastructure): TSV line -> raw batch.

Restates ``_CsvDataset`` (reference python/lib/dataset.py:86-195): fields split on TAB only,
no quoting, a field that is empty or equals the NA token ``-`` takes its default ('' for string features,
0 for identity features and the label, 0.0 for continuous), multi-valued string fields split on ',' with
empty tokens dropped.  Pure-Python loops: meant for small fixtures.
"""
import numpy as np

from .columns import encode_tokens


def parse_lines(lines, schema, feature_conf, is_pred=False, multivalue=True, na_value="-"):
    """schema: {1-based idx: name} incl. label 'clk'.  Returns (raw_batch dict, labels float32[B] | None)."""
    names = [schema[k] for k in sorted(schema)]
    if is_pred:
        names = [n for n in names if n != "clk"]
    tokens = {f: [] for f, c in feature_conf.items() if c["type"] == "category" and c["transform"] != "identity"}
    ints = {f: [] for f, c in feature_conf.items() if c["type"] == "category" and c["transform"] == "identity"}
    flts = {f: [] for f, c in feature_conf.items() if c["type"] == "continuous"}
    labels = []
    for line in lines:
        line = line.rstrip("\n").rstrip("\r")
        fields = line.split("\t")
        if len(fields) != len(names):
            raise ValueError("Expect %d fields but have %d in record" % (len(names), len(fields)))
        for name, raw in zip(names, fields):
            na = (raw == "" or raw == na_value)
            if name == "clk":
                labels.append(0 if na else int(raw))
            elif name in tokens:
                s = "" if na else raw
                if multivalue:
                    tokens[name].append([t for t in s.split(",") if t != ""])
                else:
                    tokens[name].append([s] if s != "" else [])
            elif name in ints:
                ints[name].append(0 if na else int(raw))
            elif name in flts:
                flts[name].append(0.0 if na else float(raw))
    batch = {}
    for f, rows in tokens.items():
        batch[f] = encode_tokens(rows)
    for f, v in ints.items():
        batch[f] = np.asarray(v, dtype=np.int64)
    for f, v in flts.items():
        batch[f] = np.asarray(v, dtype=np.float32)
    lab = None if is_pred else (np.asarray(labels) == 1).astype(np.float32)   # dataset.py:158
    return batch, lab
