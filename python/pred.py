#!/usr/bin/env python
"""Prediction entry point — drop-in for the reference's python/pred.py (reference pred.py:52-74)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from wide_deep_b200.config import Config  # noqa: E402
from wide_deep_b200.dataset import input_fn  # noqa: E402
from wide_deep_b200.estimator import build_estimator  # noqa: E402

CONF = Config()
CONFIG = CONF.train
parser = argparse.ArgumentParser(description="Wide and Deep Model Prediction")
parser.add_argument("--model_dir", type=str, default=CONFIG["model_dir"])
parser.add_argument("--model_type", type=str, default=CONFIG["model_type"])
parser.add_argument("--data_dir", type=str, default="../data/pred")
parser.add_argument("--image_data_dir", type=str, default=None)
parser.add_argument("--batch_size", type=int, default=CONFIG["batch_size"])
parser.add_argument("--checkpoint_path", type=str, default=CONFIG["checkpoint_path"])

if __name__ == "__main__":
    FLAGS, unparsed = parser.parse_known_args()
    model_dir = os.path.join(FLAGS.model_dir, FLAGS.model_type)
    model = build_estimator(model_dir, FLAGS.model_type, config=CONF, max_batch=FLAGS.batch_size)
    preds = model.predict(input_fn=lambda: input_fn(FLAGS.data_dir, None, "pred", FLAGS.batch_size, config=CONF, plan=model.plan),
                          checkpoint_path=FLAGS.checkpoint_path)
    for pred_dict in preds:
        cid = int(pred_dict["class_ids"][0])
        print("Prediction is \"{}\" ({:.1f}%)".format(cid, 100 * float(pred_dict["probabilities"][cid])))
