#!/usr/bin/env python
"""Evaluation entry point — drop-in for the reference's python/eval.py (same flags; prints the sorted metric
dict, reference eval.py:56-83).  Evaluates what train.py wrote under <model_dir>/<model_type>."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from wide_deep_b200.config import Config  # noqa: E402
from wide_deep_b200.dataset import input_fn  # noqa: E402
from wide_deep_b200.estimator import build_estimator  # noqa: E402

CONF = Config()
CONFIG = CONF.train
parser = argparse.ArgumentParser(description="Evaluate Wide and Deep Model.")
parser.add_argument("--model_dir", type=str, default=CONFIG["model_dir"], help="Model checkpoint dir for evaluating.")
parser.add_argument("--model_type", type=str, default=CONFIG["model_type"], help="Valid model types: {'wide', 'deep', 'wide_deep'}.")
parser.add_argument("--test_data", type=str, default=CONFIG["test_data"], help="Evaluating data dir.")
parser.add_argument("--image_test_data", type=str, default=CONFIG.get("image_test_data"))
parser.add_argument("--batch_size", type=int, default=CONFIG["batch_size"], help="Number of examples per batch.")
parser.add_argument("--checkpoint_path", type=str, default=CONFIG["checkpoint_path"],
                    help="Path of a specific checkpoint to evaluate. If None, the latest checkpoint in model_dir is used.")


def main():
    print("Using wide_deep_b200 (CUDA sm_100a) in place of TensorFlow")
    print("Model type: {}".format(FLAGS.model_type))
    model_dir = os.path.join(FLAGS.model_dir, FLAGS.model_type)
    print("Model directory: {}".format(model_dir))
    model = build_estimator(model_dir, FLAGS.model_type, config=CONF, max_batch=FLAGS.batch_size)
    if not (FLAGS.checkpoint_path or model.latest_checkpoint()):
        raise ValueError("No model checkpoint found, please check the model dir.")
    print("INFO: " + "=" * 30 + " START TESTING" + "=" * 30)
    s_time = time.time()
    results = model.evaluate(input_fn=lambda: input_fn(FLAGS.test_data, None, "eval", FLAGS.batch_size, config=CONF, plan=model.plan),
                             checkpoint_path=FLAGS.checkpoint_path)
    print("INFO: " + "=" * 30 + "FINISH TESTING, TAKE {} mins".format(round((time.time() - s_time) / 60, 2)) + "=" * 30)
    print("-" * 80)
    for key in sorted(results):
        print("%s: %s" % (key, results[key]))


if __name__ == "__main__":
    FLAGS, unparsed = parser.parse_known_args()
    main()
