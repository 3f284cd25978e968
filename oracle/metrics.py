"""ORACLE (test infrastructure): eval metric dict of the binary-logistic head.

Restates the metric ops the reference's head adds in EVAL mode (reference python/lib/joint.py:402-406 ->
tensorflow ``_binary_logistic_head_with_sigmoid_cross_entropy_loss``): streaming sums
over eval batches, ``auc`` / ``auc_precision_recall`` = ``tf.metrics.auc`` with 200 thresholds, trapezoidal.
"""
import numpy as np

EPS = 1e-7
NUM_THRESHOLDS = 200


def thresholds(n=NUM_THRESHOLDS):
    return np.array([0.0 - EPS] + [(i + 1) * 1.0 / (n - 1) for i in range(n - 2)] + [1.0 + EPS])


class EvalAccumulator(object):
    def __init__(self):
        self.thr = thresholds()
        self.tp = np.zeros(len(self.thr)); self.fp = np.zeros(len(self.thr))
        self.tn = np.zeros(len(self.thr)); self.fn = np.zeros(len(self.thr))
        self.sw = self.sloss = self.slabel = self.spred = self.scorrect = 0.0
        self.tp5 = self.fp5 = self.fn5 = 0.0
        self.batch_losses = []

    def update(self, logits, labels, weights=None):
        x = logits.astype(np.float64).reshape(-1)
        z = labels.astype(np.float64).reshape(-1)
        w = np.ones_like(x) if weights is None else weights.astype(np.float64).reshape(-1)
        p = (1.0 / (1.0 + np.exp(-x))).astype(np.float32).astype(np.float64)   # predictions are fp32 in TF
        loss = np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))
        self.batch_losses.append(float((w * loss).sum()))
        self.sw += w.sum(); self.sloss += (w * loss).sum()
        self.slabel += (w * z).sum(); self.spred += (w * p).sum()
        cls = (x > 0).astype(np.float64)
        self.scorrect += (w * (cls == z)).sum()
        self.tp5 += (w * cls * z).sum(); self.fp5 += (w * cls * (1 - z)).sum(); self.fn5 += (w * (1 - cls) * z).sum()
        pos = z > 0.5
        for i, t in enumerate(self.thr):
            pr = p > t
            self.tp[i] += w[pr & pos].sum(); self.fp[i] += w[pr & ~pos].sum()
            self.fn[i] += w[~pr & pos].sum(); self.tn[i] += w[~pr & ~pos].sum()

    def result(self):
        tp, fp, tn, fn = self.tp, self.fp, self.tn, self.fn
        rec = (tp + EPS) / (tp + fn + EPS)
        fpr = fp / (fp + tn + EPS)
        prec = (tp + EPS) / (tp + fp + EPS)
        auc = float(((fpr[:-1] - fpr[1:]) * (rec[:-1] + rec[1:]) / 2.0).sum())
        aupr = float(((rec[:-1] - rec[1:]) * (prec[:-1] + prec[1:]) / 2.0).sum())
        lm = self.slabel / self.sw
        div = lambda a, b: float(a / b) if b > 0 else 0.0
        return {
            "accuracy": float(self.scorrect / self.sw),
            "accuracy_baseline": float(max(lm, 1 - lm)),
            "auc": auc,
            "auc_precision_recall": aupr,
            "average_loss": float(self.sloss / self.sw),
            "label/mean": float(lm),
            "loss": float(np.mean(self.batch_losses)),
            "precision": div(self.tp5, self.tp5 + self.fp5),
            "prediction/mean": float(self.spred / self.sw),
            "recall": div(self.tp5, self.tp5 + self.fn5),
        }
