"""Estimator-shaped shim: the object ``python/train.py`` / ``eval.py`` / ``pred.py`` drive.

Mirrors ``build_custom_estimator(model_dir, model_type)`` -> ``WideAndDeepClassifier`` (reference
python/lib/build_estimator.py:264-294, python/lib/joint.py:272-432): ``.train(input_fn)``, ``.evaluate(input_fn)
-> dict of the head's metric keys``, ``.predict(input_fn) -> iterator of dicts``; state lives under ``model_dir``
and every ``train`` call resumes from the latest checkpoint there (hence train.py's keep_train / rmtree logic).
The reference's eval.py / pred.py build the *canned* TF estimator whose variable names do not match what
train.py wrote (quirk Q9, pred.py:5-6); here all three entry points share this one class, so eval evaluates what
train trained.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from .config import Config
from .model import WideDeepModel
from .plan import compile_plan

CKPT_PREFIX = "model.ckpt-"


class WideAndDeepClassifier(object):
    def __init__(self, model_dir, model_type, config=None, device=0, max_batch=None, seed=None, tf_compat_pad=None,
                 gemm_engine="auto", shard_world=1, shard_rank=0, group=None):
        if model_type not in ("wide", "deep", "wide_deep"):
            raise ValueError("Invalid model type: {}, must be one of `wide`, `deep`, `wide_deep`".format(model_type))
        self.config = config or Config()
        self.model_dir, self.model_type = model_dir, model_type
        run = self.config.runconfig or {}
        self.seed = run.get("tf_random_seed", 123) if seed is None else seed
        self.keep_checkpoint_max = run.get("keep_checkpoint_max") or 5
        mb = max_batch or self.config.train["batch_size"]
        # quirk Q2: the reference pads multi-valued string fields with '' per batch and the padding takes
        # part in SparseCross.  The file-based entry points reproduce that by default (train.yaml key `tf_compat_pad`, default
        # true), so train.py / eval.py / pred.py produce the ids the reference produces on the same file.
        if tf_compat_pad is None:
            tf_compat_pad = bool(self.config.train.get("tf_compat_pad", True))
        self.tf_compat_pad = bool(tf_compat_pad)
        slack = 8 if self.config.train.get("multivalue") else 1
        if self.tf_compat_pad and self.config.train.get("multivalue"):
            slack *= 4                                   # '' padding multiplies the ids of crosses over multi-valued fields
        # multi-GPU (python/train.py under torchrun): rank `shard_rank` of `shard_world`; large tables are row-sharded over the
        # ranks, every rank trains on its shard of the input (wide_deep_b200/sharded.py)
        self.shard_world, self.shard_rank, self.group = int(shard_world), int(shard_rank), group
        self._trainer = None
        self.plan = compile_plan(self.config, model_type, mb, tf_compat_pad=tf_compat_pad, gemm_engine=gemm_engine,
                                 shard_world=self.shard_world, shard_rank=self.shard_rank, shard_slack=float(max(2, self.shard_world)),
                                 max_nnz=mb * len(self.config.read_feature_conf()) * 4 * slack + mb * 64,
                                 max_keys=mb * max(1, len(self.config.read_feature_conf())) * slack)
        self._model = None
        self.device = device

    # ------------------------------------------------------------------ checkpoints
    def latest_checkpoint(self):
        if not os.path.isdir(self.model_dir):
            return None
        steps = sorted(int(f[len(CKPT_PREFIX):-4]) for f in os.listdir(self.model_dir)
                       if f.startswith(CKPT_PREFIX) and f.endswith(".npz"))
        return os.path.join(self.model_dir, "%s%d.npz" % (CKPT_PREFIX, steps[-1])) if steps else None

    def _ensure_model(self, checkpoint_path=None, need_trained=False):
        if self._model is None:
            path = checkpoint_path or self.latest_checkpoint()
            if need_trained and not path:
                # tf.estimator raises the same way: "Could not find trained model in model_dir"
                raise ValueError("Could not find trained model in model_dir: {}.".format(self.model_dir))
            self._model = WideDeepModel(self.plan, device=self.device)
            if path:
                self.restore(path)
            else:
                self._model.init(self.seed)
            if self.shard_world > 1:
                from .sharded import ShardedTrainer
                self._trainer = ShardedTrainer(self._model, self.group)
        elif checkpoint_path:
            self.restore(checkpoint_path)
        return self._model

    def save(self):
        """Flat .npz of every variable (TensorFlow variable names) + optimizer slots; keeps the newest
        keep_checkpoint_max files (reference conf/train.yaml runconfig)."""
        m = self._model
        # multi-GPU: a collective — row-sharded tensors are gathered from all ranks; rank 0 alone writes
        get = self._trainer.get_tensor if self._trainer is not None else m.get_tensor
        blob = {"global_step": np.asarray(m.global_step)}
        for name in m.tensor_names():
            blob[name] = get(name)
            for s in range(m.n_slots(name)):
                blob["%s/slot%d" % (name, s + 1)] = get(name, slot=s + 1)
        if self.shard_rank != 0:
            return None
        os.makedirs(self.model_dir, exist_ok=True)
        path = os.path.join(self.model_dir, "%s%d.npz" % (CKPT_PREFIX, m.global_step))
        tmp = path + ".tmp.%d" % os.getpid()                 # written beside, then renamed: a crash never leaves a truncated
        with open(tmp, "wb") as fh:                          # checkpoint that latest_checkpoint() would pick up
            np.savez(fh, **blob)
            fh.flush()
            os.fsync(fh.fileno())
        os.replace(tmp, path)
        olds = sorted(int(f[len(CKPT_PREFIX):-4]) for f in os.listdir(self.model_dir) if f.startswith(CKPT_PREFIX) and f.endswith(".npz"))
        for st in olds[:-self.keep_checkpoint_max]:
            os.remove(os.path.join(self.model_dir, "%s%d.npz" % (CKPT_PREFIX, st)))
        return path

    def restore(self, path):
        m = self._model
        with np.load(path) as z:
            have = set(z.files)
            want = []
            for name in m.tensor_names():
                want.append((name, 0, tuple(self.plan.tensor_names[name][3])))
                for s in range(m.n_slots(name)):
                    want.append(("%s/slot%d" % (name, s + 1), s + 1, tuple(self.plan.tensor_names[name][3])))
            missing = [k for k, _, _ in want if k not in have]
            if missing or "global_step" not in have:
                raise ValueError("checkpoint {} does not match this model (feature conf, model_type or optimizers changed?): "
                                 "missing {} of {} tensors, e.g. {}".format(path, len(missing), len(want), missing[:3]))
            for key, slot, shape in want:
                if tuple(z[key].shape) != shape:
                    raise ValueError("checkpoint {}: tensor {} has shape {}, the model expects {}".format(path, key, tuple(z[key].shape), shape))
            m.global_step = int(z["global_step"])
            m.set_opt_step(m.global_step)
            for key, slot, _ in want:
                m.set_tensor(key if slot == 0 else key[:key.rindex("/slot")], z[key], slot=slot)

    # ------------------------------------------------------------------ estimator API
    def train(self, input_fn, hooks=None, steps=None, max_steps=None, saving_listeners=None):
        """Runs ``input_fn()`` to exhaustion (one pass = one epoch, like Estimator.train on a one-shot iterator),
        restoring the latest checkpoint first and saving one at the end."""
        m = self._ensure_model()
        n, t0, loss = 0, time.time(), float("nan")
        run = self.config.runconfig or {}
        log_every = run.get("log_step_count_steps") or 1000
        # checkpoint cadence of tf.estimator.RunConfig (reference conf/train.yaml:80-98): every `save_checkpoints_steps` global
        # steps, else every `save_checkpoints_secs` seconds (600 when neither is set); the step-count test is deterministic, so
        # in a multi-GPU job every rank takes the collective save() at the same step; the timer of rank 0 decides for all ranks
        ck_steps, ck_secs = run.get("save_checkpoints_steps"), run.get("save_checkpoints_secs")
        if ck_steps and ck_secs:
            raise ValueError("Can not provide both save_checkpoints_steps and save_checkpoints_secs.")      # RunConfig's message
        if not ck_steps and not ck_secs:
            ck_secs = 600
        last_save = time.time()
        # one batch of look-ahead, as the reference's input_fn prefetches (python/lib/dataset.py:181-184): while step i runs on the
        # GPU, batch i+1 is parsed and its host->device copy issued (wd_batch_prefetch_slot, two alternating slots)
        from .dataset import Prefetcher
        it = Prefetcher(input_fn(), depth=2)             # batches are parsed on a background thread, two ahead of the step
        cur = next(it, None)
        slot = 0
        if cur is not None:
            m.prefetch_slot(slot, cur)
        while cur is not None:
            if self._trainer is not None:
                self._trainer.step_slot(slot, want_loss=False)   # collective: every rank steps on its shard of the batch
            else:
                m.train_step_slot(slot, want_loss=False)     # enqueue step i ...
            nxt = next(it, None)                             # ... parse batch i+1 on the host while it runs ...
            if nxt is not None:
                m.prefetch_slot(1 - slot, nxt)               # ... start its copy on the upload stream ...
            loss = m.last_loss()                             # ... and only then wait for step i's loss
            n += 1
            if n % log_every == 0:
                print("INFO: global_step %d: loss = %.6g (%.1f steps/sec)" % (m.global_step, loss, n / (time.time() - t0)))
            if (steps and n >= steps) or (max_steps and m.global_step >= max_steps):
                break
            if nxt is not None and self._checkpoint_due(m.global_step, ck_steps, ck_secs, last_save):
                self.save()
                last_save = time.time()
            cur, slot = nxt, 1 - slot
        print("INFO: Loss for final step: %s." % loss)
        self.save()
        return self

    def _checkpoint_due(self, global_step, ck_steps, ck_secs, last_save):
        if ck_steps:
            return global_step % int(ck_steps) == 0
        due = (time.time() - last_save) >= float(ck_secs)
        if self.shard_world > 1:                              # save() is a collective: rank 0's clock decides for everyone
            import torch
            import torch.distributed as dist
            dev = torch.device("cuda", self.device) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
            flag = torch.tensor([1 if due else 0], dtype=torch.int32, device=dev)
            dist.broadcast(flag, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            due = bool(int(flag.item()))
        return due

    def evaluate(self, input_fn, steps=None, hooks=None, checkpoint_path=None, name=None):
        if self.shard_world > 1:
            raise ValueError("evaluate() on a multi-GPU (row-sharded) estimator: the reference's distributed mode trains only "
                             "(train.py:215-216); evaluate with a single-process estimator on the saved checkpoint")
        m = self._ensure_model(checkpoint_path, need_trained=True)
        m.eval_reset()
        n = 0
        for batch in input_fn():
            if batch.label is None:
                raise ValueError("evaluate needs labelled data (the reference's `pred`-mode test call, train.py:96-101, "
                                 "fails the same way inside TensorFlow)")
            m.eval_accumulate(batch)
            n += 1
            if steps and n >= steps:
                break
        out = m.eval_finish()
        out["global_step"] = m.global_step
        return out

    def predict(self, input_fn, predict_keys=None, hooks=None, checkpoint_path=None):
        m = self._ensure_model(checkpoint_path, need_trained=True)
        for batch in input_fn():
            logits, _ = m.forward(batch)
            p = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
            for i in range(batch.batch_size):
                yield {"logits": np.array([logits[i]], dtype=np.float32), "logistic": np.array([p[i]], dtype=np.float32),
                       "probabilities": np.array([1 - p[i], p[i]], dtype=np.float32),
                       "class_ids": np.array([int(logits[i] > 0)]), "classes": np.array([str(int(logits[i] > 0)).encode()])}

    def get_variable_names(self):
        return self.plan.tensor_names.keys()

    def get_variable_value(self, name):
        return self._ensure_model().get_tensor(name)


def build_custom_estimator(model_dir, model_type, **kw):
    """Same signature as the reference's build_custom_estimator (build_estimator.py:264-294)."""
    return WideAndDeepClassifier(model_dir, model_type, **kw)


# eval.py / pred.py of the reference call build_estimator (the canned TF classes); here it is the same object
build_estimator = build_custom_estimator
