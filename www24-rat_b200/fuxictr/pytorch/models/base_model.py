"""BaseModel: the FuxiCTR model API of the reference (fuxictr/pytorch/models/base_model.py:31-301) on top of the
B200 engine.  What run_expid.py touches keeps its name, arguments, return values and side effects:
`model_class(feature_map, **params)`, `count_parameters()`, `fit_generator(train_gen, validation_data=..., **params)`,
`checkpoint`, `load_weights(path)`, `evaluate_generator(gen)` -> {"AUC":..,"logloss":..}, `predict_generator(gen)`.

Differences that are deliberate (DESIGN.md "Boundary"):
  * the training step is ONE fused device pipeline (`train_step`) instead of zero_grad/backward/clip/step; the
    regulariser, global-norm clip and dense Adam semantics are identical (kernels K5-K8);
  * there is no CPU path: constructing a model without an sm_100 device raises;
  * per-step host syncs of the reference (`loss.item()`, `.cpu().numpy()` per eval batch) are hoisted to one read
    per epoch / per evaluation.
"""
import logging
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from rat_native import call, current_stream
from rat_native.engine import EngineSpec, FeatureSpec, RatEngine

from ...metrics import evaluate_metrics
from ...utils import Monitor
from ..data_generator import DeviceBatch
from ..torch_utils import get_device, l2_lambda


class _FusedAdamHandle(object):
    """what `model.optimizer` is: the reference code only touches `param_groups[...]["lr"]`."""

    def __init__(self, model, lr):
        self._model = model
        self.param_groups = [{"lr": float(lr)}]

    def zero_grad(self):
        self._model._engine.store.G.zero_()

    def step(self):
        self._model._sync_lr()
        self._model._engine.optimizer_step()

    def state_dict(self):
        e = self._model._engine
        return {"step": float(e.opt_state[4]), "exp_avg": e.store.M, "exp_avg_sq": e.store.Vv,
                "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        e = self._model._engine
        e.store.M.copy_(sd["exp_avg"])
        e.store.Vv.copy_(sd["exp_avg_sq"])
        e.opt_state[4] = float(sd["step"])
        self.param_groups[0]["lr"] = float(sd["lr"])


class BaseModel(nn.Module):
    def __init__(self, feature_map, model_id="BaseModel", gpu=-1, monitor="AUC", save_best_only=True,
                 monitor_mode="max", patience=2, every_x_epochs=1, embedding_regularizer=None, net_regularizer=None,
                 reduce_lr_on_plateau=True, embedding_initializer="torch.nn.init.normal_(std=1e-4)",
                 retrieval_augmented=False, retrieval_configs=None, **kwargs):
        super(BaseModel, self).__init__()
        self.device = get_device(gpu)
        if self.device.type != "cuda":
            raise RuntimeError("the B200 RAT path has no CPU implementation: pass --gpu <index> on a machine with an "
                               "sm_100 device (gpu={}, cuda available={})".format(gpu, torch.cuda.is_available()))
        torch.cuda.set_device(self.device)
        self._monitor = Monitor(kv=monitor)
        self._monitor_mode = monitor_mode
        self._patience = patience
        self._every_x_epochs = every_x_epochs
        self._save_best_only = save_best_only
        self._embedding_regularizer = embedding_regularizer
        self._net_regularizer = net_regularizer
        self._reduce_lr_on_plateau = reduce_lr_on_plateau
        self._embedding_initializer = embedding_initializer
        self._retrieval_augmented = retrieval_augmented
        if self._retrieval_augmented:
            assert retrieval_configs is not None, \
                "retrieval-augmented mode requires a dataset with retrieval configurations"
            self._labelwise_retrieval = retrieval_configs["label_wise"]
        self._feature_map = feature_map
        self.model_id = model_id
        self.model_dir = os.path.join(kwargs["model_root"], feature_map.dataset_id)
        self.checkpoint = os.path.abspath(os.path.join(self.model_dir, self.model_id + ".model"))
        self._validation_metrics = kwargs["metrics"]
        self._verbose = kwargs["verbose"]
        self._seed = int(kwargs.get("seed", 2021))
        self._engine = None
        self._params = None

    # ------------------------------------------------------------------ construction helpers
    def _feature_specs(self):
        feats = []
        for name, spec in self._feature_map.feature_specs.items():
            typ = spec["type"]
            if typ == "numeric":
                raise NotImplementedError("numeric features are not used by any RAT configuration")
            for unsupported in ("pretrained_emb", "share_embedding", "embedding_dim"):
                if unsupported in spec:
                    raise NotImplementedError("feature option '{}' is outside the RAT hot path".format(unsupported))
            if typ == "sequence" and spec.get("encoder", None) != "MaskedSumPooling":
                raise NotImplementedError("sequence encoder {} (RAT configs use MaskedSumPooling)".format(spec.get("encoder")))
            feats.append(FeatureSpec(name, typ, int(spec["vocab_size"]), int(spec.get("max_len", 1)),
                                     spec.get("padding_idx", None)))
        return feats

    def _build_engine(self, spec: EngineSpec):
        self._engine = RatEngine(spec, str(self.device))
        self._params = OrderedDict((k, nn.Parameter(v, requires_grad=True)) for k, v in self._engine.p.items())

    def compile(self, optimizer, loss, lr):
        if not (isinstance(optimizer, str) and optimizer.lower() == "adam"):
            raise NotImplementedError("optimizer={} is not supported (fused Adam only).".format(optimizer))
        if loss not in ["bce", "binary_crossentropy", "binary_cross_entropy"]:
            raise NotImplementedError("loss={} is not supported.".format(loss))
        self.optimizer = _FusedAdamHandle(self, lr)
        self.loss_fn = torch.nn.functional.binary_cross_entropy
        self._engine.lr.fill_(float(lr))

    def _sync_lr(self):
        self._engine.lr.fill_(float(self.optimizer.param_groups[0]["lr"]))

    def reset_parameters(self):
        """reference base_model.py:101-123: embedding tables by `embedding_initializer` (padding row untouched = 0),
        nn.Linear xavier_normal / zero bias, LayerNorm & BatchNorm affine (1, 0); the label table keeps N(0,1)."""
        e, spec = self._engine, self._engine.spec
        pads = {f.name: f.pad for f in spec.features}
        with torch.no_grad():
            for k, v in e.p.items():
                if k == "label_embedding_layer.weight":
                    v.normal_(0.0, 1.0)
                elif "embedding_layer.embedding_layer.embedding_layer." in k:
                    feat = k.split(".")[-2]
                    pad = pads[feat]
                    if pad is not None and pad != v.shape[0] - 1:
                        # reference semantics for a padding_idx that is not the last row (embedding.py:96-100 +
                        # base_model.py:110-112): nn.Embedding's N(0,1) init with the padding row zeroed, then rows
                        # [0:-1] -- including the padding row -- re-drawn by the initializer; the last row keeps N(0,1)
                        v.normal_(0.0, 1.0)
                        v[pad].zero_()
                    else:
                        v.zero_()
                    w = v[0:-1, :] if pad is not None else v
                    if self._embedding_initializer is not None:
                        try:
                            eval(self._embedding_initializer.replace("(", "(w,", 1))
                        except Exception:
                            raise NotImplementedError("embedding_initializer={} is not supported."
                                                      .format(self._embedding_initializer))
                elif k.endswith("norm.weight") or self._is_bn(k, "weight"):
                    v.fill_(1.0)
                elif k.endswith(".bias"):
                    v.zero_()
                elif v.ndim == 2:
                    nn.init.xavier_normal_(v)
            if e.store.shard is not None:          # row-sharded tables: every rank initialises its own rows
                import re
                m = re.search(r"std\s*=\s*([0-9.eE+-]+)", str(self._embedding_initializer or ""))
                if self._embedding_initializer is not None and m is None:
                    raise NotImplementedError("embedding_initializer={} is not supported with shard_embeddings"
                                              .format(self._embedding_initializer))
                e.init_sharded_tables(float(m.group(1)) if m else 0.0, self._seed)
            for k, b in e.buffers.items():
                if k.endswith("running_mean"):
                    b.zero_()
                elif k.endswith("running_var"):
                    b.fill_(1.0)
                else:
                    b.zero_()

    def _is_bn(self, key, leaf):
        from rat_native.engine import dnn_layout
        layers, _ = dnn_layout(self._engine.spec)
        return any(key == "dnn.dnn.{}.{}".format(bn, leaf) for _, bn in layers if bn is not None)

    def model_to_device(self):
        pass    # parameters are created in HBM

    # ------------------------------------------------------------------ nn.Module surface over the flat buffers
    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        for k, v in self._params.items():
            yield (prefix + ("." if prefix else "") + k, v)

    def parameters(self, recurse=True):
        for _, v in self.named_parameters():
            yield v

    def _alias_keys(self):
        return {}

    def state_dict(self, *args, **kwargs):
        """reference key set.  With shard_embeddings the per-field tables are assembled from all ranks (collective:
        every rank must call it)."""
        sd = OrderedDict((k, v.detach()) for k, v in self._engine.p.items())
        if self._engine.store.shard is not None:
            sd.update(self._engine.gather_tables())
        for alias, target in self._alias_keys().items():
            sd[alias] = sd[target]
        for k, v in self._engine.buffers.items():
            sd[k] = v
        return sd

    def load_state_dict(self, state_dict, strict=True):
        own = self.state_dict()
        if self._engine.store.shard is not None:
            self._engine.scatter_tables(state_dict)
            own = OrderedDict((k, v) for k, v in own.items() if k in self._engine.p or k in self._engine.buffers)
            state_dict = OrderedDict((k, v) for k, v in state_dict.items()
                                     if "embedding_layer.embedding_layer.embedding_layer." not in k)
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError("load_state_dict: missing keys {} ; unexpected keys {}".format(missing[:8], unexpected[:8]))
        with torch.no_grad():
            for k, v in state_dict.items():
                if k in own and k not in self._alias_keys():
                    own[k].copy_(v.to(own[k].device).reshape(own[k].shape))
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def count_parameters(self, count_embedding=True):
        total = 0
        for name, p in self.named_parameters():
            if not count_embedding and "embedding" in name:
                continue
            if p.requires_grad:
                total += p.numel()
        if self._engine.store.shard is not None and count_embedding:      # sharded tables are not nn.Parameters
            sp = self._engine.spec
            total += sp.V * (sp.embedding_dim + (1 if sp.use_wide else 0))
        logging.info("Total number of parameters: {}.".format(total))
        return total

    def get_output_activation(self, task="binary_classification"):
        if task == "binary_classification":
            return nn.Sigmoid()
        raise NotImplementedError("task={} is not supported.".format(task))

    # ------------------------------------------------------------------ batches
    def _load_batch(self, inputs, training):
        e = self._engine
        if isinstance(inputs, DeviceBatch):
            g = inputs.gen
            B, T = inputs.size, g.K + 1
            ws = e._workspace(B, T, training)
            e.err_flag.zero_()
            call("rat_assemble_ids", g.q_ids, g.q_labels, inputs.rows, inputs.row0, g.pool_ids, g.pool_labels, g.nbr,
                 g.n_pool, ws["ids"], ws["labels"], ws["y_true"], B, T, g.L, e.err_flag, current_stream())
            return ws, B, T
        if self._retrieval_augmented:
            X, y, retrieved_values, retrieved_lens = inputs
            assert retrieved_lens.ndim == 1, "RIM does not support label-wise retrieval-enhanced training"
        else:
            raise NotImplementedError("RAT models are retrieval-augmented (retrieval_augmented: true)")
        assert X.ndim == 3, "retrieval augmented mode requires input_shape like [Bx(1+K)xF]"
        Xd = X.to(self.device, dtype=torch.float64, non_blocking=True)
        yd = y.to(self.device, dtype=torch.float64, non_blocking=True)
        self.batch_size = y.size(0)
        ws = e.load_wire(Xd, yd, training)
        return ws, X.shape[0], X.shape[1]

    def inputs_to_device(self, inputs):
        """kept for API parity (reference base_model.py:125-139)."""
        X, y, retrieved_values, retrieved_lens = inputs
        self.batch_size = y.size(0)
        return (X.to(self.device), y.float().unsqueeze(-1).to(self.device), retrieved_values.to(self.device),
                retrieved_lens.int().to(self.device))

    # ------------------------------------------------------------------ forward / loss / train step
    def forward(self, inputs):
        ws, B, T = self._load_batch(inputs, training=False)
        y_pred = self._engine.forward_ids(ws, B, T, training=False)
        return {"y_true": ws["y_true"].view(B, 1), "y_pred": y_pred.view(B, 1)}

    def add_loss(self, inputs, reduction="mean"):
        rd = self.forward(inputs)
        return self.loss_fn(rd["y_pred"], rd["y_true"], reduction=reduction)

    def add_regularization(self):
        e = self._engine
        lam_n, lam_e = float(e.spec.net_regularizer), float(e.spec.embedding_regularizer)
        W = e.store.W
        return 0.5 * lam_n * (W[:e.store.net_end] ** 2).sum() + 0.5 * lam_e * (W[e.store.net_end:] ** 2).sum()

    def get_total_loss(self, inputs):
        return self.add_loss(inputs) + self.add_regularization()

    def train_step(self, inputs):
        """zero_grad -> loss -> backward -> clip_grad_norm_ -> Adam.step of base_model.py:221-225 as one device
        pipeline.  Returns the 0-dim device tensor BCE(mean) + regularisation (the reference's `loss`)."""
        ws, B, T = self._load_batch(inputs, training=True)
        e = self._engine
        e.spec.max_gradient_norm = float(getattr(self, "_max_gradient_norm", 10.0))
        e.train_step_ids(ws, B, T)
        return ws["loss"][1] + e.opt_state[5]

    # ------------------------------------------------------------------ training loop (reference :144-230)
    def on_batch_end(self, batch, logs={}):
        self._total_batches += 1
        if (batch + 1) % self._every_x_batches == 0 or (batch + 1) % self._batches_per_epoch == 0:
            epoch = round(float(self._total_batches) / self._batches_per_epoch, 2)
            val_logs = self.evaluate_generator(self.valid_gen)
            self.checkpoint_and_earlystop(epoch, val_logs)
            self.train()
            logging.info("--- {}/{} batches finished ---".format(batch + 1, self._batches_per_epoch))

    def lr_decay(self, factor=0.1, min_lr=1e-6):
        for param_group in self.optimizer.param_groups:
            reduced_lr = max(param_group["lr"] * factor, min_lr)
            param_group["lr"] = reduced_lr
        self._sync_lr()
        return reduced_lr

    def checkpoint_and_earlystop(self, epoch, logs, min_delta=1e-6):
        monitor_value = self._monitor.get_value(logs)
        worse = (self._monitor_mode == "min" and monitor_value > self._best_metric - min_delta) or \
                (self._monitor_mode == "max" and monitor_value < self._best_metric + min_delta)
        if worse:
            self._stopping_steps += 1
            logging.info("Monitor({}) STOP: {:.6f} !".format(self._monitor_mode, monitor_value))
            if self._reduce_lr_on_plateau:
                current_lr = self.lr_decay()
                logging.info("Reduce learning rate on plateau: {:.6f}".format(current_lr))
        else:
            self._stopping_steps = 0
            self._best_metric = monitor_value
            if self._save_best_only:
                logging.info("Save best model: monitor({}): {:.6f}".format(self._monitor_mode, monitor_value))
                self.save_weights(self.checkpoint)
        if self._stopping_steps * self._every_x_epochs >= self._patience:
            self._stop_training = True
            logging.info("Early stopping at epoch={:g}".format(epoch))
        if not self._save_best_only:
            self.save_weights(self.checkpoint)

    def fit_generator(self, data_generator, epochs=1, validation_data=None, verbose=0, max_gradient_norm=10., **kwargs):
        self.valid_gen = validation_data
        self._max_gradient_norm = max_gradient_norm
        self._best_metric = np.inf if self._monitor_mode == "min" else -np.inf
        self._stopping_steps = 0
        self._total_batches = 0
        self._batches_per_epoch = len(data_generator)
        self._every_x_batches = int(np.ceil(self._every_x_epochs * self._batches_per_epoch))
        self._stop_training = False
        self._verbose = verbose
        logging.info("Start training: {} batches/epoch".format(self._batches_per_epoch))
        logging.info("************ Epoch=1 start ************")
        for epoch in range(epochs):
            epoch_loss = self.train_one_epoch(data_generator, epoch)
            logging.info("Train loss: {:.6f}".format(epoch_loss))
            if self._stop_training:
                break
            logging.info("************ Epoch={} end ************".format(epoch + 1))
        logging.info("Training finished.")

    def train_one_epoch(self, data_generator, epoch):
        self.train()
        loss_acc = torch.zeros((), device=self.device)      # accumulated on the device: one host read per epoch
        for batch_index, batch_data in enumerate(data_generator):
            loss_acc += self.train_step(batch_data)
            self.on_batch_end(batch_index)
            if self._stop_training:
                break
        self._engine.check_errors()
        return float(loss_acc.item()) / self._batches_per_epoch

    # ------------------------------------------------------------------ inference (reference :232-273)
    def _predict_all(self, data_generator, want_true):
        self.eval()
        preds, trues = [], []
        for batch_data in data_generator:
            ws, B, T = self._load_batch(batch_data, training=False)
            preds.append(self._engine.forward_ids(ws, B, T, training=False).clone())
            if want_true:
                trues.append(ws["y_true"].clone())
        self._engine.check_errors()
        y_pred = torch.cat(preds).double().cpu().numpy()          # single device->host read for the whole pass
        y_true = torch.cat(trues).double().cpu().numpy() if want_true else None
        return y_pred, y_true

    def evaluate_generator(self, data_generator):
        """reference :232-247.  AUC / logloss of the whole pass are computed on the device (rat_auc_logloss: exact
        tie-aware AUC from a radix sort of the scores, float64 logloss with the reference's 1e-7 clip) and only the two
        numbers travel to the host; any other metric falls back to evaluate_metrics on host copies."""
        metrics = list(self._validation_metrics)
        if not all(m in ("AUC", "logloss", "binary_crossentropy") for m in metrics):
            y_pred, y_true = self._predict_all(data_generator, True)
            return self.evaluate_metrics(y_true, y_pred, metrics)
        self.eval()
        preds, trues = [], []
        for batch_data in data_generator:
            ws, B, T = self._load_batch(batch_data, training=False)
            preds.append(self._engine.forward_ids(ws, B, T, training=False).clone())
            trues.append(ws["y_true"].clone())
        self._engine.check_errors()
        auc, ll = self._engine.auc_logloss(torch.cat(preds), torch.cat(trues))
        result = dict()
        for m in metrics:
            result[m] = auc if m == "AUC" else ll
        logging.info("[Metrics] " + " - ".join("{}: {:.6f}".format(k, v) for k, v in result.items()))
        return result

    def evaluate_metrics(self, y_true, y_pred, metrics):
        return evaluate_metrics(y_true, y_pred, metrics)

    def predict_generator(self, data_generator):
        return self._predict_all(data_generator, False)[0]

    # ------------------------------------------------------------------ checkpoints (reference :275-284)
    def save_weights(self, checkpoint):
        os.makedirs(os.path.dirname(checkpoint), exist_ok=True)
        torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in self.state_dict().items()), checkpoint)

    def load_weights(self, checkpoint):
        state_dict = torch.load(checkpoint, map_location="cpu")
        # strict like the reference (base_model.py:281): a checkpoint of another variant / use_wide / batch_norm setting fails loudly.
        # query_proj is a dead parameter that older checkpoints of this package did not store.
        for k, v in self.state_dict().items():
            if k.startswith("query_proj") and k not in state_dict:
                state_dict[k] = v.detach().cpu()
        self.load_state_dict(state_dict, strict=True)
        del state_dict
        torch.cuda.empty_cache()
