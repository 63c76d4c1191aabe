"""RAT_m0 / RAT_m1 / RAT_m2 / RAT_m3 model classes with the reference constructor signatures
(fuxictr/pytorch/models/RAT_m{0,1,2,3}.py:29-56) and state_dict key sets."""
from rat_native.engine import EngineSpec

from ..torch_utils import l2_lambda
from .base_model import BaseModel


class _RATBase(BaseModel):
    _variant = "RAT_m2"

    def __init__(self, feature_map, model_id=None, gpu=-1, task="binary_classification", learning_rate=1e-3,
                 embedding_dim=10, dnn_hidden_units=[64, 64, 64], dnn_activations="ReLU", attention_layers=2,
                 num_heads=1, attention_dim=8, net_dropout=0, batch_norm=False, layer_norm=False, use_scale=False,
                 use_wide=False, use_residual=True, embedding_regularizer=None, net_regularizer=None, depth=4, heads=4,
                 pool="cls", dim_head=10, dropout=0., emb_dropout=0., scale_dim=4, **kwargs):
        super().__init__(feature_map, model_id=model_id or self._variant, gpu=gpu,
                         embedding_regularizer=embedding_regularizer, net_regularizer=net_regularizer, **kwargs)
        if str(dnn_activations).lower() != "relu":
            raise NotImplementedError("dnn_activations={} (RAT configs use relu)".format(dnn_activations))
        if dropout and float(dropout) > 0:
            raise NotImplementedError("attention dropout > 0 is not used by any RAT configuration")
        if num_heads == 1 and dim_head == embedding_dim:
            raise NotImplementedError("identity out-projection (heads==1 and dim_head==dim) is not supported")
        _ = kwargs["retrieval_configs"]["topK"]           # required key, unused at run time like the reference
        # arithmetic of the projections / DNN GEMMs (not a reference option): config key `precision`, default fp16
        # tensor-core operands with fp32 accumulation; "fp32" reproduces the reference arithmetic exactly
        if kwargs.get("precision") is not None:
            import logging
            from rat_native.engine import set_precision
            set_precision(str(kwargs["precision"]))
            logging.info("RAT precision mode: {}".format(kwargs["precision"]))
        spec = EngineSpec(
            features=self._feature_specs(), model=self._variant, embedding_dim=int(embedding_dim),
            num_heads=int(num_heads), dim_head=int(dim_head), scale_dim=int(scale_dim), depth=int(depth),
            dnn_hidden_units=tuple(dnn_hidden_units or ()), batch_norm=bool(batch_norm), use_wide=bool(use_wide),
            emb_dropout=float(emb_dropout), net_dropout=float(net_dropout),
            embedding_regularizer=l2_lambda(embedding_regularizer), net_regularizer=l2_lambda(net_regularizer),
            learning_rate=float(learning_rate), seed=self._seed,
            shard_tables=bool(kwargs.get("shard_embeddings", False)))
        self._build_engine(spec)
        self.output_activation = self.get_output_activation(task)
        self.compile(kwargs["optimizer"], loss=kwargs["loss"], lr=learning_rate)
        self.reset_parameters()
        self.model_to_device()


class RAT_m2(_RATBase):
    """default RAT: cascaded intra-sample / cross-sample attention blocks (RAT_m2.py:204-259)."""
    _variant = "RAT_m2"


class RAT_m0(_RATBase):
    """RAT_JM: one joint Transformer over all (1+K)(F+1) tokens (RAT_m0.py:123-127)."""
    _variant = "RAT_m0"


class RAT_m1(_RATBase):
    """RAT_CE: intra Transformer per row, then cross Transformer over the pooled row tokens (RAT_m1.py:123-129)."""
    _variant = "RAT_m1"


class RAT_m3(_RATBase):
    """RAT_PA: parallel intra || cross attention sharing W_q (RAT_m3.py:164-242)."""
    _variant = "RAT_m3"

    def _alias_keys(self):
        out = {}
        for l in range(self._engine.spec.depth):
            pre = "encoder.encoder.{}.".format(l)
            out[pre + "intra_attention.fn.W_q.weight"] = pre + "W_q.weight"
            out[pre + "intra_attention.fn.W_k.weight"] = pre + "W_k_s.weight"
            out[pre + "intra_attention.fn.W_v.weight"] = pre + "W_v_s.weight"
            out[pre + "cross_attention.fn.W_q.weight"] = pre + "W_q.weight"
            out[pre + "cross_attention.fn.W_k.weight"] = pre + "W_k_t.weight"
            out[pre + "cross_attention.fn.W_v.weight"] = pre + "W_v_t.weight"
        return out
