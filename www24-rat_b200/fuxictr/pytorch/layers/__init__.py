"""The reference's layer zoo (EmbeddingLayer, LR_Layer, MLP_Layer, ...) has no module objects here: on the B200
path the parameters of those layers live in the engine's flat HBM buffer and their math runs in librat_b200.so
(gather K1, DNN head K3).  state_dict keys keep the reference's layer paths."""
