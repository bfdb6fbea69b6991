"""matinvent_b200 — B200-native (sm_100a) implementation of MatInvent's hot path: the batched
reverse-diffusion crystal sampler and the reward-weighted fine-tuning step of the DiffCSP back-end,
behind the reference's `models/suite` + `pipeline/base.py` plugin API.  All arithmetic runs in
libmatinvent_b200.so (hand-written CUDA, include/matinvent_b200.h); there is no CPU fallback."""
__version__ = "0.1.0"
