"""mmidas_b200 — B200-native (sm_100a) drop-in for the training step of MMIDAS's coupled
mixture-VAE (reference package: ``mmidas`` in AllenInstitute/distributed-vae).

Mirrors ``mmidas.nn_model`` (mixVAE_model, VAEConfig, mk_vae), ``mmidas.cpl_mixvae`` (cpl_mixVAE)
and ``mmidas._dist_utils``; everything numeric runs in ``libmixvae_b200.so``.
"""
from .nn_model import VAEConfig, mixVAE_model, mk_vae  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .augmentation import Augmenter_smartseq, mk_augmenter  # noqa: F401
