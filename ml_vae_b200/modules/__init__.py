"""Drop-in counterparts of the reference's src/modules/{fc_block,vanilla_vae,decoder}.py
with identical constructor kwargs, forward() dict contracts and state_dict keys."""
from .fc_block import FCBlock  # noqa: F401
from .vanilla_vae import VanillaVAE  # noqa: F401
from .decoder import Decoder  # noqa: F401
from .gmm_vae import GMMVAE  # noqa: F401
from .h_vae import HierarchicalVAE  # noqa: F401
from .lstm import LSTM  # noqa: F401
