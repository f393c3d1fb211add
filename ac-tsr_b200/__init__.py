"""ac-tsr_b200: B200-native (sm_100a) AC-SASRec training + full-sort evaluation hot path.

Drop-in surface (names mirror the reference, AIM-SE/AC-TSR = RecBole 1.0.1 fork):
    ACSASRec                      recbole/model/sequential_recommender/acsasrec.py
    AcBERT4Rec, ACSSEPT, ACTiSASRec   .../acbert4rec.py, acssept.py, actisasrec.py (sibling models on the same kernels)
    AttackRTransformerEncoder...  recbole/model/layers.py:614-1131
    ACSASRecTrainer               recbole/trainer/trainer.py:505-1044
    Config / Interaction          recbole/config/configurator.py, recbole/data/interaction.py
All arithmetic runs in csrc/libacsr.so (C ABI: include/acsr.h).  No CPU fallback.

The directory name contains a hyphen; import it as `ac_tsr_b200` (shim module at the repo root).
"""
from . import _lib, build, compat, data, dataset, evaluator, layers, ops, trainer, acsasrec, acbert4rec, acssept, actisasrec, transformer_layers, fused_step, dist, quick_start    # noqa: F401
from ._lib import LIB, AcsrError                                                      # noqa: F401
from .acsasrec import ACSASRec                                                        # noqa: F401
from .acbert4rec import AcBERT4Rec                                                    # noqa: F401
from .acssept import ACSSEPT                                                          # noqa: F401
from .actisasrec import ACTiSASRec                                                    # noqa: F401
from .compat import Config, Interaction, ModelType                                    # noqa: F401
from .layers import (AttackRMultiHeadAttention, AttackRTransformerEncoder,            # noqa: F401
                     AttackRTransformerLayer, FeedForward)
from .trainer import ACSASRecTrainer, AcBERT4RecTrainer, ACSSEPTTrainer, ACTiSASRecTrainer, FlatAdam                     # noqa: F401
from .dataset import SequentialDataset, create_dataset, data_preparation             # noqa: F401
from .quick_start import run_recbole                                                  # noqa: F401

__version__ = '0.1.0'
