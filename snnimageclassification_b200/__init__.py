"""B200-native spiking hot path behind the Python surface of JeremieGince/SNNImageClassification.

The public names mirror the reference so that ``from src.modules.snn import SNN`` becomes
``from snnimageclassification_b200.modules.snn import SNN`` and nothing else changes:

    snnimageclassification_b200.modules.snn            SNN, LoadCheckpointMode          (src/modules/snn.py)
    snnimageclassification_b200.modules.spiking_layers LayerType, LIFLayer, ALIFLayer,  (src/modules/spiking_layers.py)
                                                       ReadoutLayer, LayerType2Layer
    snnimageclassification_b200.modules.spike_funcs    SpikeFuncType, SpikeFuncType2Func, (src/modules/spike_funcs.py)
                                                       HeavisideSigmoidApprox, HeavisidePhiApprox
    snnimageclassification_b200.datasets.datasets      ToSpikes, DatasetId              (src/datasets/datasets.py)

Every FLOP of the path runs in hand-written sm_100a CUDA kernels (csrc/, C ABI in include/snnk.h).  There is no
CPU, eager-PyTorch or Triton fallback: without the built extension or off a B200 the compute calls raise.
"""
from ._cabi import build_extension  # noqa: F401
from .datasets.datasets import DatasetId, ToSpikes, unpack_raster  # noqa: F401
from .modules.optim import FusedAdam  # noqa: F401
from .modules.snn import SNN, LoadCheckpointMode  # noqa: F401
from .modules.spike_funcs import (  # noqa: F401
	HeavisidePhiApprox, HeavisideSigmoidApprox, SpikeFuncType, SpikeFuncType2Func, SpikeFunction)
from .modules.spiking_layers import (  # noqa: F401
	ALIFLayer, LayerType, LayerType2Layer, LIFLayer, ReadoutLayer, RNNLayer)

__version__ = "0.1.0"
