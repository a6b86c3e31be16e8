"""dnncancerannotator_b200 -- B200-native conv-stack hot path of DNNCancerAnnotator.

Drop-in for the reference's ``annotator.models.tf_models`` model classes and the
``WeightedCrossentropy`` loss, executed by hand-written sm_100a CUDA kernels
(``libdnnca.so``, C ABI in ``include/dnnca.h``).  No TensorFlow, no CPU fallback.
"""
from . import native  # noqa: F401
from . import models  # noqa: F401

__version__ = '0.1.0'
