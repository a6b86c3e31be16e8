# same exported names as the reference registry (annotator/models/tf_models/__init__.py:1-2);
# engine.py:267-268 looks models up with getattr(tf_models, model_name)
from .unet import UNet, UNetAnnotator, MulmoUNet, MulmoUNetAnnotator  # noqa: F401
from .multiresunet import MultiResUnet, MultiResBlock  # noqa: F401
