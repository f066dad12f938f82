from .resnet import *
from .wideresnet import *
from torchvision.models import *
from .efficientnet import *  # B200-kernel EfficientNet-B0 replaces torchvision's eager efficientnet_b0


def get_model_choices():
    from types import ModuleType

    for key, value in list(globals().items()):
        if not key.startswith("__") and not isinstance(value, ModuleType) and callable(value):
            yield key
