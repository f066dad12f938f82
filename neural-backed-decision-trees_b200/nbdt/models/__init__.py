"""Backbone factories of the package namespace (`nbdt.models.<name>`), same names as the reference exposes: the B200-kernel ResNets
and WideResNets, torchvision's model zoo, and -- shadowing torchvision's eager one -- the B200-kernel EfficientNet-B0."""
import types as _types

import torchvision.models as _tv

from . import efficientnet as _efficientnet
from . import resnet as _resnet
from . import wideresnet as _wideresnet


def _export(module):
    names = getattr(module, "__all__", None) or [n for n in vars(module) if not n.startswith("_")]
    for n in names:
        globals()[n] = getattr(module, n)


# later modules win: the package's EfficientNet-B0 replaces torchvision.models.efficientnet_b0
for _m in (_resnet, _wideresnet, _tv, _efficientnet):
    _export(_m)
from .resnet import *  # noqa: E402,F401,F403  (static names for linters / IDEs)
from .wideresnet import *  # noqa: E402,F401,F403
from .efficientnet import *  # noqa: E402,F401,F403


def get_model_choices():
    """Names a `--arch` flag may take: every public callable of this namespace (reference: nbdt/models/__init__.py:7-11)."""
    return (name for name, obj in sorted(globals().items())
            if name[:2] != "__" and callable(obj) and not isinstance(obj, _types.ModuleType) and not name.startswith("_"))
