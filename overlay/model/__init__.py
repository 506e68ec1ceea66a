"""Overlay for the reference's ``model`` package (see overlay/README.md): ``model.model_factory`` and
``model.vit_autoenc`` resolve to the B200 implementation, every other sub-module (``model.vit``, ``model.model_utils``)
falls through to the reference checkout named by VITAE_REFERENCE_ROOT."""
import os
import sys

_ref = os.environ.get("VITAE_REFERENCE_ROOT")
if _ref and os.path.isdir(os.path.join(_ref, "model")):
    __path__.append(os.path.join(_ref, "model"))

from vit_ae_plus_plus_b200.model import model_factory, vit_autoenc  # noqa: E402

sys.modules[__name__ + ".model_factory"] = model_factory
sys.modules[__name__ + ".vit_autoenc"] = vit_autoenc
