"""Overlay for the reference's ``utils`` package (see overlay/README.md): ``utils.train_one_epoch``, ``utils.misc`` and
``utils.lr_sched`` resolve to the B200 implementation, every other sub-module (``utils.feature_extraction``,
``utils.lr_decay``, ``utils.custom_loss``, ``utils.used_metrics``) falls through to the reference checkout named by
VITAE_REFERENCE_ROOT."""
import os
import sys

_ref = os.environ.get("VITAE_REFERENCE_ROOT")
if _ref and os.path.isdir(os.path.join(_ref, "utils")):
    __path__.append(os.path.join(_ref, "utils"))

from vit_ae_plus_plus_b200.utils import lr_sched, misc, train_one_epoch  # noqa: E402

sys.modules[__name__ + ".train_one_epoch"] = train_one_epoch
sys.modules[__name__ + ".misc"] = misc
sys.modules[__name__ + ".lr_sched"] = lr_sched
