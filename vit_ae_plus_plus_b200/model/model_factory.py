"""Drop-in for the reference's ``model/model_factory.py`` (:9-17): same ``get_models(model_name, args)`` dispatch for
the auto-encoder entries.  The classifier entries ('vit', 'contrastive', model_factory.py:19-27) are downstream
workloads outside this package's path and raise NotImplementedError naming that."""
from . import vit_autoenc


def get_models(model_name, args):
    if model_name in ("autoenc", "autoenc_contr"):
        print(f"Number of channels is {args.in_channels}")
        return vit_autoenc.__dict__[args.model](volume_size=args.volume_size, in_chans=args.in_channels,
                                                patch_size=args.patch_size, args=args)
    if model_name in ("vit", "contrastive"):
        raise NotImplementedError(
            f"model_name={model_name!r}: the VisionTransformer3D classifiers are not part of the B200 pre-training "
            "path; use the reference's model.vit for them")
    raise NotImplementedError("Only AE model supported till now")
