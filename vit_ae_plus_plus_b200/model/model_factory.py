"""Drop-in for the reference's ``model/model_factory.py`` (:9-17): same ``get_models(model_name, args)`` dispatch for
the auto-encoder entries and for 'vit' (the encoder-only feature extractor of the post-training half of the k-fold scripts,
SURVEY.md row f-3).  'contrastive' (model_factory.py:23-27) is a fine-tuning workload outside this package's path and raises
NotImplementedError naming that."""
from . import vit_autoenc


def get_models(model_name, args):
    if model_name in ("autoenc", "autoenc_contr"):
        print(f"Number of channels is {args.in_channels}")
        return vit_autoenc.__dict__[args.model](volume_size=args.volume_size, in_chans=args.in_channels,
                                                patch_size=args.patch_size, args=args)
    if model_name == "vit":      # the feature extractor the k-fold scripts load the MAE checkpoint into (model_factory.py:19-22)
        from functools import partial

        from torch import nn

        from .vit import VisionTransformer3D
        return VisionTransformer3D(volume_size=args.volume_size, in_chans=args.in_channels, num_classes=args.nb_classes,
                                   patch_size=args.patch_size, global_pool=args.global_pool,
                                   norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_path_rate=args.drop_path)
    if model_name == "contrastive":
        raise NotImplementedError(
            "model_name='contrastive': VisionTransformer3DContrastive (supervised contrastive fine-tuning, model/vit.py:301-337) "
            "is not part of the B200 pre-training path; use the reference's model.vit for it")
    raise NotImplementedError("Only AE model supported till now")
