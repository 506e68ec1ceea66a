"""Host-side mirror of the reference's ``model`` package for the MAE pre-training path (model_factory, vit_autoenc)."""
