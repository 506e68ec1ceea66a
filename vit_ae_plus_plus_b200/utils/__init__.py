"""Host-side mirror of the reference's ``utils`` package for the MAE pre-training loop (train_one_epoch, misc, lr_sched)."""
