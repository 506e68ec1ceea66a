"""vit_ae_plus_plus_b200 -- B200-native (sm_100a) training path for the 3D ViT masked autoencoder of ViT-AE++.

Host side is Python/PyTorch (allocation, streams, torch.distributed); the arithmetic is hand-written CUDA in
libvitae_b200.so behind the C ABI declared in include/vitae_b200.h.  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
