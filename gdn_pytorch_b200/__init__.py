"""gdn_pytorch_b200: B200-native (sm_100a) implementation of the GDN-Pytorch autoencoder hot path."""
__all__ = ["AE_model_unet"]
