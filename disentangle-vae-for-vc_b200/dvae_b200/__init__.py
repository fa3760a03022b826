"""dvae_b200: B200-native (sm_100a) kernels for the Disentangled-VAE voice-conversion hot path.

Host side is Python; all compute goes through the C-ABI shared library `libdvae_b200.so`
(declared in include/dvae_b200.h).  There is no CPU or PyTorch fallback: importing `dvae_b200.lib`
without the built library raises.
"""
