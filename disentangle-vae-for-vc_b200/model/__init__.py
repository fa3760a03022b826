"""Drop-in `model` package (same import paths as the reference) backed by dvae_b200 kernels."""
