"""Drop-in `autovc_replicate` package (proposed AutoVC-style generator) on dvae_b200 kernels."""
