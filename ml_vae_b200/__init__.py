"""ml_vae_b200: B200-native (sm_100a) implementation of the ML-VAE data-parallel
training hot path -- fused fbank front-end + VAE latent block -- behind the
reference's own module / recipe interface.  See DESIGN.md and INTEGRATION.md.

The arithmetic lives in libmlvae_b200.so (C ABI: include/mlvae_b200.h); this
package is the host-side mirror of the reference interface.  There is no CPU
fallback: importing is cheap, but any compute call without the CUDA library or
on CPU tensors raises.
"""
__version__ = "0.1.0"
