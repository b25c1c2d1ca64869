"""B200 override of the YAML target `ldm.models.autoencoder1d.AutoencoderKL`."""
from versband_b200.modules import AutoencoderKL, DiagonalGaussianDistribution  # noqa: F401
