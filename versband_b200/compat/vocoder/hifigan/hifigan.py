"""B200 override of the YAML target `vocoder.hifigan.hifigan.HifiGAN`."""
from versband_b200.modules import HifiGAN, HifiGanGenerator  # noqa: F401
