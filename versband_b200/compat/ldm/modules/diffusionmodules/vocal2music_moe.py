"""B200 override of the YAML target `ldm.modules.diffusionmodules.vocal2music_moe.TxtFlagLargeImprovedDiTV2`."""
from versband_b200.modules import TxtFlagLargeImprovedDiTV2  # noqa: F401
