"""whmr_b200 -- B200-native (sm_100a) body-model hot path for W-HMR.

Drop-in replacements, with the reference's own call signatures, for the functions SURVEY.md
section 8 puts on the path: `SMPL.forward`, `projection`, `perspective_projection`,
`convert_pare_to_full_img_cam`, `MAF_Extractor.sampling/forward/project`, the Regressor's
vertex read-outs (H36M joints, mesh down-sampling, SSM markers) and the evaluation metrics.
All compute runs in hand-written CUDA kernels behind the C-ABI library declared in
`include/whmr_b200.h`; there is no CPU or PyTorch fallback -- a missing library raises.
"""
from . import constants  # noqa: F401

__all__ = ["constants"]
__version__ = "0.1.0"
