"""divergen_b200 -- B200-native Stable-Diffusion denoising hot path behind the diffusers class surface.

Python here is host plumbing only (argument checking, ctypes, torch for device memory / streams / distributed);
all arithmetic runs in hand-written sm_100a CUDA inside libdivergen_b200.so (include/divergen_b200.h).
"""
from .scheduler import DDIMScheduler, DDIMSchedulerOutput  # noqa: F401
from .unet import SD15_CONFIG, SD21_CONFIG, UNet2DConditionModel, UNet2DConditionOutput  # noqa: F401
from .vae import SD_VAE_CONFIG, AutoencoderKL, DecoderOutput  # noqa: F401
from .clip import SD21_CLIP_CONFIG, SD_CLIP_CONFIG, CLIPScorer, CLIPTextModel  # noqa: F401
from .preprocess import clip_preprocess, resize_u8  # noqa: F401
from .pipeline import StableDiffusionPipeline, StableDiffusionPipelineOutput, pt_to_pil  # noqa: F401
from .loading import DiffusionPipeline  # noqa: F401

__version__ = "0.1.0"
