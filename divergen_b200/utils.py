"""`diffusers.utils` names the reference imports (`from diffusers.utils import pt_to_pil`,
DiverGen/generation/txt2img_diffusers_stages_from_txt.py:8)."""
from .pipeline import pt_to_pil  # noqa: F401
