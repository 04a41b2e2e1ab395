from .model import model
from .eval import latents
