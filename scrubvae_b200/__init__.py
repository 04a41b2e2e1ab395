"""scrubvae_b200 — B200-native (sm_100a) SC-VAE training step behind the scrubvae API."""
from . import model
from . import get
from . import train
from . import data
from . import eval  # noqa: A004
