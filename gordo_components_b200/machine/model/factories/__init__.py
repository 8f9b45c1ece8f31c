from .feedforward_autoencoder import *  # noqa: F401,F403
from .lstm_autoencoder import *  # noqa: F401,F403
