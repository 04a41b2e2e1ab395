from . import dataset
from .dataset import get_window_indices, preprocess_windows, DevicePoseWindows
