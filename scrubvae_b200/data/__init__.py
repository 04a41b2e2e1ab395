from . import dataset
from .dataset import get_window_indices, preprocess_windows, DevicePoseWindows
from . import skeleton
from .skeleton import read_skeleton
