from .framedata import FrameData  # noqa: F401
from .io import save_frame_data_to_file  # noqa: F401
