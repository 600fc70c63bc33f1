from .base import MergeTensorsBase  # noqa: F401
from .fast_fourier import FourierMerge  # noqa: F401
