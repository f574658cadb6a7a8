"""GPU drop-ins for the hot-path members of pdspy.interferometry
(pdspy/interferometry/__init__.py:1-18): same names and call signatures."""
from .visibilities import Visibilities, VisibilitiesObject
from .interpolate_model import (interpolate_model, model_visibilities, loglike_image, loglike_images,
                                loglike_image_fft, loglike_image_nufft)
from .grid import grid, freqcorrect, chisq
from .average import average, center
from .invert import invert
from .cube import postprocess_channels
from .clean import clean
