"""pdspy_b200: B200-native (sm_100a) model-to-visibility path for pdspy.

  interferometry.interpolate_model / grid / freqcorrect / chisq / Visibilities
  utils.emcee.lnlike, utils.dynesty.lnlike (visibility term on the GPU)

Python here is a thin mirror of the reference's call signatures over the C-ABI library
libpdsb.so (include/pdsb.h).  Importing the package does not touch CUDA; the first compute call
loads the library and initialises the device, and raises if either is missing (no CPU fallback).
"""
from . import constants
from . import interferometry
from . import utils
from .imaging import Image, UnstructuredImage
from .device import Dataset, clear_cache, release, set_cache_policy, set_dft_kernel
from ._lib import PdsbError, DeviceBuffer, PinnedArray

__version__ = "0.1.0"
