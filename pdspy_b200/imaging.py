"""Minimal Image container with the attributes the hot path reads, mirroring
pdspy/imaging/libimaging.pyx:7-48 (image[ny,nx,nfreq,npol] fp64, x, y in arcsec, freq/wave
deriving each other through c).  File I/O is out of scope; any object with .image, .x, .freq
(for instance pdspy's own Image) is accepted by interpolate_model."""
import numpy

_C = 2.99792458e10      # cm/s (pdspy/constants/physics.py)


class Image(object):

    def __init__(self, image=None, x=None, y=None, header=None, wave=None, freq=None, unc=None,
                 velocity=None, wcs=None):
        if image is not None:
            if not isinstance(image, numpy.ndarray) or image.dtype != numpy.float64 or image.ndim != 4:
                raise ValueError("image must be a 4-D float64 array [ny, nx, nfreq, npol]")
            self.image = image
        if x is not None:
            self.x = x
            self.y = y
        if header is not None:
            self.header = header
        if unc is not None:
            self.unc = unc
        if velocity is not None:
            self.velocity = velocity
        if (wave is None) and (freq is not None):
            self.freq = freq
            self.wave = _C / freq
        elif (wave is not None) and (freq is None):
            self.wave = wave
            self.freq = _C / wave
        elif (wave is not None) and (freq is not None):
            self.wave = wave
            self.freq = freq
        if wcs is not None:
            self.wcs = wcs


class UnstructuredImage(object):
    """Scattered-point image, mirroring pdspy/imaging/libimaging.pyx:186-222: image [npts, nfreq] (Jy/sr),
    x, y [npts] in arcsec, freq/wave deriving each other through c."""

    def __init__(self, image=None, x=None, y=None, wave=None, freq=None, unc=None, velocity=None):
        if image is not None and (not isinstance(image, numpy.ndarray) or image.dtype != numpy.float64 or image.ndim != 2):
            raise ValueError("image must be a 2-D float64 array [npts, nfreq]")
        self.image, self.x, self.y, self.unc, self.velocity = image, x, y, unc, velocity
        if (wave is None) and (freq is not None):
            self.freq = freq
            self.wave = _C / freq
        elif (wave is not None) and (freq is None):
            self.wave = wave
            self.freq = _C / wave
        elif (wave is not None) and (freq is not None):
            self.wave = wave
            self.freq = freq
