"""Visibilities container: host-side mirror of VisibilitiesObject / Visibilities
(pdspy/interferometry/libinterferometry.pyx:9-52, 54-149).

Same constructor signature, attributes and pickling behaviour:
  * u, v, freq must be 1-D float64 arrays, real, imag, weights 2-D float64 (the reference's
    typed `numpy.ndarray[double, ndim=...]` arguments reject anything else with ValueError);
  * uvdist = sqrt(u**2+v**2), amp = sqrt(real**2+imag**2), phase = arctan2(imag, real),
    weights default to ones (:27, :35-36, :41).  The reference computes them eagerly in the
    constructor; here they are computed on first read (same values, and assignable as in the
    reference) so that a 1M x 64 model does not pay three extra host passes per likelihood call;
  * __reduce__ rebuilds a *VisibilitiesObject* (the base class), as the reference's
    `rebuild` does (:46-52) - what crosses MPI / dynesty checkpoints is the base class.
No CUDA handle is reachable from these objects (samplers pickle them).

Device-resident results: interpolate_model leaves the model visibilities on the GPU and returns a Visibilities
whose real / imag / weights are filled in on first read (`_device_token` is a plain integer naming the buffers in
pdspy_b200.device; it means nothing in another process).  utils.emcee.lnlike / visibility_lnlike consume the
device copy directly, so the unchanged-signature chain interpolate_model(...) -> lnlike(...) never moves the
[nuv, nf] arrays over PCIe; anything else that touches .real / .imag / .weights / .amp / .phase, or pickles
the object, gets ordinary numpy arrays.
"""
import weakref

import numpy


def _check(name, a, ndim):
    if a is None:
        return None
    if not isinstance(a, numpy.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(a).__name__))
    if a.dtype != numpy.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got '%s'" % a.dtype)
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    return a


class VisibilitiesObject(object):

    def __init__(self, u=None, v=None, freq=None, real=None, imag=None, weights=None,
                 baseline=None, array_name="CARMA"):
        u = _check("u", u, 1)
        v = _check("v", v, 1)
        freq = _check("freq", freq, 1)
        real = _check("real", real, 2)
        imag = _check("imag", imag, 2)
        weights = _check("weights", weights, 2)
        self._real = self._imag = self._weights = None
        self._device_token = None
        self._shape = None
        self.u = self.v = self.freq = None
        self._uvdist = self._amp = self._phase = None

        if (u is not None) and (v is not None):
            self.u = u
            self.v = v

        if freq is not None:
            self.freq = freq

        if (real is not None) and (imag is not None):
            self.real = real
            self.imag = imag
            if weights is not None:
                self.weights = weights
            else:
                self.weights = numpy.ones((self.real.shape[0], self.real.shape[1]))

        self.baseline = baseline
        self.array_name = array_name

    # real / imag / weights: plain attributes in the reference; here they may still be on the device
    def _materialise(self):
        token = self._device_token
        if token is None:
            return
        self._device_token = None
        from .. import device
        ent = device.model_buffers(token)
        if ent is None:
            raise RuntimeError("the device copy of these model visibilities is gone (cache cleared or another process)")
        re, im, shape = ent
        self._real = re.download(shape)
        self._imag = im.download(shape)
        device.release_model(token)

    @property
    def real(self):
        self._materialise()
        return self._real

    @real.setter
    def real(self, value):
        self._materialise()
        self._real = value

    @property
    def imag(self):
        self._materialise()
        return self._imag

    @imag.setter
    def imag(self, value):
        self._materialise()
        self._imag = value

    @property
    def weights(self):
        if self._weights is None and self._shape is not None:
            self._weights = numpy.ones(self._shape)          # interpolate_model.py:57
        return self._weights

    @weights.setter
    def weights(self, value):
        self._weights = value

    @classmethod
    def _from_device(cls, u, v, freq, token, shape):
        """A result whose real / imag still live on the device (pdspy_b200.device.register_model)."""
        self = cls(u, v, freq)
        self._device_token = token
        self._shape = tuple(shape)
        from .. import device
        weakref.finalize(self, device.release_model, token)
        return self

    # derived arrays: same definitions as libinterferometry.pyx:27,35-36
    @property
    def uvdist(self):
        if self._uvdist is None and self.u is not None and self.v is not None:
            self._uvdist = numpy.sqrt(self.u ** 2 + self.v ** 2)
        return self._uvdist

    @uvdist.setter
    def uvdist(self, value):
        self._uvdist = value

    @property
    def amp(self):
        if self._amp is None and self.real is not None and self.imag is not None:
            self._amp = numpy.sqrt(self.real ** 2 + self.imag ** 2)
        return self._amp

    @amp.setter
    def amp(self, value):
        self._amp = value

    @property
    def phase(self):
        if self._phase is None and self.real is not None and self.imag is not None:
            self._phase = numpy.arctan2(self.imag, self.real)
        return self._phase

    @phase.setter
    def phase(self, value):
        self._phase = value

    def __reduce__(self):
        return (rebuild, (self.u, self.v, self.freq, self.real, self.imag,
                          self.weights, self.baseline, self.array_name))


def rebuild(u, v, freq, real, imag, weights, baseline, array_name):
    return VisibilitiesObject(u, v, freq, real, imag, weights, baseline, array_name)


class Visibilities(VisibilitiesObject):

    def get_baselines(self, num):
        incl = self.baseline == num

        return Visibilities(self.u[incl], self.v[incl], self.freq,
                            self.real[incl, :], self.imag[incl, :],
                            self.weights[incl, :], baseline=self.baseline[incl])

    def set_header(self, header):
        self.header = header

    def read(self, filename=None, usefile=None):
        """HDF5 reader (libinterferometry.pyx:91-116): keys u v freq real imag weights, cast
        to float64.  File formats are outside the hot path; needs h5py."""
        import h5py
        f = h5py.File(filename, "r") if usefile is None else usefile
        u = v = freq = real = imag = weights = None
        if ('u' in f) and ('v' in f):
            u = f['u'][...].astype(numpy.double)
            v = f['v'][...].astype(numpy.double)
        if 'freq' in f:
            freq = f['freq'][...].astype(numpy.double)
        if ('real' in f) and ('imag' in f):
            real = f['real'][...].astype(numpy.double)
            imag = f['imag'][...].astype(numpy.double)
            if 'weights' in f:
                weights = f['weights'][...].astype(numpy.double)
            else:
                weights = numpy.ones(real.shape)
        self.__init__(u, v, freq, real, imag, weights)
        if usefile is None:
            f.close()

    def write(self, filename=None, usefile=None):
        """HDF5 writer (libinterferometry.pyx:118-149); needs h5py."""
        import h5py
        f = h5py.File(filename, "w") if usefile is None else usefile
        for name in ("u", "v", "freq", "real", "imag"):
            a = getattr(self, name)
            if a is not None:
                f.create_dataset(name, a.shape, dtype='float64')[...] = a
        if self.weights is not None and self.real is not None:
            if not numpy.all(self.weights == 1.0):
                f.create_dataset("weights", self.weights.shape, dtype='float64')[...] = self.weights
        if usefile is None:
            f.close()
