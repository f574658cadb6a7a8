"""lnlike with the reference's dynesty signature (pdspy/utils/dynesty.py:11-76): the same
visibility term as utils/emcee.py, with the parameter vector unpacked the way dynesty.py:16-28
does (free parameters in sorted key order)."""
import numpy

from . import emcee as _emcee


def lnlike(p, visibilities, images, spectra, parameters, plot, model="disk", ncpus=1,
           ncpus_highmass=1, with_hyperion=False, timelimit=3600, source="ObjName", nice=19,
           verbose=False, ftcode="galario", run_model=None):
    keys = []
    for key in sorted(parameters.keys()):
        if not parameters[key]["fixed"]:
            keys.append(key)
    params = dict(zip(keys, p))
    return _emcee.lnlike(params, visibilities, images, spectra, parameters, plot, model=model, ncpus=ncpus,
                         ncpus_highmass=ncpus_highmass, with_hyperion=with_hyperion, timelimit=timelimit,
                         source=source, nice=nice, verbose=verbose, ftcode=ftcode, run_model=run_model)
