"""lnlike with the reference's signature (pdspy/utils/emcee.py:6-61).

The model runners (run_disk_model / run_flared_model: RADMC-3D orchestration) are outside the
hot path; pass one as `run_model=`, or have pdspy importable to use its own.  Everything after
the model exists - the visibility term (:31-43) - runs on the GPU; the image and SED terms
(:46-57) are the reference's small numpy expressions."""
import numpy

from .likelihood import lnlike_visibilities


def _default_runner(model):
    try:
        from pdspy.modeling import run_disk_model, run_flared_model
    except Exception as e:         # pragma: no cover - pdspy itself is not a dependency
        raise RuntimeError("no model runner: pass run_model=... or install pdspy (%s)" % e)
    return run_disk_model if model == "disk" else run_flared_model


def lnlike(params, visibilities, images, spectra, parameters, plot, model="disk", ncpus=1,
           ncpus_highmass=1, with_hyperion=False, timelimit=3600, source="ObjName", nice=19,
           verbose=False, ftcode="galario", run_model=None):

    runner = run_model if run_model is not None else _default_runner(model)
    if model == "disk":
        m = runner(visibilities, images, spectra, params, parameters, plot, ncpus=ncpus,
                   ncpus_highmass=ncpus_highmass, with_hyperion=with_hyperion, timelimit=timelimit,
                   source=source, nice=nice, verbose=verbose, ftcode=ftcode)
    elif model == "flared":
        m = runner(visibilities, params, parameters, plot, ncpus=ncpus, source=source, nice=nice,
                   ftcode=ftcode)

    # Catch whether the model timed out (:22-23).
    if isinstance(m, float) and m == 0.:
        return -numpy.inf

    chisq = lnlike_visibilities(visibilities, m)

    for j in range(len(images["file"])):
        chisq.append(-0.5 * (numpy.sum((images["data"][j].image - m.images[images["lam"][j]].image) ** 2 /
                                       images["data"][j].unc ** 2)))

    if "total" in spectra:
        chisq.append(-0.5 * (numpy.sum((spectra["total"].flux - m.spectra["SED"].flux) ** 2 /
                                       spectra["total"].unc ** 2)))

    return numpy.array(chisq).sum()
