from . import emcee, dynesty
from .likelihood import visibility_lnlike, lnlike_visibilities
