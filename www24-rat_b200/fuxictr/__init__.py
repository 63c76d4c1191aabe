"""Drop-in `fuxictr` package: the FuxiCTR model API of the reference (run_expid.py works unchanged) with the
RAT hot path executed by librat_b200.so on a B200.  Only what the RAT path needs is provided."""
from .version import __version__
