"""Version marker.  run_expid.py:15 of the reference asserts `fuxictr.__version__.startswith("1.2")`; the local
suffix tells this B200-native implementation apart from the reference package."""
__version__ = "1.2.3+b200"
