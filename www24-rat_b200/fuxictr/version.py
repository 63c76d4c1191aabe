__version__ = "1.2.3+b200"
