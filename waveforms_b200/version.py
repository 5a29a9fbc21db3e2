__version__ = "0.1.0"
# API level of the reference this package is a drop-in for
REFERENCE_VERSION = "2.2.3"
