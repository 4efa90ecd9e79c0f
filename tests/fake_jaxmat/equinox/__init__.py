"""Stand-in for equinox (tests/fake_jaxmat/README.md)."""
__version__ = "0.0-standin"
