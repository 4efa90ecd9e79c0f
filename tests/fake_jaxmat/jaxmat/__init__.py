"""Stand-in for jaxmat (tests/fake_jaxmat/README.md): behaviour descriptors with jaxmat's constructor signatures."""
__version__ = "0.0-standin (oracle/jaxmat_form.py)"
from . import materials  # noqa: F401
