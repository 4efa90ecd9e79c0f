"""Stand-in for jax (tests/fake_jaxmat/README.md): only what make_golden_jaxmat.py touches."""
__version__ = "0.0-standin"


class _Config:
    def update(self, key, value):
        assert key == "jax_enable_x64" and value is True


config = _Config()
