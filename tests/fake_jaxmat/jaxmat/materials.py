"""Constructor signatures as the reference's call sites use them (plane_elastoplasticity.py:67-71,
tests/test_FeFp_jax.py:17-19)."""


class LinearElasticIsotropic:
    def __init__(self, E, nu):
        self.E, self.nu = E, nu


class VoceHardening:
    def __init__(self, sig0, sigu, b):
        self.sig0, self.sigu, self.b = sig0, sigu, b


class _Plastic:
    finite_strain = False

    def __init__(self, elasticity, yield_stress):
        self.elasticity, self.yield_stress = elasticity, yield_stress

    def props(self):
        y = self.yield_stress
        return dict(E=self.elasticity.E, nu=self.elasticity.nu, sig0=y.sig0, sigu=y.sigu, b=y.b)


class vonMisesIsotropicHardening(_Plastic):
    pass


class FeFpJ2Plasticity(_Plastic):
    finite_strain = True
