"""``JAXMaterial`` look-alike (tests/fake_jaxmat/README.md): the members make_golden_jaxmat.py calls on the reference's
class (dolfinx_materials/jaxmat.py:141-234), with the arithmetic of oracle/jaxmat_form.py and the state carried from
one increment to the next."""
import numpy as np

from oracle import jaxmat_form as jf


class _DataManager:
    def __init__(self, mat):
        self._m = mat

    def update(self):
        self._m.s0 = self._m.s1


class JAXMaterial:
    def __init__(self, behavior):
        self.behavior = behavior
        self.finite = behavior.finite_strain
        self.data_manager = _DataManager(self)

    @property
    def internal_state_variable_names(self):
        return ["p", "be_bar"] if self.finite else ["p", "epsp"]

    def set_data_manager(self, n):
        if self.finite:
            F = np.zeros((n, 9))
            F[:, :3] = 1.0
            be = np.zeros((n, 6))
            be[:, :3] = 1.0
            self.s0 = {"F": F, "PK1": np.zeros((n, 9)), "p": np.zeros(n), "be_bar": be}
        else:
            self.s0 = {"strain": np.zeros((n, 6)), "stress": np.zeros((n, 6)), "p": np.zeros(n), "epsp": np.zeros((n, 6))}
        self.s1 = self.s0

    def integrate(self, gradients, dt=0):
        props = self.behavior.props()
        n = len(gradients)
        if self.finite:
            r = jf.fefp_integrate(gradients, self.s0, props)
            self.s1 = {"F": np.array(gradients), "PK1": r["PK1"], "p": r["p"], "be_bar": r["be_bar"]}
            return r["PK1"], np.concatenate([r["p"].reshape(n, 1), r["be_bar"]], axis=1), r["Ct"]
        r = jf.j2_integrate(gradients, self.s0, props)
        self.s1 = {"strain": np.array(gradients), "stress": r["stress"], "p": r["p"], "epsp": r["epsp"]}
        return r["stress"], np.concatenate([r["p"].reshape(n, 1), r["epsp"]], axis=1), r["Ct"]

    def get_final_state_dict(self):
        return dict(self.s1)
