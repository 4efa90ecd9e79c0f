"""Stand-in for the reference package (tests/fake_jaxmat/README.md): only ``dolfinx_materials.jaxmat.JAXMaterial``."""
