"""SE(3) + categorical diffusers with the reference's `diffuser` package API (diffuser/*.py), running on
the sm_100a kernels of libabx_b200."""
from abx_b200.diffuser.full_diffuser import FullDiffuser  # noqa: F401
