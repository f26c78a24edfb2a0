"""``import dolfin`` -> the fedm_b200 shim (put ``<repo>/compat`` on sys.path / PYTHONPATH)."""
from fedm_b200.dolfin import *          # noqa: F401,F403
from fedm_b200.dolfin import MPI, File, XDMFFile, parameters, set_backend, get_backend   # noqa: F401
