"""``import fedm`` -> this repo's mirror of the FEDM problem-definition API (``fedm_b200.functions`` / ``file_io`` ...),
for scripts written against the reference package (put ``<repo>/compat`` on sys.path)."""
from . import functions, file_io, physical_constants, utils   # noqa: F401
from .file_io import files                                     # noqa: F401
