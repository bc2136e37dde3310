"""Scene compiler: MJCF + STL -> model arrays -> blob (see mjcf.py, blob.py)."""
from .mjcf import compile_mjcf, model_to_blob_arrays  # noqa: F401
from .blob import pack, unpack  # noqa: F401
