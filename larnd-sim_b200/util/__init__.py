"""Drop-in for ``larndsim.util`` pieces that sit on the charge path: :class:`CudaDict`."""
from .cuda_dict import CudaDict  # noqa: F401
