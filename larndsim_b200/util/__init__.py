"""Drop-in for ``larndsim.util`` pieces that sit on the charge path: :class:`CudaDict`, :class:`TPCBatcher`."""
from .cuda_dict import CudaDict  # noqa: F401
from .batching import TPCBatcher, TrackSegmentBatcher  # noqa: F401
