"""hydrotools.cuda -- B200-native (sm_100a) drop-in for the raster-hydrology hot
path of bluegeo/hydro-tools:

* ``flow_direction_accumulation`` / ``FlowAccumulation``
  (/root/reference/src/hydrotools/watershed.py:30-75, 321-409)
* ``slope`` / ``aspect`` / ``terrain_ruggedness_index``
  (/root/reference/src/hydrotools/elevation.py:24-56, 59-80, 106-128)
* ``raster_filter`` (moving-window reducers;
  /root/reference/src/hydrotools/interpolate.py:110-247)
* ``reach_labels`` / ``upstream_stats`` / ``basin`` (consumers of the direction grid;
  /root/reference/src/hydrotools/streams.py:320-388, morphology.py:50-153, 292-360,
  watershed.py:292-318)

Python host code over a C-ABI shared library of hand-written CUDA kernels
(include/hydrotools_b200.h).  No CPU fallback: importing this package without
the built library, or calling it without a CUDA device, raises.
"""
from . import _native  # noqa: F401  (raises ImportError if the .so is missing)
from .elevation import slope, aspect, terrain_ruggedness_index, terrain  # noqa: F401
from .interpolate import raster_filter, window_filter  # noqa: F401
from .streams import (  # noqa: F401
    basin,
    basin_mask,
    reach_labels,
    upstream_mean,
    upstream_stats,
)
from .pipeline import FlowPipeline, flow_direction_accumulation_many  # noqa: F401
from .watershed import (  # noqa: F401
    FlowAccumulation,
    flow_accumulation_arrays,
    flow_direction_accumulation,
    flow_direction_accumulation_arrays,
    validate_conditioned,
)

__all__ = [
    "slope",
    "aspect",
    "terrain_ruggedness_index",
    "terrain",
    "flow_direction_accumulation",
    "flow_direction_accumulation_arrays",
    "flow_direction_accumulation_many",
    "FlowPipeline",
    "FlowAccumulation",
    "flow_accumulation_arrays",
    "validate_conditioned",
    "raster_filter",
    "window_filter",
    "reach_labels",
    "upstream_mean",
    "upstream_stats",
    "basin_mask",
    "basin",
]
