"""volpath-b200: B200-native (sm_100a) implementation of CUDA-volpath's render hot path behind the
reference's own extern "C" host interface.  See DESIGN.md.  Importing the package does not load the CUDA
library; `Renderer` / `lib.load()` do, and fail loudly when it is missing (there is no CPU path)."""
from .param import Param, default_param, mat, MATERIALS  # noqa: F401
from .camera import inv_view_matrix  # noqa: F401
from .sunsky import default_sunsky, default_sky_state, constant_sky  # noqa: F401
from .sharding import frames_for_rank, reduce_accumulators, init_nccl_from_torch, init_nccl_via_store, split_frames  # noqa: F401
from . import io, lib  # noqa: F401
from .lib import (VOXEL_U8, VOXEL_F16, VOXEL_F32, BOUNDS_VOXEL, BOUNDS_CELL, BOUNDS_EXACT, MODE_PARITY, MODE_FAST, MODE_WAVE,  # noqa: F401
                  VolpathError)
from .renderer import Renderer  # noqa: F401
