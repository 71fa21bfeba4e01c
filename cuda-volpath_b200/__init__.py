"""volpath-b200: B200-native (sm_100a) implementation of CUDA-volpath's render hot path behind the
reference's own extern "C" host interface.  See DESIGN.md."""
from .param import Param, default_param, mat, MATERIALS  # noqa: F401
from .camera import inv_view_matrix  # noqa: F401
from .sunsky import default_sunsky, constant_sky  # noqa: F401
