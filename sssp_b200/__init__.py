"""sssp_b200 -- B200-native connect / collide / roadmap hot path of Kei18/sssp (MRMP.jl).

Host-side mirror of the reference's closure API (gen_connect / gen_collide /
gen_check_goal, src/MRMP.jl:98) on top of the C-ABI CUDA library
(include/mrmp_b200.h).  There is no CPU fallback.
"""
from .lib import (  # noqa: F401
    MRMPError, MODEL_POINT2D, MODEL_POINT3D, MODEL_ARM22, MODEL_LINE2D, MODEL_SNAKE2D, MODEL_ARM33,
    MODEL_CAPSULE3D,
)
from .closures import (  # noqa: F401
    StatePoint2D, StatePoint3D, StateArm22, StateLine2D, StateSnake2D, StateArm33, StateCapsule3D,
    CircleObstacle2D, CircleObstacle3D, Node,
    Context, Roadmaps, SsspRoadmaps, gen_connect, gen_collide, gen_check_goal, PRMs,
)
from . import closures  # noqa: F401
from .solvers import SSSP, get_distance_table, get_distance_tables  # noqa: F401
from .validation import validate, is_valid_instance  # noqa: F401
from .pp import PP, PP_on_roadmaps  # noqa: F401

__all__ = [
    "MRMPError", "MODEL_POINT2D", "MODEL_POINT3D", "MODEL_ARM22",
    "StatePoint2D", "StatePoint3D", "StateArm22", "StateLine2D", "StateSnake2D", "StateArm33",
    "StateCapsule3D", "CircleObstacle2D", "CircleObstacle3D", "Node",
    "Context", "Roadmaps", "SsspRoadmaps", "SSSP", "PP", "PP_on_roadmaps", "validate", "is_valid_instance", "gen_connect", "gen_collide", "gen_check_goal", "PRMs",
]
