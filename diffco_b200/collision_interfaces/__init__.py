"""Robot descriptions for the checkers — mirror of the reference's ``diffco/collision_interfaces`` restricted to what the
hot path needs: the URDF kinematic tree (``URDFRobot``), compiled into a joint program that the CUDA kernels execute
(SURVEY.md §8 row f3).  Geometry back-ends (python-fcl, trimesh, cuRobo, ROS) are not part of this package."""
from .urdf_interface import MultiURDFRobot, URDFRobot, parse_urdf  # noqa: F401
