from .base_boundary_condition import BoundaryCondition
from .ABC import ABCBoundaryCondition
from .container import BoundaryConditions
from .essential import EssentialBoundaryCondition, PECWallsBoundaryCondition
