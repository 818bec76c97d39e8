from .EMProblem import EMProblem
from .EMVectorWaveEigenproblem import EigenProblem, DefaultEigenSolver
from .EMDrivenProblem import DrivenProblemABC
