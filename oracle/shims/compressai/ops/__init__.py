from .bound_ops import LowerBound
from .parametrizers import NonNegativeParametrizer
