"""Samplers of the hot path (mirrors of cemc.mcmc.{Montecarlo, SGCMonteCarlo,
ParallelTempering, MCParameterSweep}, /root/reference/cemc/mcmc/)."""
from .averager import Averager  # noqa: F401
from .mc_observers import (EnergyEvolution, EnergyHistogram,  # noqa: F401
                           LowestEnergyStructure, MCObserver,
                           PairCorrelationObserver, SGCObserver,
                           SiteOrderParameter)
from .montecarlo import (CanNotFindLegalMoveError,  # noqa: F401
                         DidNotReachEquillibriumError, Montecarlo,
                         TooFewElementsError)
from .sgc_montecarlo import InvalidChemicalPotentialError, SGCMonteCarlo  # noqa: F401
from .parallel_tempering import ParallelTempering  # noqa: F401
from .mc_parameter_sweep import MCParameterSweep  # noqa: F401
