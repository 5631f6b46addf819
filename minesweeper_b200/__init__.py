"""minesweeper_b200 -- B200-native likelihood hot path of MINESweeper.

Public surface mirrors the reference for this path:
``likelihood.likelihood``, ``fastMISTmod.GenMIST``, ``genmod.GenMod``,
``fitstar.lnprobfn`` (+ batched variants).  Importing the package does not
touch the GPU; constructing an ``Engine`` / ``likelihood`` does and fails
loudly when libmsb200.so or an sm_100 device is missing.
"""
__version__ = '0.1.0'
