"""Scalar-flux holders (mirror of reference phi.py).

The reference seeds NumPy's GLOBAL generator when this module is first imported (phi.py:2) and
every Phi draws its initial guess from it, so the order and count of draws is part of the
contract: OneGroup.__init__ draws Z*G numbers (group.py:18), then each Zone draws G more
(zone.py:11).  Both stay on the host, in NumPy, in that order.
"""
import numpy as np

np.random.seed(2021)        # phi.py:2


class Phi:
    """`new` / `old` / `converged`, as phi.py:5-13.

    Phi(n)     -> 1-D guess 2*rand(n)                      (phi.py:8, used per Zone)
    Phi(n, g)  -> (n, g) guess made symmetric about the slab centre: r + flip(r)  (phi.py:10-11)
    """

    def __init__(self, num_to_create, num2=None, _draw=None):
        if _draw is not None:                   # pre-drawn values (vectorised Zone construction)
            self.new = _draw
        elif num2 is None:
            self.new = np.random.rand(num_to_create) * 2
        else:
            half = np.random.rand(num_to_create, num2)
            self.new = half + np.flip(half, 0)
        self.old = self.new.copy()
        self.converged = False

    def reset_scalar_flux(self):                # phi.py:15-18
        self.old = self.new.copy()
        self.new[:] = 0.0 + 1e-12
        self.converged = False

    def power_iteration_flux_update(self, k):   # phi.py:20-21
        self.new = self.new / k


class ZoneFlux:
    """Three-bin energy condensation container (phi.py:24-28)."""

    def __init__(self, num_to_create):
        self.thermal = np.zeros(num_to_create)
        self.epithermal = np.zeros(num_to_create)
        self.fast = np.zeros(num_to_create)
