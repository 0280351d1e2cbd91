"""Per-angle holder (mirror of the live part of reference one_direction.py:8-14).

The reference also builds a dense (Z+1)^2 "coefficient matrix" per angle (one_direction.py:34-39)
that is never used and costs 80 GB per angle at 1e5 cells; it is not reproduced.  psi_midpoint is
fetched from the device on first access after a solve instead of being written by every sweep.

`boundary_condition` (one_direction.py:12) is the flux entering the slab at this angle's upstream boundary: the reference
puts it into the first / last entry of the right-hand side it builds in set_positive / set_negative_boundary_condition
(one_direction.py:24-31) and never sets it to anything but 0 (vacuum, which its live sweep hard-codes, group.py:48,54).
Here a non-zero value (scalar, or one per group) is honoured by every sweep: it is the initial value of the recurrence.
"""
import numpy as np


class OneDirection(object):
    def __init__(self, number, slab, group):
        self.number = number
        self.slab = slab
        self.group = group
        self.mu = self.slab.mu[number]
        self.weight = self.slab.weight[number]
        self.boundary_condition = 0
        self._psi_midpoint = None
        self._psi_zone_boundaries = None

    @property
    def psi_midpoint(self):
        """(Z, G) cell-centre angular flux of the last sweep of this angle (group.py:59)."""
        self.group._refresh_psi()
        if self._psi_midpoint is None:
            self._psi_midpoint = np.zeros((self.slab.num_zones, self.slab.num_groups))
        return self._psi_midpoint

    @psi_midpoint.setter
    def psi_midpoint(self, value):
        self._psi_midpoint = value

    @property
    def psi_zone_boundaries(self):
        if self._psi_zone_boundaries is None:
            self._psi_zone_boundaries = np.zeros(self.slab.num_zones + 1)
        return self._psi_zone_boundaries
