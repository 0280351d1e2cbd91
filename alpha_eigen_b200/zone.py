"""Per-cell record (mirror of reference zone.py)."""
import numpy as np

from .phi import Phi


class Zone:
    def __init__(self, slab, midpoint, _draw=None):
        self.slab = slab
        self.material = None
        self.midpoint = midpoint
        self.region_name = "undefined"
        self.phi = Phi(self.slab.num_groups, _draw=_draw)      # zone.py:11 (G draws from the global RNG)
        self.source = np.zeros(self.slab.num_groups)
        self.fission_source = self.source.copy()
        self.out_group_scatter_xs = None

    def assign_material(self, midpoint):
        """reflector | fuel | moderator | fuel | reflector map (zone.py:16-30)."""
        s = self.slab
        if midpoint <= s.reflector1_right:
            self.material, self.region_name = s.reflector_material, "reflector"
        elif s.reflector1_right < midpoint < s.fuel1_right:
            self.material, self.region_name = s.fuel_material, "fuel"
        elif s.fuel1_right <= midpoint <= s.moderator_right:
            self.material, self.region_name = s.moderator_material, "moderator"
        elif s.moderator_right < midpoint < s.fuel2_right:
            self.material, self.region_name = s.fuel_material, "fuel"
        elif s.fuel2_right <= midpoint <= s.reflector2_right:
            self.material, self.region_name = s.reflector_material, "reflector"
        # in-group scattering is treated separately: zero the diagonal (zone.py:32-36)
        sc = np.array(self.material.scatter_xs, dtype=np.float64, ndmin=2)
        self.out_group_scatter_xs = sc - np.diag(np.diag(sc))

    def compute_fission_source(self):
        """0.5*chi*sum_g' nu_sigma_f(g')*phi_old(g')  (zone.py:38-42)."""
        total = np.sum(self.material.nu_sigmaf * self.phi.old)
        self.fission_source = 0.5 * self.material.chi * total

    def reset_scattering_source(self):
        self.source = self.fission_source

    def compute_scattering_source(self):
        """zone.py:48-51."""
        for g_prime in range(self.slab.num_groups):
            self.source = self.source + 0.5 * (
                self.phi.new[g_prime] * self.out_group_scatter_xs[g_prime, :]
                + self.material.chi * self.phi.new[g_prime] * self.material.nu_sigmaf[g_prime])
