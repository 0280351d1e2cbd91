"""Slab geometry and quadrature (mirror of reference slab.py)."""
import numpy as np


class Slab:
    """1-D slab of `length` mean free paths cut into length*cells_per_mfp zones (slab.py:5-12)."""

    def __init__(self, length, cells_per_mfp, num_angles):
        self.length = length
        self.num_zones = self.length * cells_per_mfp
        self.hx = self.length / self.num_zones
        self.num_angles = num_angles
        self.mu, self.weight = np.polynomial.legendre.leggauss(num_angles)
        self.num_groups = 1


class SymmetricHeterogeneousSlab(Slab):
    """Kornreich-Parsons fuel | moderator | absorber | moderator | fuel slab (slab.py:15-56)."""

    def __init__(self, length, cells_per_mfp, num_angles, mod_mat, fuel_mat, abs_mat):
        super().__init__(length, cells_per_mfp, num_angles)
        self.fuel_thick = 1
        self.mod_thick = 1
        self.num_regions = 5
        self.fuel_material = fuel_mat
        self.moderator_material = mod_mat
        self.absorber_material = abs_mat
        self.reflector1_right = 0
        self.fuel1_right = self.mod1_right = self.abs_right = self.mod2_right = self.fuel2_right = 0
        self.define_region_boundaries()
        groups = getattr(fuel_mat, "num_groups", 1)
        if groups and groups > 1:
            self.num_groups = groups

    def define_region_boundaries(self):
        """Right edges at 1, 2, 7, 8, 9 mfp (slab.py:33-39)."""
        self.fuel1_right, self.mod1_right, self.abs_right, self.mod2_right, self.fuel2_right = 1, 2, 7, 8, 9
        assert self.fuel2_right == self.length

    def region_of(self, midpoint):
        """(material, name) for a zone midpoint; closed/open interval ends as slab.py:41-56."""
        if midpoint <= self.fuel1_right:
            return self.fuel_material, "fuel"
        if self.fuel1_right < midpoint < self.mod1_right:
            return self.moderator_material, "moderator"
        if self.mod1_right <= midpoint <= self.abs_right:
            return self.absorber_material, "absorber"
        if self.abs_right < midpoint < self.mod2_right:
            return self.moderator_material, "moderator"
        if self.mod2_right <= midpoint <= self.fuel2_right:
            return self.fuel_material, "fuel"
        return None, "undefined"

    def assign_zone_material(self, zone):
        material, name = self.region_of(zone.midpoint)
        if material is not None:
            zone.material, zone.region_name = material, name

    def assign_all(self, midpoints):
        """Vectorised region_of: (materials list, names list, index array into the materials list)."""
        m = np.asarray(midpoints)
        mats = [self.fuel_material, self.moderator_material, self.absorber_material]
        names = ["fuel", "moderator", "absorber"]
        idx = np.full(m.shape, -1, dtype=np.int64)
        idx[m <= self.fuel1_right] = 0
        idx[(self.fuel1_right < m) & (m < self.mod1_right)] = 1
        idx[(self.mod1_right <= m) & (m <= self.abs_right)] = 2
        idx[(self.abs_right < m) & (m < self.mod2_right)] = 1
        idx[(self.mod2_right <= m) & (m <= self.fuel2_right)] = 0
        return mats, names, idx


class ScaledHeterogeneousSlab(SymmetricHeterogeneousSlab):
    """The same five-region layout with an arbitrary number of zones (synthetic benchmark slabs:
    SURVEY.md section 8(d) keeps the 1/9, 2/9, 7/9, 8/9 region fractions at 1e5 and 1e6 cells)."""

    def __init__(self, num_zones, num_angles, mod_mat, fuel_mat, abs_mat, length=9):
        super().__init__(length, 1, num_angles, mod_mat, fuel_mat, abs_mat)
        self.num_zones = int(num_zones)
        self.hx = self.length / self.num_zones


class MultiRegionSlab(Slab):
    """reflector | fuel | inner | fuel | reflector layout sketched at slab.py:59-68 (the reference
    version cannot be constructed: `super.__init__` at slab.py:61)."""

    def __init__(self, length, cells_per_mfp, num_angles, reflector_thick, inner_thick, num_zones, num_groups,
                 material1, material2):
        super().__init__(length, cells_per_mfp, num_angles)
        self.reflector_thick = reflector_thick
        self.inner_thick = inner_thick
        self.fuel_thick = (length - 2 * reflector_thick - inner_thick) / 2
        self.num_zones = num_zones
        self.hx = self.length / self.num_zones
        self.num_groups = num_groups
        self.metal = material1
        self.polyethylene = material2
        # edges used by Zone.assign_material (zone.py:16-30)
        self.reflector1_right = reflector_thick
        self.fuel1_right = reflector_thick + self.fuel_thick
        self.moderator_right = self.fuel1_right + inner_thick
        self.fuel2_right = self.moderator_right + self.fuel_thick
        self.reflector2_right = length
        self.reflector_material = material2
        self.moderator_material = material2
        self.fuel_material = material1

    def assign_zone_material(self, zone):
        zone.assign_material(zone.midpoint)
