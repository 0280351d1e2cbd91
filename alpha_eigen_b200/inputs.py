"""Problem builders (mirror of reference inputs.py)."""
import numpy as np

from .group import OneGroup
from .material import SimpleMaterial
from .slab import SymmetricHeterogeneousSlab
from .time_unit import TimeUnit


def kornreich_parsons_symmetric(length, cells_per_mfp, num_angles, time_steps=1, dt=0.10, nu_sigmaf=None, output=None):
    """Steady-state k or time-dependent alpha eigenvalues of the 9-mfp heterogeneous one-speed
    slab of Kornreich & Parsons (2005), solved as in McClarren (2019); same positional arguments, prints
    and defaults as inputs.py:7-52 (dt = 0.10, inputs.py:39).

    Extras beyond the reference, all optional:
      dt ......... time step of the march (the reference hard-codes 0.10)
      nu_sigmaf .. fuel nu*Sigma_f instead of 0.7 (material.py:13); 0.3 is the sub-critical variant of the
                   report's Table 1 (docs/ME_599_Final__alpha_eigen-6.pdf p.7)
      output ..... path of an .npz that receives the zone midpoints, the scalar flux and k or the alphas
                   (the flux profiles of the report's Fig. 3)
    """
    fuel = SimpleMaterial('fuel', num_groups=1)
    moderator = SimpleMaterial('moderator', num_groups=1)
    absorber = SimpleMaterial('absorber', num_groups=1)
    if nu_sigmaf is not None:
        fuel.nu_sigmaf = float(nu_sigmaf)
    slab = SymmetricHeterogeneousSlab(length=length, cells_per_mfp=cells_per_mfp, num_angles=num_angles,
                                      mod_mat=moderator, fuel_mat=fuel, abs_mat=absorber)
    one_group = OneGroup(slab=slab)
    one_group.create_zones()
    num_steps = int(time_steps)
    midpoints = np.array([z.midpoint for z in one_group.zone]) if output else None
    if num_steps == 1:
        print("*************************====================")
        print("Initiating k-eigenvalue solver")
        print("*************************====================")
        k = one_group.perform_power_iteration()
        print(k.new)
        if output:
            np.savez(output, midpoints=midpoints, phi=one_group.phi.new, k=k.new)
        return k
    time_dependent = TimeUnit(slab=slab, group=one_group, num_steps=num_steps, dt=dt)
    time_dependent.calculate_time_dependent_flux()
    alpha_eigs = time_dependent.compute_alpha_eigenvalues()
    print("alphas = ", alpha_eigs)
    if output:
        np.savez(output, midpoints=midpoints, phi=one_group.phi.new, phi_steps=time_dependent.phi, alphas=alpha_eigs,
                 dt=dt)
    return alpha_eigs


if __name__ == "__main__":
    kornreich_parsons_symmetric(length=9, cells_per_mfp=50, num_angles=4, time_steps=50)
