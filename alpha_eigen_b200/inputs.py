"""Problem builders (mirror of reference inputs.py)."""
from .group import OneGroup
from .material import SimpleMaterial
from .slab import SymmetricHeterogeneousSlab
from .time_unit import TimeUnit


def kornreich_parsons_symmetric(length, cells_per_mfp, num_angles, time_steps=1):
    """Steady-state k or time-dependent alpha eigenvalues of the 9-mfp heterogeneous one-speed
    slab of Kornreich & Parsons (2005), solved as in McClarren (2019); same arguments, prints
    and defaults as inputs.py:7-52 (dt = 0.10, inputs.py:39)."""
    fuel = SimpleMaterial('fuel', num_groups=1)
    moderator = SimpleMaterial('moderator', num_groups=1)
    absorber = SimpleMaterial('absorber', num_groups=1)
    slab = SymmetricHeterogeneousSlab(length=length, cells_per_mfp=cells_per_mfp, num_angles=num_angles,
                                      mod_mat=moderator, fuel_mat=fuel, abs_mat=absorber)
    one_group = OneGroup(slab=slab)
    one_group.create_zones()
    num_steps = int(time_steps)
    dt = 0.10
    if num_steps == 1:
        print("*************************====================")
        print("Initiating k-eigenvalue solver")
        print("*************************====================")
        k = one_group.perform_power_iteration()
        print(k.new)
        return k
    time_dependent = TimeUnit(slab=slab, group=one_group, num_steps=num_steps, dt=dt)
    time_dependent.calculate_time_dependent_flux()
    alpha_eigs = time_dependent.compute_alpha_eigenvalues()
    print("alphas = ", alpha_eigs)
    return alpha_eigs


if __name__ == "__main__":
    kornreich_parsons_symmetric(length=9, cells_per_mfp=50, num_angles=4, time_steps=50)
