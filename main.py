"""Command-line front end with the reference's flags (main.py:9-13) and defaults (main.py:18-26).

    python main.py --problem kornreich-parsons-symmetric [--length 9] [--cells_per_mfp 50]
                   [--num_angles 16] [--steps 1] [--dt 0.1] [--nu_sigmaf 0.7] [--output flux.npz]
"""
import argparse

from alpha_eigen_b200.inputs import kornreich_parsons_symmetric


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--problem', type=str, required=True, help='options: kornreich-parsons-symmetric')
    parser.add_argument('--length', type=int, required=False)
    parser.add_argument('--cells_per_mfp', type=int, required=False)
    parser.add_argument('--num_angles', type=int, required=False)
    parser.add_argument('--steps', type=int, required=False)        # the reference declares str (defect D1)
    # beyond the reference (all optional; defaults reproduce it)
    parser.add_argument('--dt', type=float, required=False, help='time step of the march (reference: 0.10)')
    parser.add_argument('--nu_sigmaf', type=float, required=False,
                        help='fuel nu*Sigma_f (reference: 0.7; 0.3 = sub-critical case of the report, Table 1)')
    parser.add_argument('--output', type=str, required=False, help='.npz for midpoints, flux and eigenvalues')
    args = parser.parse_args(argv)
    if args.problem == "kornreich-parsons-symmetric":
        kornreich_parsons_symmetric(length=args.length or 9, cells_per_mfp=args.cells_per_mfp or 50,
                                    num_angles=args.num_angles or 16, time_steps=args.steps or 1,
                                    dt=args.dt or 0.10, nu_sigmaf=args.nu_sigmaf, output=args.output)


if __name__ == "__main__":
    main()
