"""Command-line front end with the reference's flags (main.py:9-13) and defaults (main.py:18-26).

    python main.py --problem kornreich-parsons-symmetric [--length 9] [--cells_per_mfp 50]
                   [--num_angles 16] [--steps 1]
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
    args = parser.parse_args(argv)
    if args.problem == "kornreich-parsons-symmetric":
        kornreich_parsons_symmetric(length=args.length or 9, cells_per_mfp=args.cells_per_mfp or 50,
                                    num_angles=args.num_angles or 16, time_steps=args.steps or 1)


if __name__ == "__main__":
    main()
