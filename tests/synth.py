"""The synthetic workload generators live in the product package (tomofast-x_b200/synth.py) so that bench.py does not
depend on test code; the tests import them through this alias."""
from tomofastx_b200.synth import *          # noqa: F401,F403
from tomofastx_b200.synth import Problem, depth_weight_type1, make_problem, regular_grid, station_lattice  # noqa: F401
