import sys
sys.path.insert(0, '.')
import numpy as np
import perphil_b200 as pb
from tests.util import make_problem
cells = tuple(int(c) for c in sys.argv[1:]) or (4, 4, 4)
W, p, bcs, osys = make_problem(cells, 1)
sol = pb.solve_dpp(W, p, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
print("its", sol.iteration_number, "rnorm", sol.residual_error)
