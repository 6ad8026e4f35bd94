import sys; sys.path.insert(0, '.')
import numpy as np
import perphil_b200 as pb
from oracle import dpp_oracle as orc
from tests.util import make_problem, configured_handle
for cells in [(4,4,4),(8,8,8),(9,3,3),(3,9,3),(3,3,9)]:
    W,p,bcs,osys = make_problem(cells,1)
    ref = orc.solve_dpp_oracle(osys,"cg","jacobi")
    for every in (1,8):
        sol = pb.solve_dpp(W,p,bcs,solver_parameters={**pb.B200_CG_JACOBI_PARAMS,"b200_history":64,"b200_check_every":every})
        info = pb.last_solve_info()
        print(cells, every, 'its', sol.iteration_number, ref.iteration_number, 'reason', info.converged_reason)
        print('  gpu ', info.history[:5]); print('  ref ', np.array(ref.history[:5]))
