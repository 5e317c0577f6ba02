"""Drop-in for dmm/modules/submodules/relax_match.py: relax_matching / hungarian_matching with the reference's
signatures and return values, solved by the persistent warp-per-problem kernel (K3)."""
import time

import numpy as np
import torch

from ... import ops


def relax_matching(C, max_iter=100, proj_iter=100, lr=0.1, return_time=0):
    """C [n templates, m proposals] -> (X, cost, X_list, inner_projection_error[, seconds])
    (reference relax_match.py:36-105).  X_list[0] is the greedy start, X_list[k] the iterate right after
    gradient step k (before projection); cost[0] == 0.  Differentiable w.r.t. C through X (the kernel's backward);
    the list entries are detached views of the recorded iterates.
    The 4th value (per-sweep ||dX|| of the last outer step) is diagnostic only and never read by the reference's
    callers; it is returned as an empty list."""
    assert C.dim() == 2, C.shape
    stime = time.time()
    R, _, _, _, X, _, n_list, xlist, cost = ops.relax_solve(C[None], None, max_iter=max_iter, proj_iter=proj_iter, lr=lr,
                                                            negate=False, pad_rule=False, is_test=True, want_xlist=True)
    L = int(n_list[0].item())
    X_list = [xlist[0, k] for k in range(L)]
    cost_list = [0] + cost[0, 1:L].tolist()
    etime = time.time() - stime
    if return_time:
        return X[0], cost_list, X_list, [], etime
    return X[0], cost_list, X_list, []


def relax_matching_mean(C, max_iter, proj_iter, lr):
    """sum(X_list)/len(X_list) without materialising the list (what MatchModel consumes, match_model.py:121); differentiable."""
    return ops.relax_solve(C[None], None, max_iter=max_iter, proj_iter=proj_iter, lr=lr, negate=False, pad_rule=False)[0][0]


def hungarian_matching(cost):
    """SciPy LSAP on the host, one-hot back on the device (reference relax_match.py:120-126; non-differentiable)."""
    from scipy.optimize import linear_sum_assignment
    dev = cost.device
    c = cost.detach().cpu().numpy()
    row_ind, col_ind = linear_sum_assignment(c)
    X_0 = np.zeros_like(c)
    X_0[row_ind, col_ind] = 1
    return torch.from_numpy(X_0).float().to(dev), None, None, None
