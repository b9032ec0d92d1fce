"""Routine to fit SQFA filters with L-BFGS -- drop-in for `sqfa._optim`.

The optimiser driver, stopping rule and progress reporting are host code with the semantics of the
reference (/root/reference/src/sqfa/_optim.py:33-145). The closure body -- the hot loop -- runs
natively: for the built-in distances the whole `get_class_distances -> check -> -mean(tril) ->
backward` sequence is ONE fused native evaluation (`_ops.FusedLoss`), for any other
`distance_fun` the projection is native and the user's function runs on device tensors.
"""

import time

import torch
from torch import optim  # noqa: F401  (kept: the reference module exposes it)
from tqdm import tqdm

from ._lbfgs import LBFGS

__all__ = ["fitting_loop"]


def __dir__():
    return __all__


_NAN_MSG = "Some distances between classes are NaN. Try using float64 or a different regularization parameter."
_INF_MSG = "Some distances between classes are inf. Try using float64 or a different regularization parameter."


def check_distances_valid(distances):
    """
    Check if off-diagonal distances are valid. Raise an error if they are not
    (reference _optim.py:16-30; only the strict lower triangle is inspected).
    """
    n_classes = distances.shape[0]
    i, j = torch.tril_indices(n_classes, n_classes, offset=-1, device=distances.device)
    tril = distances[i, j]
    if torch.isnan(tril).any():
        raise ValueError(_NAN_MSG)
    if torch.isinf(tril).any():
        raise ValueError(_INF_MSG)


def fitting_loop(
    model,
    data_statistics,
    max_epochs=200,
    lr=0.1,
    atol=1e-6,
    show_progress=True,
    return_loss=False,
    **kwargs,
):
    """
    Learn SQFA filters using the LBFGS optimizer.

    Same contract as the reference `fitting_loop` (_optim.py:33-145): `model.parameters()` are
    optimised with `torch.optim.LBFGS(lr=lr, **kwargs)`, one epoch is one `optimizer.step`, training
    stops after 3 consecutive epochs whose loss change is below `atol`. Returns
    `(loss per epoch, elapsed time per epoch)` tensors when `return_loss` is True, else None.
    """
    optimizer = LBFGS(model.parameters(), lr=lr, **kwargs)  # torch.optim.LBFGS, direction update on the device

    if isinstance(data_statistics, dict):
        n_classes = data_statistics["means"].shape[0]
    else:
        n_classes = data_statistics.shape[0]

    fused = model._fused_loss_plan(data_statistics) if hasattr(model, "_fused_loss_plan") else None
    direct = model._fused_direct_plan(data_statistics) if hasattr(model, "_fused_direct_plan") else None
    tril_ind = None

    def closure():
        nonlocal tril_ind
        if direct is not None:
            # native loss + gradient, gradient written to .grad without an autograd graph; ONE host
            # read per evaluation, which also carries max|grad| for the optimiser's first stopping test
            loss_value, bad, grad_absmax = direct().tolist()
            if bad != 0 or loss_value != loss_value:
                raise ValueError(_NAN_MSG if loss_value != loss_value else _INF_MSG)
            optimizer._last_grad_absmax = grad_absmax
            return torch.tensor(loss_value)
        optimizer.zero_grad()
        if fused is not None:
            out = fused()  # [loss, #non-finite pair distances] -- one native evaluation
            loss_value, bad = out.detach().tolist()  # the single host read of this evaluation
            if bad != 0 or loss_value != loss_value:
                raise ValueError(_NAN_MSG if loss_value != loss_value else _INF_MSG)
            epoch_loss = out[0]
            epoch_loss.backward()
            # hand LBFGS a host scalar so its float(loss) does not synchronise again
            return torch.tensor(loss_value)
        distances = model.get_class_distances(data_statistics, regularized=True)
        check_distances_valid(distances)
        if tril_ind is None:
            tril_ind = torch.tril_indices(n_classes, n_classes, offset=-1, device=distances.device)
        epoch_loss = -torch.mean(distances[tril_ind[0], tril_ind[1]])
        epoch_loss.backward()
        return epoch_loss

    loss_list = []
    training_time = []
    total_start_time = time.time()

    prev_loss = 0.0
    consecutive_stopping_criteria_met = 0

    for e in tqdm(range(max_epochs), desc="Epochs", unit="epoch", disable=not show_progress):
        epoch_loss = float(optimizer.step(closure))
        epoch_time = time.time() - total_start_time

        loss_change = abs(prev_loss - epoch_loss)
        if loss_change < atol:
            consecutive_stopping_criteria_met += 1
        else:
            consecutive_stopping_criteria_met = 0

        prev_loss = epoch_loss
        training_time.append(epoch_time)
        loss_list.append(epoch_loss)

        if consecutive_stopping_criteria_met >= 3:
            tqdm.write(
                f"Loss change below {atol} for 3 consecutive epochs. Stopping training at epoch {e + 1}/{max_epochs}."
            )
            break
    else:
        print(
            f"Reached max_epochs ({max_epochs}) without meeting stopping criteria."
            + "Consider increasing max_epochs, changing initialization or using dtype=torch.float64."
        )

    if return_loss:
        return torch.tensor(loss_list), torch.tensor(training_time)
    return None
