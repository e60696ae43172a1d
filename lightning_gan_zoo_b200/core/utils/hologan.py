"""Mirror of the reference's `core.utils.hologan` (core/utils/hologan.py:3-9): the `_target_` of
conf/lr_scheduler/hologan.yaml."""
from torch.optim.lr_scheduler import LambdaLR

from ...training import hologan_lr_lambda


def create_hologan_lr_scheduler(total_epochs, optimizer):
    """LambdaLR: factor 1 until epoch <= total_epochs / 2, then linear decay to 0 at total_epochs."""
    return LambdaLR(optimizer, hologan_lr_lambda(total_epochs))
