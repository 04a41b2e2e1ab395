from . import losses
from . import trainer
from .losses import get_batch_loss
from .trainer import (train, train_epoch, train_test_epoch, test_epoch, predict_batch, get_optimizer_and_lr_scheduler,
                      clip_grad_norm_)
