"""magic_b200 -- B200-native MAGIC pretraining / distillation hot path.

Public surface (mirrors the reference's module API for this path):
  GlocalTextPathCMTPreTraining      drop-in for pretrain_src/model/pretrain_goat.py (absent upstream)
  kd_loss.{mse_loss, kd_loss, exponential_decay, invert_normalized_losses}   pretrain_src/optim/kd_loss.py
  makd.compute_kd_losses / train_step.PretrainStepper / train_loop.train     map_nav_src/r2r/agent.py:546-719; the loop
                                                                             train_r2r_magic.py sets up but omits
  optim.FusedAdamW / build_optimizer / get_lr_sched                          pretrain_src/optim/*
"""
from .model import GlocalTextPathCMTPreTraining, GlocalTextPathCMT, stack_attns  # noqa: F401
from .graph_index import prepare_batch, batch_to_device, flatten_batch, INDEX_KEY, FLAT_KEY  # noqa: F401
