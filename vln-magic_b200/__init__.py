"""magic_b200 -- B200-native MAGIC pretraining / distillation hot path."""
