"""Distillation service: DistillReader data plane, teacher servers, discovery + balancing."""
