"""Model zoo: the workloads of the reference's examples, rebuilt on the fused sm_100a ops."""
from .resnet_vd import (ResNetVd, ResNet18_vd, ResNet34_vd, ResNet50_vd, ResNet101_vd,
                        ResNet152_vd, ResNet200_vd, ConvBNAct, to_train_dtype)

__all__ = ["ResNetVd", "ResNet18_vd", "ResNet34_vd", "ResNet50_vd", "ResNet101_vd", "ResNet152_vd",
           "ResNet200_vd", "ConvBNAct", "to_train_dtype"]
