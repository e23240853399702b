"""Model zoo: the workloads of the reference's examples, rebuilt on the fused sm_100a ops."""
from .resnet_vd import (ResNetVd, ResNet18_vd, ResNet34_vd, ResNet50_vd, ResNet101_vd,
                        ResNet152_vd, ResNet200_vd, ConvBNAct, to_train_dtype)
from .resnet import ResNet, ResNet18, ResNet34, ResNet50, ResNet101, ResNet152
from .vgg import VGG, VGG11, VGG13, VGG16, VGG19
from .ctr_dnn import CtrDnn, DeepFM

__all__ = ["ResNet", "ResNet18", "ResNet34", "ResNet50", "ResNet101", "ResNet152", "VGG", "VGG11", "VGG13",
           "VGG16", "VGG19", "ResNetVd", "ResNet18_vd", "ResNet34_vd", "ResNet50_vd", "ResNet101_vd", "ResNet152_vd",
           "ResNet200_vd", "ConvBNAct", "to_train_dtype", "CtrDnn", "DeepFM"]
